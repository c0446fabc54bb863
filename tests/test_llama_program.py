"""CPU tests of the host-side LLaMA lowering (zgml_b200/host/llama.py) against the oracle executor:
structure of the emitted DeviceProgram and prefill == step on the reference semantics."""
import numpy as np

from llama_reference import OracleBackend
from zgml_b200 import abi
from zgml_b200.host.llama import SMOLLM_135M, DeviceLlamaSession, LlamaConfig, build_program, linear_shapes, rope_tables, synthetic_weights

TINY = LlamaConfig(vocab_size=64, d_model=32, n_layers=2, n_heads=4, n_kv_heads=2, d_ff=64, max_seq_len=16)


def test_program_structure_matches_reference_lowering():
    w = synthetic_weights(TINY, "q8_0", seed=1)
    lp = build_program(TINY, w, 1)
    tags = [o.tag for o in lp.program.ops]
    assert tags.count(abi.OP_QMATMUL) == 7 * TINY.n_layers          # q,k,v,o,gate,up,down (llama_transformer.zig:204-206,240,130-132)
    assert tags.count(abi.OP_ATTENTION) == TINY.n_heads * TINY.n_layers
    assert tags.count(abi.OP_ROPE) == (TINY.n_heads + TINY.n_kv_heads) * TINY.n_layers
    assert tags.count(abi.OP_MATMUL) == 1                           # tied LM head stays dense f32
    assert len(lp.slice_assign_ops) == 2 * TINY.n_kv_heads * TINY.n_layers
    for i in lp.slice_assign_ops:
        assert lp.program.ops[i].u.slice_assign.patch_stride == TINY.d_head
    assert lp.n_qmatmul == len(lp.program.qweights)
    shapes = linear_shapes(SMOLLM_135M)
    assert sum(k * n for k, n in shapes.values()) * SMOLLM_135M.n_layers == 106_168_320  # 106.2 M quantized weights


def test_rope_table_matches_reference_formula():
    cos, sin = rope_tables(TINY)
    d = TINY.d_head
    p, i = 5, 3
    f = np.float32(p) / np.power(np.float32(TINY.rope_base), np.float32(2 * i) / np.float32(d), dtype=np.float32)
    assert cos[p, i] == np.cos(np.float32(f)).astype(np.float32) and cos[p, i + d // 2] == cos[p, i]
    assert sin[p, i] == np.sin(np.float32(f)).astype(np.float32)


def test_prefill_equals_step_on_reference_executor():
    w = synthetic_weights(TINY, "q4_0", seed=2, embed_scale=1.0)
    toks = [1, 7, 33]
    s = DeviceLlamaSession(OracleBackend(), TINY, w, 1)
    for t in toks:
        last = s.step(t).copy()
    p = DeviceLlamaSession(OracleBackend(), TINY, w, len(toks))
    got = p.execute_at(toks, 0).copy()
    s.close(); p.close()
    assert np.max(np.abs(got - last)) <= 1e-4 * np.max(np.abs(last))
    assert np.isfinite(got).all() and np.ptp(got) > 0
