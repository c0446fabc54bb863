"""Whole-program parity: LLaMA decode / prefill DevicePrograms (zgml_b200/host/llama.py, the mirror of
src/device_inference.zig's lowering of LLaMA.forwardCachedMasked) on the CUDA backend vs the oracle
executor on the same synthetic GGUF-direct weights.  Bars (BASELINE.json north_star): logits within
1e-3 relative, greedy argmax tokens identical."""
import numpy as np
import pytest

from llama_reference import OracleBackend
from zgml_b200.host.llama import SMOLLM_135M, DeviceLlamaSession, LlamaConfig, synthetic_weights

pytestmark = pytest.mark.gpu

TINY = LlamaConfig(vocab_size=256, d_model=64, n_layers=2, n_heads=4, n_kv_heads=2, d_ff=128, max_seq_len=32)
TINY_UNTIED = LlamaConfig(vocab_size=256, d_model=64, n_layers=2, n_heads=4, n_kv_heads=4, d_ff=160, max_seq_len=16,
                          rope_base=5e5, tied_lm_head=False)


def rel(got, want):
    return float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64))) / (np.max(np.abs(want)) + 1e-30))


def greedy(sess, first, n):
    toks, logs, t = [], [], first
    for _ in range(n):
        lg = sess.step(t).copy()
        t = int(np.argmax(lg))
        toks.append(t)
        logs.append(lg)
    return toks, logs


@pytest.mark.parametrize("cfg", [TINY, TINY_UNTIED], ids=["tied-gqa", "untied-mha"])
@pytest.mark.parametrize("kind", ["q8_0", "q4_0"])
@pytest.mark.parametrize("graph", [True, False], ids=["graph", "eager"])
def test_tiny_llama_greedy_decode_matches_oracle(cuda_backend, cfg, kind, graph):
    w = synthetic_weights(cfg, kind, seed=3, embed_scale=1.0)
    cuda_backend.set_graph_mode(graph)
    dev, ref = DeviceLlamaSession(cuda_backend, cfg, w), DeviceLlamaSession(OracleBackend(), cfg, w)
    got_t, got_l = greedy(dev, 1, 8)
    want_t, want_l = greedy(ref, 1, 8)
    dev.close(); ref.close()
    cuda_backend.set_graph_mode(True)
    assert got_t == want_t
    for g, wv in zip(got_l, want_l):
        assert rel(g, wv) < 1e-3


def test_prefill_chunk_equals_token_by_token(cuda_backend):  # reference llama_inference.zig:983-1107 (prefill == step, 1e-4)
    cfg = TINY
    w = synthetic_weights(cfg, "q8_0", seed=5, embed_scale=1.0)
    toks = [3, 17, 250, 9]
    step = DeviceLlamaSession(cuda_backend, cfg, w, 1)
    for t in toks:
        last = step.step(t).copy()
    pre = DeviceLlamaSession(cuda_backend, cfg, w, len(toks))
    got = pre.execute_at(toks, 0).copy()
    ref = DeviceLlamaSession(OracleBackend(), cfg, w, len(toks))
    want = ref.execute_at(toks, 0).copy()
    step.close(); pre.close(); ref.close()
    assert rel(got, want) < 1e-3
    assert rel(got, last) < 1e-3 and int(np.argmax(got)) == int(np.argmax(last))


def test_smollm_135m_shape_q8_0_greedy_decode(cuda_backend):  # BASELINE.json configs[0] shapes, 3 greedy tokens
    cfg = SMOLLM_135M
    w = synthetic_weights(cfg, "q8_0", seed=0)
    dev, ref = DeviceLlamaSession(cuda_backend, cfg, w), DeviceLlamaSession(OracleBackend(native=True), cfg, w)
    got_t, got_l = greedy(dev, 1, 3)
    want_t, want_l = greedy(ref, 1, 3)
    dev.close(); ref.close()
    assert got_t == want_t
    for g, wv in zip(got_l, want_l):
        assert rel(g, wv) < 1e-3


def test_prefill_chunk_on_tensor_core_path(cuda_backend):  # token_len > 8: every linear runs the tcgen05 GEMM
    cfg = TINY
    w = synthetic_weights(cfg, "q4_0", seed=9, embed_scale=1.0)
    toks = [(7 * i + 3) % cfg.vocab_size for i in range(24)]
    pre = DeviceLlamaSession(cuda_backend, cfg, w, len(toks))
    got = pre.execute_at(toks, 0).copy()
    ref = DeviceLlamaSession(OracleBackend(), cfg, w, len(toks))
    want = ref.execute_at(toks, 0).copy()
    pre.close(); ref.close()
    assert rel(got, want) < 1e-3 and int(np.argmax(got)) == int(np.argmax(want))


@pytest.mark.parametrize("token_len", [1, 3])
def test_every_program_buffer_matches_oracle_with_fused_patterns(cuda_backend, token_len):
    """The backend evaluates the lowering's fixed op runs (add/rmsnorm/repeat/mul, SiLU chain * up, attention + head
    store, chained small ops) in single passes; every DeviceOp's output buffer must still hold what op-by-op
    execution (the oracle executor) leaves there — not just the logits."""
    from zgml_b200 import ProgramIO
    cfg = TINY_UNTIED
    w = synthetic_weights(cfg, "q8_0", seed=21, embed_scale=1.0)
    dev, ref = DeviceLlamaSession(cuda_backend, cfg, w, token_len), DeviceLlamaSession(OracleBackend(), cfg, w, token_len)
    sizes = dev.lp.program.buffer_sizes
    got = [np.zeros(n, np.float32) for n in sizes]
    want = [np.zeros(n, np.float32) for n in sizes]
    dev.outputs = [ProgramIO(i, got[i]) for i in range(len(sizes))]
    ref.outputs = [ProgramIO(i, want[i]) for i in range(len(sizes))]
    toks = [5, 77, 130][:token_len]
    for pos in (0, token_len):
        dev.execute_at(toks, pos)
        ref.execute_at(toks, pos)
        for i, (g, wv) in enumerate(zip(got, want)):
            scale = float(np.max(np.abs(wv[np.isfinite(wv)]))) if np.isfinite(wv).any() else 0.0
            fin = np.isfinite(wv)
            assert np.array_equal(np.isfinite(g), fin), f"buffer {i}"
            assert np.max(np.abs(g[fin] - wv[fin]), initial=0.0) <= 1e-4 * scale + 1e-6, f"buffer {i} at pos {pos}"
    dev.close(); ref.close()


@pytest.mark.parametrize("knobs", [{"ZG_CUDA_GEMV_FUSE": "3"}, {"ZG_CUDA_CHAIN": "0", "ZG_CUDA_FUSE": "0", "ZG_CUDA_GEMV_BATCH": "1", "ZG_CUDA_ATTN_SPLIT": "0", "ZG_CUDA_PDL": "0"}],
                         ids=["matvec-prologue-fusion", "every-scheduling-feature-off"])
def test_scheduling_knobs_do_not_change_results(knobs):
    """The scheduling features are switches read at context creation: the optional matvec-prologue fusion
    (off by default, measured slower) and the plain one-launch-per-op configuration must give the same program buffers."""
    import os
    from zgml_b200 import CudaBackend, ProgramIO
    old = {k: os.environ.get(k) for k in knobs}
    os.environ.update(knobs)
    try:
        be = CudaBackend(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    try:
        cfg = TINY_UNTIED
        w = synthetic_weights(cfg, "q4_0", seed=31, embed_scale=1.0)
        dev, ref = DeviceLlamaSession(be, cfg, w, 1), DeviceLlamaSession(OracleBackend(), cfg, w, 1)
        sizes = dev.lp.program.buffer_sizes
        got = [np.zeros(n, np.float32) for n in sizes]
        want = [np.zeros(n, np.float32) for n in sizes]
        dev.outputs = [ProgramIO(i, got[i]) for i in range(len(sizes))]
        ref.outputs = [ProgramIO(i, want[i]) for i in range(len(sizes))]
        for pos, tok in enumerate([9, 200, 31]):
            dev.execute_at([tok], pos)
            ref.execute_at([tok], pos)
            for i, (g, wv) in enumerate(zip(got, want)):
                fin = np.isfinite(wv)
                scale = float(np.max(np.abs(wv[fin]))) if fin.any() else 0.0
                assert np.array_equal(np.isfinite(g), fin), f"buffer {i}"
                assert np.max(np.abs(g[fin] - wv[fin]), initial=0.0) <= 1e-4 * scale + 1e-6, f"buffer {i} at pos {pos}"
        dev.close(); ref.close()
    finally:
        be.close()
