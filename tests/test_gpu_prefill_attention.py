"""Prefill attention (csrc/ops.cu k_attention_prefill: 64 x 64 tiles, fp32, online softmax) against the oracle executor's
attention (src/backend/reference.zig:568-672): causal and arbitrary additive masks, ragged seq_q / seq_kv, strided
(column-major) query / key / value / output layouts like the LLaMA lowering emits, all-masked rows."""
import numpy as np
import pytest

from oracle import oracle
from zgml_b200 import DeviceOp, DeviceProgram, ProgramIO

pytestmark = pytest.mark.gpu


def run_both(be, prog, ins, out_buf, n):
    want = np.zeros(n, np.float32)
    oracle.run_program(prog, ins, [ProgramIO(out_buf, want)])
    h = be.compile_program(prog)
    assert h is not None
    got = np.zeros(n, np.float32)
    be.execute_program(h, ins, [ProgramIO(out_buf, got)])
    be.free_program(h)
    return got, want


@pytest.mark.parametrize("dh", [64, 128])
@pytest.mark.parametrize("seq_q,seq_kv,S", [(16, 16, 16), (64, 64, 64), (100, 130, 160), (192, 192, 256), (33, 257, 300)])
@pytest.mark.parametrize("mask_kind", ["causal", "random", "none"])
def test_prefill_attention_matches_oracle(cuda_backend, dh, seq_q, seq_kv, S, mask_kind):
    r = np.random.default_rng(dh + seq_q + seq_kv)
    n_heads = 2
    # buffers: 0 q [T][n_heads*dh] (head h = columns h*dh..), 1 k cache [S][dh] per head slab, 2 v cache, 3 mask [S x T] (column per query), 4 out
    qb = r.standard_normal(seq_q * n_heads * dh).astype(np.float32)
    kb = r.standard_normal(n_heads * S * dh).astype(np.float32)
    vb = r.standard_normal(n_heads * S * dh).astype(np.float32)
    mask = np.zeros((seq_q, S), np.float32)
    if mask_kind == "causal":
        off = seq_kv - seq_q
        for i in range(seq_q):
            mask[i, max(i + off + 1, 0):] = -np.inf
    elif mask_kind == "random":
        mask = np.where(r.random((seq_q, S)) < 0.3, -np.inf, r.standard_normal((seq_q, S)) * 0.1).astype(np.float32)
        mask[3, :] = -np.inf          # a fully masked query row: zeros out
    ops = []
    for h in range(n_heads):
        ops.append(DeviceOp.attention(4, 0, 1, 2, 3, mask_kind != "none", dh, seq_q, seq_kv, float(1.0 / np.sqrt(dh)),
                                      h * dh, h * S * dh, h * S * dh, 0, h * dh, 1, n_heads * dh, 1, dh, 1, dh, 1, S, 1, n_heads * dh))
    prog = DeviceProgram(ops, [qb.size, kb.size, vb.size, mask.size, seq_q * n_heads * dh], [], [])
    ins = [ProgramIO(0, qb), ProgramIO(1, kb), ProgramIO(2, vb), ProgramIO(3, np.ascontiguousarray(mask).ravel())]
    got, want = run_both(cuda_backend, prog, ins, 4, seq_q * n_heads * dh)
    assert np.isfinite(got).all()
    assert np.max(np.abs(got - want)) <= 2e-5 * max(1.0, float(np.max(np.abs(want))))
    if mask_kind == "random":
        assert np.all(got.reshape(seq_q, n_heads * dh)[3] == 0.0)
