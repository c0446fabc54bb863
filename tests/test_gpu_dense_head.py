"""bf16 LM-head matvec with the argmax-safe exact recompute (csrc/dense_head.cu, zg_cuda_program_promote_dense; SURVEY.md §8f-4,
precedent src/backend/wgpu.zig:1068-1106).  Bars: the argmax and the winning logit are the unpromoted program's, bit for bit;
every other logit within 2e-4 (f16) / 2e-3 (bf16) of the output scale — f16 stays inside the 1e-3 whole-program logit budget; non-finite activations behave like the
exact path; an input that overwrites the operand refreshes the copy."""
import numpy as np
import pytest

from zgml_b200 import DeviceOp, DeviceProgram, ProgramIO
from zgml_b200.host.llama import DeviceLlamaSession, LlamaConfig, synthetic_weights

pytestmark = pytest.mark.gpu


def head_program(W, K, N):
    # buffers: 0 x [K], 1 W [N, K] (k-contiguous rows: the tied head x @ token_embed^T), 2 logits [N]
    op = DeviceOp.matmul(2, 0, 1, 1, N, K, K, 1, 1, K)
    return DeviceProgram([op], [K, N * K, N], [ProgramIO(1, W)], [])


def run(be, h, x, N, extra_inputs=()):
    out = np.zeros(N, np.float32)
    be.execute_program(h, [ProgramIO(0, x)] + list(extra_inputs), [ProgramIO(2, out)])
    return out


@pytest.mark.parametrize("fmt,tol", [("f16", 2e-4), ("bf16", 2e-3)])
@pytest.mark.parametrize("K,N", [(576, 4096), (2048, 8192), (64, 1024)])
def test_16bit_head_keeps_argmax_and_winning_logit(cuda_backend, K, N, fmt, tol):
    r = np.random.default_rng(K + N)
    W = (r.standard_normal((N, K)) * 0.05).astype(np.float32)
    W[5] = W[900]                                   # an exact tie among the rows
    prog = head_program(W.ravel(), K, N)
    exact, fast = cuda_backend.compile_program(prog), cuda_backend.compile_program(prog)
    assert cuda_backend.promote_dense_weights(fast, fmt) == 1
    assert cuda_backend.promote_dense_weights(fast, fmt) == 0          # idempotent
    for trial in range(12):
        x = r.standard_normal(K).astype(np.float32) * (10.0 ** r.integers(-3, 4))
        if trial % 3 == 0:
            x = (W[r.integers(0, N)] * 40 + r.standard_normal(K) * 0.3).astype(np.float32)   # a clear winner
        if trial % 4 == 1:
            x = (W[5] * 25).astype(np.float32)                                                 # the tied rows win together
        y0, y1 = run(cuda_backend, exact, x, N), run(cuda_backend, fast, x, N)
        top = int(np.argmax(y0))
        assert int(np.argmax(y1)) == top
        assert y1[top].view(np.uint32) == y0[top].view(np.uint32)
        assert y1[5].view(np.uint32) == y1[900].view(np.uint32) or abs(y0[5] - y0[top]) > 0.1 * abs(y0[top])
        assert np.max(np.abs(y1 - y0)) <= tol * np.max(np.abs(y0))
        # every column the bound could not exclude carries the exact bits: in particular all columns within 1e-4 of the top
        near = np.abs(y0 - y0[top]) <= 1e-4 * abs(y0[top])
        assert np.array_equal(y1[near].view(np.uint32), y0[near].view(np.uint32))
    st = cuda_backend.program_stats(fast)
    assert st["kernels"] == 1 and st["dense_bytes_saved_per_execution"] == 2 * N * K - 6 * N
    cuda_backend.free_program(exact); cuda_backend.free_program(fast)


def test_bf16_head_degenerate_and_non_finite_inputs(cuda_backend):
    K, N = 256, 2048
    r = np.random.default_rng(1)
    W = np.tile((r.standard_normal(K) * 0.1).astype(np.float32), (N, 1))        # all rows equal: nothing can be excluded
    prog = head_program(W.ravel(), K, N)
    exact, fast = cuda_backend.compile_program(prog), cuda_backend.compile_program(prog)
    assert cuda_backend.promote_dense_weights(fast) == 1
    x = r.standard_normal(K).astype(np.float32)
    assert np.array_equal(run(cuda_backend, fast, x, N).view(np.uint32), run(cuda_backend, exact, x, N).view(np.uint32))
    assert not run(cuda_backend, fast, np.zeros(K, np.float32), N).any()
    x[7] = np.nan
    assert np.isnan(run(cuda_backend, fast, x, N)).all()
    x[7] = np.inf
    y0, y1 = run(cuda_backend, exact, x, N), run(cuda_backend, fast, x, N)
    assert np.array_equal(np.isnan(y0), np.isnan(y1)) and np.array_equal(y0[~np.isnan(y0)], y1[~np.isnan(y1)])
    cuda_backend.free_program(exact); cuda_backend.free_program(fast)


def test_bf16_head_copy_follows_an_overwritten_operand(cuda_backend):
    K, N = 128, 1024
    r = np.random.default_rng(2)
    W1 = (r.standard_normal((N, K)) * 0.05).astype(np.float32)
    W2 = (r.standard_normal((N, K)) * 0.05).astype(np.float32)
    fast = cuda_backend.compile_program(head_program(W1.ravel(), K, N))
    exact2 = cuda_backend.compile_program(head_program(W2.ravel(), K, N))
    assert cuda_backend.promote_dense_weights(fast) == 1
    x = (W2[77] * 30).astype(np.float32)
    y1 = run(cuda_backend, fast, x, N, [ProgramIO(1, W2.ravel())])            # the operand arrives as an execute input
    y0 = run(cuda_backend, exact2, x, N)
    assert int(np.argmax(y1)) == int(np.argmax(y0)) == 77 and y1[77].view(np.uint32) == y0[77].view(np.uint32)
    assert np.max(np.abs(y1 - y0)) <= 1e-3 * np.max(np.abs(y0))
    cuda_backend.free_program(fast); cuda_backend.free_program(exact2)


def test_rejects_operands_the_program_writes_and_small_heads(cuda_backend):
    K, N = 128, 1024
    W = np.zeros(N * K, np.float32)
    # the operand is produced by an op of the program: not a constant, not promoted
    ops = [DeviceOp.elementwise("add", 1, 1, 1, N * K), DeviceOp.matmul(2, 0, 1, 1, N, K, K, 1, 1, K)]
    h = cuda_backend.compile_program(DeviceProgram(ops, [K, N * K, N], [ProgramIO(1, W)], []))
    assert cuda_backend.promote_dense_weights(h) == 0
    cuda_backend.free_program(h)
    h = cuda_backend.compile_program(head_program(np.zeros(512 * K, np.float32), K, 512))   # below 1024 columns
    assert cuda_backend.promote_dense_weights(h) == 0
    cuda_backend.free_program(h)


@pytest.mark.parametrize("fmt,tol", [("f16", 1e-3), ("bf16", 4e-3)])
def test_tied_head_model_decodes_the_same_tokens(cuda_backend, fmt, tol):
    cfg = LlamaConfig(vocab_size=2048, d_model=256, n_layers=2, n_heads=4, n_kv_heads=2, d_ff=512, max_seq_len=32)   # tied LM head
    w = synthetic_weights(cfg, "q8_0", seed=3, embed_scale=1.0)
    a, b = DeviceLlamaSession(cuda_backend, cfg, w), DeviceLlamaSession(cuda_backend, cfg, w)
    assert cuda_backend.promote_dense_weights(b.handle, fmt) == 1
    ta = tb = 1
    for _ in range(8):
        la, lb = a.step(ta).copy(), b.step(tb).copy()
        assert int(np.argmax(la)) == int(np.argmax(lb))
        assert np.max(np.abs(la - lb)) <= tol * np.max(np.abs(la))   # f16: inside the whole-program budget; bf16: argmax-safe only
        ta = tb = int(np.argmax(la))
    a.close(); b.close()
