"""GPU parity tests of the W8A8 decode path (prepareTransposed, quantizeInput, gemv / gemvRange — reference
src/quant.zig:274-459) through the C-ABI, against the CPU oracle on the same seeded inputs.

The whole path is integer / IEEE-exact arithmetic with a fixed float accumulation order (K-blocks ascending per
output), so the bar is BIT-EXACT: transposed int8 data and scales, quantized activations, and the gemv outputs.
The reference's own envelope tests (src/quant.zig:1165-1209: gemv vs matmul and vs the float product, 0.15) are
repeated on top."""
import numpy as np
import pytest

from oracle import oracle
from zgml_b200 import BackendError, QuantizedWeight
from zgml_b200.backend import quantize_input

pytestmark = pytest.mark.gpu


def rng(seed):
    return np.random.default_rng(seed)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def make_pair(be, K, N, bs, seed):
    w = rng(seed).uniform(-1, 1, K * N).astype(np.float32)
    o = oracle.QuantizedWeight.from_slice(w, K, N, bs)
    g = QuantizedWeight.upload(be, o.data, o.scales, K, N, bs)
    return w, o, g


SHAPES = [(64, 64, 32), (100, 37, 32), (96, 160, 32), (576, 1536, 32), (2048, 512, 64), (128, 96, 16), (4096, 512, 128),
          (1024, 64, 512), (512, 40, 256), (33, 7, 16), (9, 12, 4), (48, 20, 48)]


@pytest.mark.parametrize("K,N,bs", SHAPES)
def test_prepare_transposed_bit_identical(cuda_backend, K, N, bs):
    _, o, g = make_pair(cuda_backend, K, N, bs, K + N)
    o.prepare_transposed()
    t_data, t_scales = g.prepare_transposed(return_host=True)
    assert np.array_equal(t_data, o.t_data)
    assert np.array_equal(bits(t_scales), bits(o.t_scales))
    again = g.prepare_transposed(return_host=True)          # idempotent
    assert np.array_equal(again[0], t_data)
    g.free()


@pytest.mark.parametrize("ggml_type", [8, 2])
def test_prepare_transposed_from_gguf_residency(cuda_backend, ggml_type):
    """The transposed form is built from the packed records (Q8_0 / Q4_0 residency), not from a host copy."""
    K, N = 576, 192
    r = rng(3)
    nb, payload = K * N // 32, (32 if ggml_type == 8 else 16)
    raw = np.zeros((nb, 2 + payload), np.uint8)              # GGUF blocks: f16 scale + 32 x i8 (Q8_0) / 16 nibble bytes (Q4_0)
    raw[:, 0:2] = r.uniform(1e-3, 1e-2, nb).astype(np.float16).view(np.uint8).reshape(nb, 2)
    raw[:, 2:] = r.integers(0, 256, (nb, payload)).astype(np.uint8)
    if ggml_type == 8:
        raw[:, 2:][raw[:, 2:] == 0x80] = 0x81                # keep q in [-127, 127]
    raw = raw.ravel()
    o = oracle.QuantizedWeight.from_gguf(raw, ggml_type, K, N)
    g = QuantizedWeight.from_gguf_blocks(cuda_backend, raw, ggml_type, K, N)
    o.prepare_transposed()
    t_data, t_scales = g.prepare_transposed(return_host=True)
    assert np.array_equal(t_data, o.t_data) and np.array_equal(bits(t_scales), bits(o.t_scales))
    x = rng(5).standard_normal(K).astype(np.float32)
    assert np.array_equal(bits(g.gemv(x)), bits(o.gemv(x)))
    g.free()


@pytest.mark.parametrize("K,bs", [(64, 32), (100, 32), (4096, 32), (16384, 32), (2048, 64), (33, 16), (7, 32), (512, 1)])
def test_quantize_input_bit_identical(cuda_backend, K, bs):
    x = rng(K + bs).standard_normal(K).astype(np.float32)
    x[:: max(1, K // 7)] = 0.0
    if K >= 64:
        x[32:64] = 0.0                                      # an all-zero block: scale 1, q 0
    q, s = quantize_input(cuda_backend, x, bs)
    oq, os_ = oracle.quantize_input(x, bs)
    assert np.array_equal(q, oq) and np.array_equal(bits(s), bits(os_))


@pytest.mark.parametrize("K,N,bs", SHAPES + [(4096, 4096, 32), (16384, 64, 32)])
def test_gemv_w8a8_bit_identical(cuda_backend, K, N, bs):
    w, o, g = make_pair(cuda_backend, K, N, bs, 3 * K + N)
    o.prepare_transposed()
    g.prepare_transposed()
    for seed in (1, 2):
        x = rng(seed).standard_normal(K).astype(np.float32)
        got, want = g.gemv(x), o.gemv(x)
        assert np.array_equal(bits(got), bits(want))
    # the reference's envelope (src/quant.zig:1165-1209): W8A8 vs the W8.f32 matmul and vs the float product
    y_mm = g.matmul(x, 1)[0]
    y_f = x.astype(np.float64) @ w.reshape(K, N).astype(np.float64)
    tol = 0.08 * float(np.max(np.abs(y_f))) + 0.05          # two int8 roundings of the weights + one of the activations
    assert np.max(np.abs(got - y_mm)) < tol and np.max(np.abs(got - y_f)) < tol
    g.free()


def test_gemv_w8a8_contract_errors(cuda_backend):
    _, o, g = make_pair(cuda_backend, 64, 32, 32, 1)
    with pytest.raises(BackendError):                       # no transposed form yet (src/quant.zig:444-445 asserts it)
        g.gemv(np.ones(64, np.float32))
    g.free()
    _, o, g = make_pair(cuda_backend, 16384 + 32, 32, 32, 2)
    g.prepare_transposed()
    with pytest.raises(BackendError):                       # the reference's stack buffers: K <= 16384
        g.gemv(np.ones(16384 + 32, np.float32))
    with pytest.raises(BackendError):
        quantize_input(cuda_backend, np.ones(1024, np.float32), 1)   # > 512 blocks
    g.free()


def test_gemv_w8a8_zero_and_extreme_inputs(cuda_backend):
    K, N = 256, 96
    _, o, g = make_pair(cuda_backend, K, N, 32, 9)
    o.prepare_transposed()
    g.prepare_transposed()
    for x in (np.zeros(K, np.float32), np.full(K, 1e20, np.float32), np.full(K, -1e-30, np.float32)):
        assert np.array_equal(bits(g.gemv(x)), bits(o.gemv(x)))
    g.free()
