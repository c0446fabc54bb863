"""GPU parity tests proper: the CUDA path, called through the C-ABI of
include/zgml_cuda.h, against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): dequantized weights BIT-EXACT; matvec / matmul
outputs within fp32 relative tolerance 1e-3 (measured against the output scale
max|y|; the kernels land near 1e-6).
"""
import struct

import numpy as np
import pytest

from oracle import oracle
from zgml_b200 import DeviceOp, DeviceProgram, ProgramIO, QuantizedWeight, QuantizedWeightUpload, abi

pytestmark = pytest.mark.gpu

REL_TOL = 1e-3  # north_star tolerance on matvec/matmul outputs


def rng(seed):
    return np.random.default_rng(seed)


def rel_err(got, want):
    return float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64))) / (np.max(np.abs(want)) + 1e-30))


def make_q8_0_raw(K, N, seed):
    """GGUF Q8_0 blocks over flat [K,N] (reference src/gguf.zig type 8: f16 scale + 32 x i8)."""
    r = rng(seed)
    nb = K * N // 32
    raw = np.zeros((nb, 34), np.uint8)
    sc = r.uniform(1e-3, 1e-2, nb).astype(np.float16)
    raw[:, 0:2] = sc.view(np.uint8).reshape(nb, 2)
    raw[:, 2:] = r.integers(-127, 128, (nb, 32)).astype(np.int8).view(np.uint8)
    return raw.ravel()


def make_q4_0_raw(K, N, seed):
    """GGUF Q4_0 blocks (type 2: f16 scale + 16 nibble bytes), zgml nibble order on decode."""
    r = rng(seed)
    nb = K * N // 32
    raw = np.zeros((nb, 18), np.uint8)
    sc = r.uniform(1e-3, 1e-2, nb).astype(np.float16)
    raw[:, 0:2] = sc.view(np.uint8).reshape(nb, 2)
    raw[:, 2:] = r.integers(0, 256, (nb, 16)).astype(np.uint8)
    return raw.ravel()


def oracle_out(qw, x, M):
    dst = np.zeros(M * qw.cols, np.float32)
    qw.qmatmul_op(np.ascontiguousarray(x, np.float32).ravel(), dst, M)
    return dst.reshape(M, qw.cols)


# ── golden vectors through the CUDA path ──────────────────────────────────────
def test_metal_qmatvec_vector_on_cuda(cuda_backend):  # reference src/backend/metal.zig:6666-6712
    w = QuantizedWeight.upload(cuda_backend, np.arange(1, 7, dtype=np.int8), np.array([1], np.float32), 3, 2, 6)
    assert w.format == abi.QFMT_GENERIC
    out = w.matmul(np.array([10, 20, 30], np.float32), 1)
    np.testing.assert_allclose(out.ravel(), [220, 280], atol=1e-4)
    w.free()


def test_reference_executor_vector_on_cuda(cuda_backend):  # reference src/backend/reference.zig:710-761
    w = QuantizedWeight.upload(cuda_backend, np.array([2, -1, 3, 4, -2, 1, -3, 5, 2], np.int8),
                               np.array([0.5, 0.25, 1.0], np.float32), 3, 3, 4)
    out = w.matmul(np.array([1, 2, 3, -1, 0.5, 4], np.float32), 2)
    np.testing.assert_allclose(out.ravel(), [2.75, 2.25, 8.0, -3.0, 5.25, 6.625], atol=1e-6)
    w.free()


def test_gguf_block_vectors_on_cuda(cuda_backend):  # reference src/models/gguf_loader.zig:484-572
    raw = bytearray(18)
    raw[0:2] = struct.pack("<e", 0.25)
    raw[2] = 0xF8
    w = QuantizedWeight.from_gguf_blocks(cuda_backend, np.frombuffer(bytes(raw), np.uint8), 2, 16, 2)
    deq = w.dequantize_to().ravel()
    assert deq[0] == 0.0 and deq[1] == 7 * 0.25 and np.all(deq[2:] == -8 * 0.25)
    w.free()
    raw = bytearray(34)
    raw[0:2] = struct.pack("<e", 0.5)
    raw[2], raw[3] = 10, (-5) & 0xFF
    w = QuantizedWeight.from_gguf_blocks(cuda_backend, np.frombuffer(bytes(raw), np.uint8), 8, 16, 2)
    deq = w.dequantize_to().ravel()
    assert deq[0] == 5.0 and deq[1] == -2.5 and np.all(deq[2:] == 0.0)
    w.free()


# ── dequantized weights: bit-exact, every residency format ────────────────────
@pytest.mark.parametrize("K,N", [(64, 64), (1, 32), (3, 96), (100, 160), (129, 32), (576, 1536), (2048, 2048)])
def test_dequant_bit_exact_native_i8_f32(cuda_backend, K, N):
    wf = rng(K * 7 + N).uniform(-1, 1, K * N).astype(np.float32)
    o = oracle.QuantizedWeight.from_slice(wf, K, N, 32)
    w = QuantizedWeight.upload(cuda_backend, o.data, o.scales, K, N, 32)
    assert w.format == abi.QFMT_I8_F32
    assert w.device_bytes >= K * N // 32 * 36
    assert np.array_equal(w.dequantize_to().view(np.uint32), o.dequantize_to().view(np.uint32))
    w.free()


@pytest.mark.parametrize("K,N", [(64, 64), (2, 32), (65, 96), (576, 576), (1536, 576), (4096, 1024)])
@pytest.mark.parametrize("ggml_type,fmt", [(8, abi.QFMT_I8_F16), (2, abi.QFMT_I4_F16)])
def test_dequant_bit_exact_gguf(cuda_backend, K, N, ggml_type, fmt):
    raw = make_q8_0_raw(K, N, K + N) if ggml_type == 8 else make_q4_0_raw(K, N, K + N)
    o = oracle.QuantizedWeight.from_gguf(raw, ggml_type, K, N)
    want = (oracle.dequant_q8_0 if ggml_type == 8 else oracle.dequant_q4_0)(raw, K * N)  # gguf_loader.zig:33-71
    assert np.array_equal(o.dequantize_to().ravel().view(np.uint32), want.view(np.uint32))
    # (a) raw GGUF blocks expanded on device
    w = QuantizedWeight.from_gguf_blocks(cuda_backend, raw, ggml_type, K, N)
    assert w.format == fmt
    assert np.array_equal(w.dequantize_to().ravel().view(np.uint32), want.view(np.uint32))
    w.free()
    # (b) the reference's host-expanded i8 + f32 arrays through the QuantizedWeightUpload boundary
    w = QuantizedWeight.upload(cuda_backend, o.data, o.scales, K, N, 32)
    assert w.format == fmt  # lossless narrowing detected at upload
    assert np.array_equal(w.dequantize_to().ravel().view(np.uint32), want.view(np.uint32))
    w.free()


@pytest.mark.parametrize("K,N,bs", [(5, 6, 2), (9, 12, 4), (4, 6, 6), (7, 48, 32), (33, 40, 8), (64, 64, 64), (16, 256, 128)])
def test_dequant_bit_exact_generic_blocks(cuda_backend, K, N, bs):
    wf = rng(K + N + bs).uniform(-2, 2, K * N).astype(np.float32)
    o = oracle.QuantizedWeight.from_slice(wf, K, N, bs)
    w = QuantizedWeight.upload(cuda_backend, o.data, o.scales, K, N, bs)
    assert w.format == abi.QFMT_GENERIC
    assert np.array_equal(w.dequantize_to().view(np.uint32), o.dequantize_to().view(np.uint32))
    w.free()


def test_format_hint_cannot_be_lossy(cuda_backend):
    from zgml_b200 import BackendError
    o = oracle.QuantizedWeight.from_slice(rng(0).uniform(-1, 1, 64 * 64).astype(np.float32), 64, 64, 32)
    with pytest.raises(BackendError):
        QuantizedWeight.upload(cuda_backend, o.data, o.scales, 64, 64, 32, fmt_hint=abi.QFMT_I4_F16)
    w = QuantizedWeight.upload(cuda_backend, o.data, o.scales, 64, 64, 32, fmt_hint=abi.QFMT_GENERIC)
    assert w.format == abi.QFMT_GENERIC
    w.free()


# ── matvec / matmul vs the oracle ─────────────────────────────────────────────
SHAPES = [(64, 64), (1, 32), (3, 96), (100, 160), (129, 32), (576, 576), (576, 1536), (1536, 576), (2048, 8192), (4096, 4096)]


@pytest.mark.parametrize("K,N", SHAPES)
@pytest.mark.parametrize("M", [1, 2, 3, 4, 5, 8])
def test_qmatmul_native_vs_oracle(cuda_backend, K, N, M):
    if K * N * M > 40e6 and M > 4:
        pytest.skip("oracle time")
    r = rng(K + 3 * N + M)
    o = oracle.QuantizedWeight.from_slice(r.uniform(-1, 1, K * N).astype(np.float32), K, N, 32)
    x = r.standard_normal((M, K)).astype(np.float32)
    w = QuantizedWeight.upload(cuda_backend, o.data, o.scales, K, N, 32)
    got = w.matmul(x, M)
    want = oracle_out(o, x, M)
    assert rel_err(got, want) < REL_TOL
    assert rel_err(got, want) < 2e-5  # what the fp32 kernel actually achieves
    w.free()


@pytest.mark.parametrize("K,N", [(64, 64), (65, 96), (576, 1536), (2048, 2048), (4096, 4096)])
@pytest.mark.parametrize("ggml_type", [8, 2])
@pytest.mark.parametrize("M", [1, 2, 7])
def test_qmatmul_gguf_vs_oracle(cuda_backend, K, N, ggml_type, M):
    raw = make_q8_0_raw(K, N, 11 + K) if ggml_type == 8 else make_q4_0_raw(K, N, 11 + K)
    o = oracle.QuantizedWeight.from_gguf(raw, ggml_type, K, N)
    x = rng(M + K).standard_normal((M, K)).astype(np.float32)
    w = QuantizedWeight.from_gguf_blocks(cuda_backend, raw, ggml_type, K, N)
    got = w.matmul(x, M)
    want = oracle_out(o, x, M)
    assert rel_err(got, want) < 2e-5
    # QuantizedWeight.matmul (src/quant.zig:475-578) and the DeviceOp restatement agree when N % 32 == 0
    assert rel_err(o.matmul(x, M), want) < 1e-5
    w.free()


@pytest.mark.parametrize("K,N,bs", [(5, 6, 2), (9, 12, 4), (4, 6, 6), (7, 48, 32), (33, 40, 8), (64, 64, 64), (256, 256, 128)])
def test_qmatmul_generic_blocks_vs_oracle(cuda_backend, K, N, bs):
    r = rng(K * N + bs)
    o = oracle.QuantizedWeight.from_slice(r.uniform(-1, 1, K * N).astype(np.float32), K, N, bs)
    x = r.standard_normal((3, K)).astype(np.float32)
    w = QuantizedWeight.upload(cuda_backend, o.data, o.scales, K, N, bs)
    assert rel_err(w.matmul(x, 3), oracle_out(o, x, 3)) < 2e-5
    w.free()


def test_qmatmul_edge_inputs(cuda_backend):
    K, N = 128, 64
    o = oracle.QuantizedWeight.from_slice(rng(5).uniform(-1, 1, K * N).astype(np.float32), K, N, 32)
    w = QuantizedWeight.upload(cuda_backend, o.data, o.scales, K, N, 32)
    assert np.array_equal(w.matmul(np.zeros((1, K), np.float32), 1), np.zeros((1, N), np.float32))
    e = np.zeros((1, K), np.float32)
    e[0, 77] = 1.0  # one-hot input reproduces row 77 of the dequantized weight (to the 2^-23 fixed-point grid of x*s)
    assert rel_err(w.matmul(e, 1).ravel(), o.dequantize_to()[77]) < 1e-6
    w.free()
    # all-zero block: scale 1.0, q = 0 (src/quant.zig:233-236)
    z = oracle.QuantizedWeight.from_slice(np.zeros(64 * 32, np.float32), 64, 32, 32)
    wz = QuantizedWeight.upload(cuda_backend, z.data, z.scales, 64, 32, 32)
    assert not wz.matmul(np.ones((1, 64), np.float32), 1).any()
    wz.free()


def test_split_k_is_deterministic_and_rearms(cuda_backend):
    K, N = 4096, 1024  # few column tiles -> split-K across CTAs
    raw = make_q8_0_raw(K, N, 3)
    w = QuantizedWeight.from_gguf_blocks(cuda_backend, raw, 8, K, N)
    x = rng(9).standard_normal((1, K)).astype(np.float32)
    outs = [w.matmul(x, 1) for _ in range(5)]
    for o in outs[1:]:
        assert np.array_equal(o.view(np.uint32), outs[0].view(np.uint32))
    w.free()


# ── full BASELINE sizes: oracle (threaded) + size-independent properties ─────
@pytest.mark.parametrize("K,N", [(4096, 4096), (4096, 14336)])
@pytest.mark.parametrize("kind", ["i8_f32", "q8_0", "q4_0"])
def test_full_size_microbench_shapes(cuda_backend, K, N, kind):
    r = rng(1)
    if kind == "i8_f32":
        o = oracle.QuantizedWeight(r.integers(-127, 128, K * N).astype(np.int8),
                                   r.uniform(1e-3, 1e-2, K * N // 32).astype(np.float32), K, N, 32)
        w = QuantizedWeight.upload(cuda_backend, o.data, o.scales, K, N, 32)
        assert w.format == abi.QFMT_I8_F32
    else:
        t = 8 if kind == "q8_0" else 2
        raw = make_q8_0_raw(K, N, 2) if t == 8 else make_q4_0_raw(K, N, 2)
        o = oracle.QuantizedWeight.from_gguf(raw, t, K, N)
        w = QuantizedWeight.from_gguf_blocks(cuda_backend, raw, t, K, N)
    x = r.standard_normal((1, K)).astype(np.float32)
    y = r.standard_normal((1, K)).astype(np.float32)
    gx, gy = w.matmul(x, 1), w.matmul(y, 1)
    want = o.matmul(x, 1, threads=8)
    assert rel_err(gx, want) < 2e-5
    # linearity: W^T(2x - y) == 2 W^T x - W^T y up to fp32 rounding
    gz = w.matmul(2 * x - y, 1)
    assert rel_err(gz, 2 * gx - gy) < 1e-4
    # batched rows equal the single-row results
    g2 = w.matmul(np.concatenate([x, y]), 2)
    assert rel_err(g2[0], gx[0]) < 1e-5 and rel_err(g2[1], gy[0]) < 1e-5
    # checksum of the dequantized weight (bit-exactness at full size)
    assert np.array_equal(w.dequantize_to().view(np.uint32), o.dequantize_to().view(np.uint32))
    w.free()


# ── DeviceOp.qmatmul contract inside a compiled program ───────────────────────
def test_program_qmatmul_offsets_strides_and_untouched_cells(cuda_backend):
    K, N, M = 96, 64, 3
    r = rng(4)
    o = oracle.QuantizedWeight.from_slice(r.uniform(-1, 1, K * N).astype(np.float32), K, N, 32)
    in_rs, out_rs, in_off, out_off = K + 5, N + 3, 7, 2
    inp = r.standard_normal(in_off + M * in_rs).astype(np.float32)
    dst0 = np.full(out_off + M * out_rs + 4, -7, np.float32)
    qw = QuantizedWeightUpload(o.data, o.scales, K, N, 32)
    prog = DeviceProgram([DeviceOp.qmatmul(1, 0, 0, M, N, K, in_off, in_rs, out_off, out_rs)],
                         [inp.size, dst0.size], [ProgramIO(0, inp), ProgramIO(1, dst0)], [qw])
    want = np.zeros_like(dst0)
    oracle.run_program(prog, [], [ProgramIO(1, want)])
    for graph in (True, False):
        cuda_backend.set_graph_mode(graph)
        h = cuda_backend.compile_program(prog)
        assert h is not None
        got = np.zeros_like(dst0)
        cuda_backend.execute_program(h, [], [ProgramIO(1, got)])
        cuda_backend.free_program(h)
        touched = want != -7
        assert np.array_equal(got[~touched], want[~touched])  # cells outside dst rows stay -7
        assert rel_err(got[touched], want[touched]) < 2e-5
    cuda_backend.set_graph_mode(True)


def test_program_rejects_bad_descriptors(cuda_backend):  # reference src/backend.zig:402-430
    o = oracle.QuantizedWeight.from_slice(np.ones(64 * 32, np.float32), 64, 32, 32)
    qw = QuantizedWeightUpload(o.data, o.scales, 64, 32, 32)
    bad_k = DeviceProgram([DeviceOp.qmatmul(1, 0, 0, 1, 32, 63)], [64, 32], [], [qw])
    assert not cuda_backend.supports_program(bad_k) and cuda_backend.compile_program(bad_k) is None
    bad_idx = DeviceProgram([DeviceOp.qmatmul(1, 0, 1, 1, 32, 64)], [64, 32], [], [qw])
    assert cuda_backend.compile_program(bad_idx) is None
    bad_buf = DeviceProgram([DeviceOp.qmatmul(5, 0, 0, 1, 32, 64)], [64, 32], [], [qw])
    assert cuda_backend.compile_program(bad_buf) is None


# ── prefill path: M > 8 on the tcgen05 tensor cores (3xBF16 operand terms, fp32 accumulate) ──────
@pytest.mark.parametrize("K,N", [(64, 64), (100, 160), (576, 1536), (1536, 576), (2048, 2048)])
@pytest.mark.parametrize("M", [9, 64, 128, 200, 257])
@pytest.mark.parametrize("kind", ["i8_f32", "q8_0", "q4_0"])
def test_qmatmul_prefill_tensor_core_path_vs_oracle(cuda_backend, K, N, M, kind):
    if K * N * M > 600e6:
        pytest.skip("oracle time")
    r = rng(K + 3 * N + M)
    if kind == "i8_f32":
        o = oracle.QuantizedWeight.from_slice(r.uniform(-1, 1, K * N).astype(np.float32), K, N, 32)
        w = QuantizedWeight.upload(cuda_backend, o.data, o.scales, K, N, 32)
    else:
        t = 8 if kind == "q8_0" else 2
        if K * N % 32:
            pytest.skip("GGUF blocks need K*N % 32 == 0")
        raw = make_q8_0_raw(K, N, 5 + K) if t == 8 else make_q4_0_raw(K, N, 5 + K)
        o = oracle.QuantizedWeight.from_gguf(raw, t, K, N)
        w = QuantizedWeight.from_gguf_blocks(cuda_backend, raw, t, K, N)
    x = r.standard_normal((M, K)).astype(np.float32)
    got = w.matmul(x, M)
    want = o.matmul(x, M, threads=8)
    e = rel_err(got, want)
    assert e < REL_TOL, e          # north-star tolerance
    assert e < 5e-5, e             # 3xBF16 (hi/lo split of both operands, 16 significant bits each; fp32 accumulation in TMEM)
    w.free()


@pytest.mark.parametrize("M,out_pad,out_off", [(40, 3, 2), (300, 3, 2), (300, 4, 4), (130, 0, 0)])
def test_program_prefill_qmatmul_strides_and_untouched_cells(cuda_backend, M, out_pad, out_off):
    """DeviceOp.qmatmul with M > 8 (tensor-core path): input / dst offsets and row strides (reference.zig:499-566);
    unaligned destinations take the epilogue's scalar row stores, 16-byte aligned ones the 128-bit stores; cells
    outside the dst rows keep their previous contents in both."""
    K, N = 96, 160
    r = rng(40 + M)
    o = oracle.QuantizedWeight.from_slice(r.uniform(-1, 1, K * N).astype(np.float32), K, N, 32)
    in_rs, out_rs, in_off = K + 5, N + out_pad, 7
    inp = r.standard_normal(in_off + M * in_rs).astype(np.float32)
    dst0 = np.full(out_off + M * out_rs + 4, -7, np.float32)
    qw = QuantizedWeightUpload(o.data, o.scales, K, N, 32)
    prog = DeviceProgram([DeviceOp.qmatmul(1, 0, 0, M, N, K, in_off, in_rs, out_off, out_rs)],
                         [inp.size, dst0.size], [ProgramIO(0, inp), ProgramIO(1, dst0)], [qw])
    want = np.zeros_like(dst0)
    oracle.run_program(prog, [], [ProgramIO(1, want)])
    h = cuda_backend.compile_program(prog)
    assert h is not None
    got = np.zeros_like(dst0)
    cuda_backend.execute_program(h, [], [ProgramIO(1, got)])
    cuda_backend.free_program(h)
    touched = want != -7
    assert np.array_equal(got[~touched], want[~touched])
    assert rel_err(got[touched], want[touched]) < 5e-5


def test_prefill_rows_match_decode_rows(cuda_backend):
    """Row i of an M = 40 tensor-core matmul equals the exact M = 1 matvec of that row within the 3xBF16 envelope."""
    K, N = 576, 1536
    raw = make_q8_0_raw(K, N, 77)
    w = QuantizedWeight.from_gguf_blocks(cuda_backend, raw, 8, K, N)
    x = rng(3).standard_normal((40, K)).astype(np.float32)
    big = w.matmul(x, 40)
    for i in (0, 7, 39):
        one = w.matmul(x[i:i + 1], 1)
        assert rel_err(big[i], one[0]) < 5e-5
    w.free()


# ── QuantizedWeight.fromSlice / matmulBias on device ────────────────
@pytest.mark.parametrize("K,N,bs", [(64, 64, 32), (100, 37, 32), (576, 1536, 32), (128, 96, 64), (33, 7, 16), (4096, 512, 128)])
def test_from_slice_on_device_is_bit_identical_to_reference(cuda_backend, K, N, bs):
    r = rng(K * 7 + N + bs)
    w = r.uniform(-2, 2, K * N).astype(np.float32)
    w[:bs] = 0.0                      # an all-zero block: scale 1.0, q 0 (src/quant.zig:236)
    w[bs + 3] = 1e-30                 # a tiny block maximum
    o = oracle.QuantizedWeight.from_slice(w, K, N, bs)
    qw, data, scales = QuantizedWeight.from_slice(cuda_backend, w, K, N, bs, return_flat=True)
    assert np.array_equal(data, o.data)
    assert np.array_equal(scales.view(np.uint32), o.scales.view(np.uint32))
    assert np.array_equal(qw.dequantize_to().view(np.uint32), o.dequantize_to().view(np.uint32))
    x = r.standard_normal((1, K)).astype(np.float32)
    assert rel_err(qw.matmul(x, 1), oracle_out(o, x, 1)) < 2e-5
    qw.free()


@pytest.mark.parametrize("M", [1, 3, 20])
def test_matmul_bias_matches_reference(cuda_backend, M):   # reference test: src/quant.zig:1211-1247 (tolerance 0.1 vs float)
    K, N = 96, 160
    r = rng(M)
    o = oracle.QuantizedWeight.from_slice(r.uniform(-1, 1, K * N).astype(np.float32), K, N, 32)
    w = QuantizedWeight.upload(cuda_backend, o.data, o.scales, K, N, 32)
    x = r.standard_normal((M, K)).astype(np.float32)
    b = r.standard_normal(N).astype(np.float32)
    assert rel_err(w.matmul_bias(x, b, M), o.matmul_bias(x, b, M)) < 5e-5
    w.free()


# ── gate | up matvec pair + activation epilogue in ONE launch (csrc/qgemv.cu qgemv_pair_kernel) ──────────────────
@pytest.mark.parametrize("K,N", [(128, 192), (576, 1536), (2048, 8192), (8192, 1024), (4096, 64)],
                         ids=["tiny", "smollm-135m", "smollm-1.7b", "long-k-split", "one-group-many-splits"])
@pytest.mark.parametrize("kind", ["i8_f32", "q8_0", "q4_0"])
@pytest.mark.parametrize("chain", ["silu", "bias-relu"])
def test_gate_up_pair_epilogue_matches_oracle_and_unpaired_path(cuda_backend, K, N, kind, chain):
    """The MLP head of a single-token program (reference src/nn.zig:38-44 lowered by src/device_inference.zig): two matvecs from
    one input, an elementwise chain on the first, times the second.  The pair launch must leave ALL FOUR buffers (gate, up,
    chain output, product) as the oracle executor does (1e-5 relative), and bit-identical to the unpaired schedule."""
    from zgml_b200 import CudaBackend
    r = rng(K * 3 + N)
    if kind == "i8_f32":
        oa = oracle.QuantizedWeight.from_slice(r.uniform(-1, 1, K * N).astype(np.float32), K, N, 32)
        ob = oracle.QuantizedWeight.from_slice(r.uniform(-1, 1, K * N).astype(np.float32), K, N, 32)
    else:
        t, mk = (8, make_q8_0_raw) if kind == "q8_0" else (2, make_q4_0_raw)
        oa, ob = oracle.QuantizedWeight.from_gguf(mk(K, N, 1), t, K, N), oracle.QuantizedWeight.from_gguf(mk(K, N, 2), t, K, N)
    x = r.standard_normal(K).astype(np.float32)
    bias = r.standard_normal(N).astype(np.float32)
    ones = np.ones(N, np.float32)
    # buffers: 0 x, 1 gate, 2 up, 3 chain output, 4 product, 5 ones, 6 bias
    if chain == "silu":      # gate * recip(exp(-gate) + 1)
        steps = [("neg", False, 0, 0), ("exp", False, 0, 0), ("add", False, 5, 0), ("recip", False, 0, 0), ("mul", True, 1, 0)]
    else:                    # relu(gate + bias) * up-in-chain, secondaries from an external buffer and from the second matvec
        steps = [("add", False, 6, 0), ("relu", False, 0, 0), ("mul", False, 2, 0)]
    ops = [DeviceOp.qmatmul(1, 0, 0, 1, N, K), DeviceOp.qmatmul(2, 0, 1, 1, N, K),
           DeviceOp.fused_elementwise(steps, N, 3, 1), DeviceOp.elementwise("mul", 4, 3, 2, N)]
    prog = DeviceProgram(ops, [K, N, N, N, N, N, N], [ProgramIO(0, x), ProgramIO(5, ones), ProgramIO(6, bias)],
                         [QuantizedWeightUpload(oa.data, oa.scales, K, N, 32), QuantizedWeightUpload(ob.data, ob.scales, K, N, 32)])
    want = [np.zeros(N, np.float32) for _ in range(4)]
    oracle.run_program(prog, [], [ProgramIO(1 + i, want[i]) for i in range(4)])

    def run(be):
        h = be.compile_program(prog)
        assert h is not None
        got = [np.zeros(N, np.float32) for _ in range(4)]
        for _ in range(2):   # twice: the split-K arrival counters must come back to zero
            be.execute_program(h, [], [ProgramIO(1 + i, got[i]) for i in range(4)])
        n_kernels = be.program_stats(h)["kernels"]   # counted when the step's graph is captured
        be.free_program(h)
        return got, n_kernels

    got, n_kernels = run(cuda_backend)
    assert n_kernels == 1
    for g, w_ in zip(got, want):
        assert rel_err(g, w_) < 1e-5
    import os
    os.environ["ZG_CUDA_GEMV_PAIR"] = "0"
    try:
        plain = CudaBackend(0)
    finally:
        del os.environ["ZG_CUDA_GEMV_PAIR"]
    try:
        got0, n0 = run(plain)
    finally:
        plain.close()
    assert n0 == 2
    for g, g0 in zip(got, got0):
        assert np.array_equal(g.view(np.uint32), g0.view(np.uint32))
