"""Backend conformance on the GPU: the reference's acceptance test for a new backend
(reference src/backend/conformance.zig:348-372 — "A CUDA backend adds one more test
here"): every core program runs on the CUDA backend through the C-ABI and is compared
with the reference executor (the oracle) at the reference's 1e-5 tolerance, plus
randomised versions of the non-quantized DeviceOps at decode/prefill-like sizes."""
import numpy as np
import pytest

from conformance_programs import core_cases
from oracle import oracle
from zgml_b200 import DeviceOp, DeviceProgram, ProgramIO

pytestmark = pytest.mark.gpu


def run_both(be, program, out_idx, out_len, inputs=()):
    want = np.zeros(out_len, np.float32)
    oracle.run_program(program, list(inputs), [ProgramIO(out_idx, want)])
    assert be.supports_program(program)
    h = be.compile_program(program)
    assert h is not None
    got = np.zeros(out_len, np.float32)
    be.execute_program(h, list(inputs), [ProgramIO(out_idx, got)])
    be.free_program(h)
    return got, want


@pytest.mark.parametrize("graph", [True, False], ids=["graph", "eager"])
@pytest.mark.parametrize("case", core_cases(), ids=lambda c: c[0])
def test_cuda_backend_conforms_to_reference_core_ops(cuda_backend, case, graph):
    name, program, out_idx, out_len, closed_form = case
    cuda_backend.set_graph_mode(graph)
    got, want = run_both(cuda_backend, program, out_idx, out_len)
    cuda_backend.set_graph_mode(True)
    np.testing.assert_allclose(got, want, atol=1e-5, rtol=0)       # conformance.zig:350
    np.testing.assert_allclose(got, closed_form, atol=1e-5, rtol=0)


def test_capabilities_match_reference_cpu_profile(cuda_backend):  # reference src/backend.zig:60-70
    c = cuda_backend.capabilities
    assert c.compiled_programs and c.qmatmul and c.fused_elementwise and c.dynamic_program_refresh
    assert c.prefill_attention and c.decode_attention and c.attention_supported and c.attention_max_d_head == 512
    assert not c.host_visible_program_memory
    assert cuda_backend.dense_matmul_f32() is False  # declines host-pointer GEMMs (device_inference.zig:750-752)


def r32(seed, n, lo=-1.0, hi=1.0):
    return np.random.default_rng(seed).uniform(lo, hi, n).astype(np.float32)


def test_random_elementwise_and_fused_chain(cuda_backend):
    n = 5000
    a, b = r32(1, n, 0.1, 2.0), r32(2, n)
    ops = [DeviceOp.elementwise(op, 2, 0, 1, n, dst_offset=i * n) for i, op in
           enumerate(["add", "mul", "neg", "abs", "relu", "sqrt", "recip", "exp", "log"])]
    prog = DeviceProgram(ops, [n, n, 9 * n], [ProgramIO(0, a), ProgramIO(1, b)])
    got, want = run_both(cuda_backend, prog, 2, 9 * n)
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-6)
    # SiLU chain as zgml lowers it (src/nn.zig:38-44): x * recip(exp(-x) + 1), `ones` as secondary
    x, ones = r32(3, n, -6, 6), np.ones(n, np.float32)
    steps = [("neg", False, 0, 0), ("exp", False, 0, 0), ("add", False, 1, 0), ("recip", False, 0, 0), ("mul", True, 0, 0)]
    prog = DeviceProgram([DeviceOp.fused_elementwise(steps, n, 2, 0)], [n, n, n], [ProgramIO(0, x), ProgramIO(1, ones)])
    got, want = run_both(cuda_backend, prog, 2, n)
    np.testing.assert_allclose(got, want, rtol=3e-6, atol=1e-6)
    np.testing.assert_allclose(got, x / (1 + np.exp(-x)), rtol=1e-5, atol=1e-6)


def test_gelu_matches_reference_within_1e5(cuda_backend):
    n = 1001
    x = r32(4, n, -4, 4)
    prog = DeviceProgram([DeviceOp.elementwise("gelu", 1, 0, 0, n)], [n, n], [ProgramIO(0, x)])
    got, want = run_both(cuda_backend, prog, 1, n)
    np.testing.assert_allclose(got, want, atol=1e-5)


@pytest.mark.parametrize("rows,cols", [(1, 576), (3, 2048), (7, 100), (2, 8192)])
def test_random_norms_softmax_reduce(cuda_backend, rows, cols):
    x = r32(rows + cols, rows * cols, -3, 3)
    ops = [DeviceOp.rmsnorm(1, 0, rows, cols, 1e-5), DeviceOp.layernorm(1, 0, rows, cols, 1e-5, dst_offset=rows * cols),
           DeviceOp.softmax(1, 0, rows, cols, dst_offset=2 * rows * cols),
           DeviceOp.reduce("sum", 1, 0, rows, cols, dst_offset=3 * rows * cols),
           DeviceOp.reduce("max", 1, 0, rows, cols, dst_offset=3 * rows * cols + rows)]
    n_out = 3 * rows * cols + 2 * rows
    prog = DeviceProgram(ops, [rows * cols, n_out], [ProgramIO(0, x)])
    got, want = run_both(cuda_backend, prog, 1, n_out)
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5)


def test_random_repeat_modes(cuda_backend):
    # rmsnorm gamma broadcast as zgml lowers it: src [d,1] -> dst [d, T]
    d, T = 96, 5
    g = r32(5, d)
    prog = DeviceProgram([DeviceOp.repeat(1, 0, d * T, (d, 1, 1, 1), (d, T, 1, 1), (1, d, d, d), (1, d, d * T, d * T))],
                         [d, d * T], [ProgramIO(0, g)])
    got, want = run_both(cuda_backend, prog, 1, d * T)
    assert np.array_equal(got, want) and np.array_equal(got, np.tile(g, T))
    # scalar fill, plain copy and general strided broadcast (every other element of [1,2T] -> [d,T])
    s = r32(6, 2 * T)
    ops = [DeviceOp.repeat(1, 0, 7, (1, 1, 1, 1), (7, 1, 1, 1), (1, 1, 1, 1), (1, 7, 7, 7), src_offset=2),
           DeviceOp.repeat(1, 0, T, (T, 1, 1, 1), (T, 1, 1, 1), (1, T, T, T), (1, T, T, T), dst_offset=7),
           DeviceOp.repeat(1, 0, d * T, (1, T, 1, 1), (d, T, 1, 1), (1, 2, 2 * T, 2 * T), (1, d, d * T, d * T), dst_offset=7 + T)]
    n_out = 7 + T + d * T
    prog = DeviceProgram(ops, [2 * T, n_out], [ProgramIO(0, s)])
    got, want = run_both(cuda_backend, prog, 1, n_out)
    assert np.array_equal(got, want)
    assert np.array_equal(got[7 + T:], np.repeat(s[::2], d))


def test_random_rope_and_slice_assign(cuda_backend):
    hd, T, heads = 32, 4, 3
    d = 2 * hd
    src = r32(7, heads * d * T)                      # [heads*d, T] column-major like zgml: row stride 1, col stride heads*d
    cs = np.concatenate([np.cos(r32(8, hd * T, 0, 6)).reshape(T, hd), np.sin(r32(8, hd * T, 0, 6)).reshape(T, hd)], 1).ravel()
    ops = [DeviceOp.rope(2, 0, 1, hd, T, h * d, 0, h * d * T, 1, heads * d, d) for h in range(heads)]
    prog = DeviceProgram(ops, [src.size, cs.size, heads * d * T], [ProgramIO(0, src), ProgramIO(1, cs.astype(np.float32))])
    got, want = run_both(cuda_backend, prog, 2, heads * d * T)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))  # separate roundings, like the reference
    # slice_assign with strides (KV cache column store)
    rows, cols, cap = d, T, 16
    prog = DeviceProgram([DeviceOp.slice_assign(1, 0, rows, cols, 0, 5 * rows, 1, rows, 3, 1, rows, rows)],
                         [3 + rows * cols, rows * cap], [ProgramIO(0, src[:3 + rows * cols])])
    got, want = run_both(cuda_backend, prog, 1, rows * cap)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("d_head,seq_q,seq_kv", [(64, 1, 1), (64, 1, 37), (64, 1, 513), (128, 5, 64), (40, 3, 9), (512, 2, 17),
                                                  (128, 1, 1000), (64, 4, 700), (32, 1, 2048), (256, 2, 300)])
def test_random_attention(cuda_backend, d_head, seq_q, seq_kv):
    q = r32(1, d_head * seq_q)
    k = r32(2, d_head * seq_kv)
    v = r32(3, d_head * seq_kv)
    mask = np.zeros((seq_q, seq_kv), np.float32)
    for i in range(seq_q):  # causal-style mask with -inf tail, zgml layout mask[s*mask_rs + qi*mask_cs]
        mask[i, max(1, seq_kv - seq_q + i + 1):] = -np.inf
    prog = DeviceProgram(
        [DeviceOp.attention(4, 0, 1, 2, 3, True, d_head, seq_q, seq_kv, 1.0 / np.sqrt(d_head), 0, 0, 0, 0, 0,
                            1, d_head, 1, d_head, 1, d_head, 1, seq_kv, 1, d_head)],
        [q.size, k.size, v.size, mask.size, d_head * seq_q],
        [ProgramIO(0, q), ProgramIO(1, k), ProgramIO(2, v), ProgramIO(3, mask.ravel())])
    got, want = run_both(cuda_backend, prog, 4, d_head * seq_q)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-6)


def test_attention_fully_masked_row_is_zero(cuda_backend):  # reference.zig:664-667: l == 0 -> zeros
    d_head, seq_kv = 16, 4
    mask = np.full(seq_kv, -np.inf, np.float32)
    prog = DeviceProgram(
        [DeviceOp.attention(4, 0, 1, 2, 3, True, d_head, 1, seq_kv, 0.25, 0, 0, 0, 0, 0, 1, d_head, 1, d_head, 1, d_head, 1, seq_kv, 1, d_head)],
        [d_head, d_head * seq_kv, d_head * seq_kv, seq_kv, d_head],
        [ProgramIO(0, r32(1, d_head)), ProgramIO(1, r32(2, d_head * seq_kv)), ProgramIO(2, r32(3, d_head * seq_kv)), ProgramIO(3, mask)])
    got, want = run_both(cuda_backend, prog, 4, d_head)
    assert not got.any() and not want.any()


@pytest.mark.parametrize("M,N,K,kmajor", [(1, 300, 64, True), (1, 4096, 576, True), (3, 50, 33, False), (2, 64, 20, True)])
def test_random_dense_matmul(cuda_backend, M, N, K, kmajor):
    a, b = r32(1, M * K), r32(2, K * N)
    if kmajor:   # tied LM head: B stored [N, K] (src/models/llama.zig:162-165)
        op = DeviceOp.matmul(2, 0, 1, M, N, K, K, 1, 1, K)
    else:
        op = DeviceOp.matmul(2, 0, 1, M, N, K, K, 1, N, 1)
    prog = DeviceProgram([op], [M * K, K * N, M * N], [ProgramIO(0, a), ProgramIO(1, b)])
    got, want = run_both(cuda_backend, prog, 2, M * N)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)


def test_refresh_patches_slice_offset_and_seq_kv(cuda_backend):
    """Per-step dynamic state (reference src/device_inference.zig:242-256): one compiled program,
    `refresh_program` with patched slice_assign.dst_offset / attention.seq_kv before every execute."""
    d, cap = 32, 8
    kv_ops = lambda pos, skv: [
        DeviceOp.slice_assign(1, 0, d, 1, 0, pos * d, 1, d, 0, 1, d, d),       # k_cache[:, pos] = x
        DeviceOp.slice_assign(2, 0, d, 1, 0, pos * d, 1, d, 0, 1, d, d),       # v_cache[:, pos] = x
        DeviceOp.attention(4, 0, 1, 2, 3, True, d, 1, skv, 0.3, 0, 0, 0, 0, 0, 1, d, 1, d, 1, d, 1, cap, 1, d)]
    mask = np.zeros(cap, np.float32)
    prog = DeviceProgram(kv_ops(0, 1), [d, d * cap, d * cap, cap, d], [ProgramIO(3, mask)])
    st = oracle.ProgramState(prog)
    for graph in (True, False):
        cuda_backend.set_graph_mode(graph)
        h = cuda_backend.compile_program(prog)
        st.close()
        st = oracle.ProgramState(prog)
        for pos in range(cap):
            ops = kv_ops(pos, pos + 1)
            x = r32(100 + pos, d)
            arr = (type(ops[0]) * len(ops))(*ops)
            want, got = np.zeros(d, np.float32), np.zeros(d, np.float32)
            st.execute(arr, len(ops), [ProgramIO(0, x)], [ProgramIO(4, want)])
            cuda_backend.refresh_program(h, ops)
            cuda_backend.execute_program(h, [ProgramIO(0, x)], [ProgramIO(4, got)])
            np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-6)
        cuda_backend.free_program(h)
    st.close()
    cuda_backend.set_graph_mode(True)


@pytest.mark.parametrize("graph", [True, False], ids=["graph", "eager"])
def test_split_kv_decode_attention_over_a_growing_context(cuda_backend, graph):
    """Three heads of one layer (batched launch, split over the kv range, head output also stored into the concatenated
    buffer by the absorbed slice_assign) against a 1024-row cache whose valid length is patched per step."""
    d, cap, heads = 64, 1024, 3
    bufs = [d * heads, d * cap * heads, d * cap * heads, cap] + [d] * heads + [d * heads]   # q | k | v | mask | out_h.. | concat
    concat = 4 + heads

    def ops_for(skv):
        ops = []
        for h in range(heads):
            ops.append(DeviceOp.attention(4 + h, 0, 1, 2, 3, True, d, 1, skv, 0.125, h * d, h * d * cap, h * d * cap, 0, 0,
                                          1, d, 1, d, 1, d, 1, cap, 1, d))
            ops.append(DeviceOp.slice_assign(concat, 4 + h, d, 1, 0, h * d, 1, d * heads, 0, 1, d, 0))
        return ops

    prog = DeviceProgram(ops_for(1), bufs, [ProgramIO(0, r32(1, d * heads)), ProgramIO(1, r32(2, d * cap * heads)),
                                            ProgramIO(2, r32(3, d * cap * heads)), ProgramIO(3, np.zeros(cap, np.float32))])
    cuda_backend.set_graph_mode(graph)
    h = cuda_backend.compile_program(prog)
    st = oracle.ProgramState(prog)
    for skv in (1, 127, 128, 129, 400, 1023, 1024, 5):
        ops = ops_for(skv)
        arr = (type(ops[0]) * len(ops))(*ops)
        want, got = np.zeros(d * heads, np.float32), np.zeros(d * heads, np.float32)
        want1, got1 = np.zeros(d, np.float32), np.zeros(d, np.float32)
        st.execute(arr, len(ops), [], [ProgramIO(concat, want), ProgramIO(5, want1)])
        cuda_backend.refresh_program(h, ops)
        cuda_backend.execute_program(h, [], [ProgramIO(concat, got), ProgramIO(5, got1)])
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-6)
        np.testing.assert_allclose(got1, want1, rtol=1e-4, atol=2e-6)
    cuda_backend.free_program(h)
    st.close()
    cuda_backend.set_graph_mode(True)


@pytest.mark.parametrize("rows,cols,with_add", [(1, 64, True), (3, 2048, True), (1, 4096, False), (1, 4100, True), (1, 8192, True), (1, 8192, False), (2, 4096, True)])
def test_norm_pattern_writes_every_buffer(cuda_backend, rows, cols, with_add):
    """[add,] rmsnorm, repeat(gamma over rows), mul — the lowering's norm block (llama_transformer.zig:118-125) is evaluated
    in one pass (chain macro up to 4096 columns, a 1024-thread kernel per row above); sum, bare, gamma_rep and the
    result must all equal op-by-op execution."""
    n = rows * cols
    a, b, gamma = r32(1, n), r32(2, n), r32(3, cols)
    ops = []
    if with_add:
        ops.append(DeviceOp.elementwise("add", 3, 0, 1, n))
    src = 3 if with_add else 0
    ops += [DeviceOp.rmsnorm(4, src, rows, cols, 1e-5),
            DeviceOp.repeat(5, 2, n, (cols, 1, 1, 1), (cols, rows, 1, 1), (1, cols, cols, cols), (1, cols, n, n)),
            DeviceOp.elementwise("mul", 6, 4, 5, n)]
    prog = DeviceProgram(ops, [n, n, cols, n, n, n, n], [ProgramIO(0, a), ProgramIO(1, b), ProgramIO(2, gamma)])
    h = cuda_backend.compile_program(prog)
    outs_g = [np.zeros(n, np.float32) for _ in range(4)]
    outs_w = [np.zeros(n, np.float32) for _ in range(4)]
    cuda_backend.execute_program(h, [], [ProgramIO(3 + i, outs_g[i]) for i in range(4)])
    cuda_backend.free_program(h)
    oracle.run_program(prog, [], [ProgramIO(3 + i, outs_w[i]) for i in range(4)])
    for g, w in zip(outs_g, outs_w):
        np.testing.assert_allclose(g, w, rtol=2e-5, atol=1e-6)


def test_runtime_profile_slot(cuda_backend):  # reference src/backend.zig:351, src/profile.zig:819-842
    name, program, out_idx, out_len, _ = core_cases()[1]
    h = cuda_backend.compile_program(program)
    assert cuda_backend.get_runtime_profile(h) is None  # disabled by default, like cpu.zig:136-138
    cuda_backend.set_profiling(True)
    out = np.zeros(out_len, np.float32)
    cuda_backend.execute_program(h, [], [ProgramIO(out_idx, out)])
    prof = cuda_backend.get_runtime_profile(h)
    assert prof is not None and prof.call_count == 1 and prof.backend_op_count == 1 and prof.fallback_op_count == 0
    cuda_backend.set_profiling(False)
    cuda_backend.free_program(h)
