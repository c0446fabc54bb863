"""The streamed single-row matvec (csrc/qgemv_stream.cu: one column group per warp, evenly sliced work, split pieces met by
the last-arriving warp) against the oracle.  The default context only hands it launches with enough work (70B-size
matvecs, the microbench batches); here a second context forces EVERY M == 1 launch onto it (ZG_GEMV_STREAM=2) with 3 chunks
per CTA, so small shapes exercise ragged blocks of column groups, column groups cut into many pieces, several staging
segments per group and batched launches.  Bars as in test_gpu_qmatmul.py: 1e-3 of the output scale (north_star), 2e-5 observed."""
import os

import numpy as np
import pytest

from oracle import oracle
from test_gpu_qmatmul import make_q4_0_raw, make_q8_0_raw, oracle_out, rel_err, rng
from zgml_b200 import DeviceOp, DeviceProgram, ProgramIO, QuantizedWeight, QuantizedWeightUpload

pytestmark = pytest.mark.gpu


def _forced_backend(chunks):
    from zgml_b200 import CudaBackend
    os.environ["ZG_GEMV_STREAM"] = "2"
    os.environ["ZG_GEMV_STREAM_CHUNKS"] = str(chunks)
    try:
        return CudaBackend(0)
    finally:
        del os.environ["ZG_GEMV_STREAM"], os.environ["ZG_GEMV_STREAM_CHUNKS"]


@pytest.fixture(scope="module", params=[3, 32], ids=["3-chunks-per-cta", "32-chunks-per-cta"])
def stream_backend(request):
    be = _forced_backend(request.param)
    yield be
    be.close()


def _weights(kind, K, N, seed):
    if kind == "i8_f32":
        o = oracle.QuantizedWeight.from_slice(rng(seed).uniform(-1, 1, K * N).astype(np.float32), K, N, 32)
        return o, lambda be: QuantizedWeight.upload(be, o.data, o.scales, K, N, 32)
    t, mk = (8, make_q8_0_raw) if kind == "q8_0" else (2, make_q4_0_raw)
    raw = mk(K, N, seed)
    return oracle.QuantizedWeight.from_gguf(raw, t, K, N), lambda be: QuantizedWeight.from_gguf_blocks(be, raw, t, K, N)


# ragged blocks of column groups (N / 32 not a multiple of 8), K not a multiple of the chunk, one chunk, several staging
# segments per column group (K > 4096), long K with few column groups, the BASELINE microbench shape
SHAPES = [(32, 32), (64, 96), (100, 160), (129, 32), (576, 1536), (1536, 576), (2048, 2048), (8192, 320), (12288, 64), (28672, 256), (4096, 4096)]


@pytest.mark.parametrize("K,N", SHAPES)
@pytest.mark.parametrize("kind", ["i8_f32", "q8_0", "q4_0"])
def test_streamed_matvec_vs_oracle(stream_backend, K, N, kind):
    o, up = _weights(kind, K, N, 3 + K + N)
    w = up(stream_backend)
    x = rng(K).standard_normal((1, K)).astype(np.float32)
    want = oracle_out(o, x, 1)
    outs = [w.matmul(x, 1) for _ in range(3)]   # repeated: the arrival counters re-arm, the piece order is fixed
    w.free()
    assert rel_err(outs[0], want) < 2e-5
    for g in outs[1:]:
        assert np.array_equal(g.view(np.uint32), outs[0].view(np.uint32))


@pytest.mark.parametrize("kind", ["q8_0", "q4_0"])
def test_streamed_matvec_edge_activations(stream_backend, kind):
    K, N = 4224, 288   # 33 chunks of q4_0 per column group, 9 column groups
    o, up = _weights(kind, K, N, 7)
    w = up(stream_backend)
    assert not w.matmul(np.zeros((1, K), np.float32), 1).any()
    e = np.zeros((1, K), np.float32)
    e[0, 4100] = 1.0      # one-hot in the second staging segment: that row of the dequantized weight
    assert rel_err(w.matmul(e, 1).ravel(), o.dequantize_to()[4100]) < 1e-6
    x = rng(1).standard_normal((1, K)).astype(np.float32)
    x[0, 17] = np.nan     # non-finite activations poison every output, like the reference's float accumulation
    assert np.isnan(w.matmul(x, 1)).all()
    x[0, 17] = np.inf
    assert not np.isfinite(w.matmul(x, 1)).any()
    x = rng(2).standard_normal((1, K)).astype(np.float32)
    x[0, :4096] *= 1e-20  # a segment of tiny activations next to a normal one: each segment has its own fixed-point scale
    assert rel_err(w.matmul(x, 1), oracle_out(o, x, 1)) < 2e-5
    x[0, 5] = 1e4         # one outlier: the error bound follows max|x| of its segment
    want = oracle_out(o, x, 1)
    assert rel_err(w.matmul(x, 1), want) < 1e-3
    w.free()


@pytest.mark.parametrize("kind", ["i8_f32", "q4_0"])
@pytest.mark.parametrize("count", [2, 3, 8])
def test_streamed_batch_in_a_program_equals_single_launches(stream_backend, kind, count):
    """`count` same-shape matvecs of one dependency level share ONE streamed launch (one linear work space across the
    matvecs).  Where a column group is cut depends on the launch's slicing, so the bar is the oracle (2e-5) and run-to-run
    bit equality, not bit equality with a single launch."""
    K, N = 1024, 416   # 13 column groups: ragged second block
    os_, ups = zip(*[_weights(kind, K, N, 20 + i) for i in range(count)])
    xs = [rng(40 + i).standard_normal(K).astype(np.float32) for i in range(count)]
    qws = [QuantizedWeightUpload(o.data, o.scales, K, N, 32) for o in os_]
    ops = [DeviceOp.qmatmul(count + i, i, i, 1, N, K) for i in range(count)]
    prog = DeviceProgram(ops, [K] * count + [N] * count, [ProgramIO(i, xs[i]) for i in range(count)], qws)
    h = stream_backend.compile_program(prog)
    assert h is not None
    runs = []
    for _ in range(2):
        got = [np.zeros(N, np.float32) for _ in range(count)]
        stream_backend.execute_program(h, [], [ProgramIO(count + i, got[i]) for i in range(count)])
        runs.append(got)
    assert stream_backend.program_stats(h)["kernels"] == 1
    stream_backend.free_program(h)
    for i in range(count):
        assert rel_err(runs[0][i], oracle_out(os_[i], xs[i][None], 1).ravel()) < 2e-5
        assert np.array_equal(runs[0][i].view(np.uint32), runs[1][i].view(np.uint32))


def test_default_context_streams_only_large_launches(cuda_backend):
    """The default plan (csrc/qgemv_stream.cu zg_qgemv_stream_plan): small decode matvecs keep the k-split kernel, the Llama-3-70B
    gate | up shape streams — both within the same tolerance of the oracle."""
    for K, N in [(2048, 2048), (8192, 28672)]:
        raw = make_q4_0_raw(K, N, 5)
        o = oracle.QuantizedWeight.from_gguf(raw, 2, K, N)
        w = QuantizedWeight.from_gguf_blocks(cuda_backend, raw, 2, K, N)
        x = rng(K).standard_normal((1, K)).astype(np.float32)
        assert rel_err(w.matmul(x, 1), o.matmul(x, 1, threads=8)) < 2e-5
        w.free()
