"""The world-size independent synthetic model (zg_cuda_qweight_synth_gguf + host twin) and full-shape parity.

* device generator == host twin, slab by slab, bit for bit (dequantized weights of column / row slabs, Q8_0 and Q4_0);
* BASELINE.json config 4 shapes in pytest: M = 2048 rows through the tcgen05 path against the oracle on sampled rows
  of 4096 x {4096, 14336} (Q8_0 and Q4_0) — every output row depends on its own activation row only, so the sampled
  rows check the full-size launch;
* config 3 shape: SmolLM-1.7B-width Q4_0 greedy decode at context 512 (layers reduced, everything else at size)
  against the oracle executor: logits 1e-3, greedy tokens identical."""
import numpy as np
import pytest

from llama_reference import OracleBackend
from oracle import oracle
from zgml_b200 import QuantizedWeight
from zgml_b200.host import llama

pytestmark = pytest.mark.gpu


def rel(got, want):
    return float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64))) / (np.max(np.abs(want)) + 1e-30))


@pytest.mark.parametrize("kind,ggml", [("q8_0", 8), ("q4_0", 2)])
@pytest.mark.parametrize("slab", [(0, 96, 0, 160), (0, 96, 64, 128), (32, 64, 0, 160), (5, 37, 32, 160)],
                         ids=["whole", "column-slab", "row-slab", "ragged-rows"])
def test_device_generator_equals_host_twin(cuda_backend, kind, ggml, slab):
    K, N = 96, 160
    k0, k1, n0, n1 = slab
    tid = llama.tensor_id(3, "w_up")
    h = QuantizedWeight.synth_gguf(cuda_backend, 7, tid, ggml, K, N, k0, k1, n0, n1)
    raw = llama.synth_gguf_blocks(7, tid, kind, K, N, k0, k1, n0, n1)
    want = oracle.QuantizedWeight.from_gguf(raw, ggml, k1 - k0, n1 - n0).dequantize_to()
    got = h.dequantize_to()
    h.free()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # ... and the slab is a slice of the whole tensor
    whole = oracle.QuantizedWeight.from_gguf(llama.synth_gguf_blocks(7, tid, kind, K, N), ggml, K, N).dequantize_to()
    assert np.array_equal(want, whole[k0:k1, n0:n1])


@pytest.mark.parametrize("kind,ggml", [("q8_0", 8), ("q4_0", 2)])
@pytest.mark.parametrize("K,N", [(4096, 4096), (4096, 14336)])
def test_prefill_2048_rows_llama3_8b_linears_vs_oracle(cuda_backend, kind, ggml, K, N):
    """Config 4: 2048 tokens x a Llama-3-8B linear on the tensor cores; 24 sampled rows (first / last tile included) vs the oracle."""
    M = 2048
    tid = llama.tensor_id(0, "w_gate")
    w = QuantizedWeight.synth_gguf(cuda_backend, 1, tid, ggml, K, N, 0, K, 0, N)
    o = oracle.QuantizedWeight.from_gguf(llama.synth_gguf_blocks(1, tid, kind, K, N), ggml, K, N)
    r = np.random.default_rng(9)
    x = r.standard_normal((M, K)).astype(np.float32)
    got = w.matmul(x, M)
    w.free()
    rows = np.unique(np.concatenate([[0, 1, 127, 128, 255, 256, 1023, 1024, 2046, 2047], r.integers(0, M, 14)]))
    want = o.matmul(np.ascontiguousarray(x[rows]), len(rows), threads=4, native=True)
    assert np.isfinite(got).all()
    assert rel(got[rows], want) < 1e-3          # north_star budget; the kernel lands near 2e-5
    assert rel(got[rows], want) < 1e-4
    # per element: |delta| <= 1e-3 |want| + a rounding floor proportional to the row's typical magnitude
    floor = 1e-4 * np.sqrt(np.mean(want.astype(np.float64) ** 2, axis=1, keepdims=True))
    assert (np.abs(got[rows].astype(np.float64) - want) <= 1e-3 * np.abs(want) + floor).all()


def test_smollm_1p7b_width_q4_0_decode_at_context_512_matches_oracle(cuda_backend):
    """Config 3 at its real width (d_model 2048, d_ff 8192, 32 heads, vocab 49152 tied, max_seq 2048), 2 layers: greedy decode
    from position 512 (cache rows below hold zeros on both sides, as scripts/bench_decode.py --context does)."""
    cfg = llama.LlamaConfig(**{**llama.SMOLLM_1_7B.__dict__, "n_layers": 2})
    wd, handles = llama.synthetic_resident_shard(cuda_backend, cfg, "q4_0", seed=3)
    wh = llama.synthetic_model_host(cfg, "q4_0", seed=3)
    dev = llama.DeviceLlamaSession(cuda_backend, cfg, wd)
    ref = llama.DeviceLlamaSession(OracleBackend(native=True), cfg, wh)
    dev.pos = ref.pos = 512
    t_d = t_r = 1
    for _ in range(3):
        lg_d, lg_r = dev.step(t_d).copy(), ref.step(t_r).copy()
        assert rel(lg_d, lg_r) < 1e-3
        t_d, t_r = int(np.argmax(lg_d)), int(np.argmax(lg_r))
        assert t_d == t_r
    dev.close(); ref.close()
    for h in handles:
        h.free()


@pytest.mark.parametrize("fmt", ["i8_f32", "q8_0", "q4_0"])
@pytest.mark.parametrize("case", ["outliers", "scale-spread", "both"])
def test_matvec_adversarial_activations_and_scales(cuda_backend, fmt, case):
    """Heavy-tailed activations (a few 1e3 x outliers per row) and per-block scales spanning three decades inside every column
    group: the decode matvec quantises x * s to a 2^-23 grid relative to max|x| * max(s) of a warp's k-range
    (csrc/qgemv.cu), so its error is bounded by that product, not by the individual terms.  Checked per element:
    |delta| <= 1e-3 |want| + 2^-22 max|x| smax sum_k |q[k, n]|  (the analytic worst case: half a grid step per term), 1e-3
    relative on every output that is at least 5 % of the row maximum, and the usual 1e-3 on the row maximum."""
    K, N = 2048, 512
    r = np.random.default_rng({"outliers": 1, "scale-spread": 2, "both": 3}[case])
    qmax = 7 if fmt == "q4_0" else 127
    q = r.integers(-qmax - (1 if fmt == "q4_0" else 0), qmax + 1, (K, N), dtype=np.int8)
    if case == "outliers":
        s = r.uniform(1e-3, 1e-2, (K, N // 32))
    else:
        s = 10.0 ** r.uniform(-5, -2, (K, N // 32))
    s = s.astype(np.float16).astype(np.float32) if fmt != "i8_f32" else s.astype(np.float32)
    x = r.standard_normal((1, K)).astype(np.float32)
    if case != "scale-spread":
        x[0, r.integers(0, K, 6)] *= 1e3
    o = oracle.QuantizedWeight(q.ravel(), s.ravel(), K, N, 32)
    w = QuantizedWeight.upload(cuda_backend, q.ravel(), s.ravel(), K, N, 32)
    assert w.format == {"i8_f32": 1, "q8_0": 2, "q4_0": 3}[fmt]
    want = np.zeros(N, np.float32)
    o.qmatmul_op(x.ravel(), want, 1)
    got = w.matmul(x, 1).ravel()
    w.free()
    assert rel(got, want) < 1e-3
    smax = np.repeat(s.max(axis=0), 32)                                   # per column: its group's largest scale
    bound = 2.0 ** -22 * float(np.abs(x).max()) * smax * np.abs(q.astype(np.float64)).sum(axis=0)
    err = np.abs(got.astype(np.float64) - want)
    assert (err <= 1e-3 * np.abs(want) + bound).all(), float((err / (np.abs(want) + 1e-30)).max())
    big = np.abs(want) >= 0.05 * np.abs(want).max()
    assert (err[big] <= 1e-3 * np.abs(want[big])).all(), float((err[big] / np.abs(want[big])).max())
