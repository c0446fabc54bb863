"""GPU parity tests of the quantized KV cache path (QuantizedKVCache.storeColumn / attentionQuantized, reference
src/quant.zig:633-1091) through the C-ABI, against the CPU oracle on the same seeded inputs.

Bars: cache contents (int8 data and f32 scales) BIT-EXACT; attention outputs within 1e-5 absolute of the oracle on
O(1) values (scores are bit-identical; the softmax groups 32 x 8 positions where the reference groups 8, and expf
differs from libm in the last place) — far inside the reference's own 0.01 envelope (src/quant.zig:1410)."""
import numpy as np
import pytest

from oracle import oracle
from zgml_b200 import BackendError
from zgml_b200.backend import QuantizedKVCache, attention_quantized

pytestmark = pytest.mark.gpu
ATOL = 1e-5


def fill(n, seed):
    return ((np.random.default_rng(seed).random(n, dtype=np.float32) - 0.5) * 2.0).astype(np.float32)


def make_caches(be, d, n_cols, bs, seed, used=None):
    """Oracle and device caches filled column by column with the same data (`used` columns, default all)."""
    used = n_cols if used is None else used
    ok_, ov_ = oracle.QuantizedKVCache(d, n_cols, bs), oracle.QuantizedKVCache(d, n_cols, bs)
    gk, gv = QuantizedKVCache(be, d, n_cols, bs), QuantizedKVCache(be, d, n_cols, bs)
    kd, vd = fill(d * used, seed), fill(d * used, seed + 1)
    for c in range(used):
        ok_.store_column(c, kd[c * d:(c + 1) * d])
        ov_.store_column(c, vd[c * d:(c + 1) * d])
    gk.store_columns(0, kd.reshape(used, d))                # batched write (prefill), src/llama_inference.zig:343-347
    for c in range(used):
        gv.store_column(c, vd[c * d:(c + 1) * d])           # one column at a time (decode)
    return ok_, ov_, gk, gv


@pytest.mark.parametrize("d,bs,n_cols", [(64, 32, 4), (128, 32, 9), (64, 64, 3), (128, 16, 5), (512, 32, 2), (8, 4, 6)])
def test_store_column_bit_identical(cuda_backend, d, bs, n_cols):
    ok_, ov_, gk, gv = make_caches(cuda_backend, d, n_cols, bs, d + bs, used=n_cols - 1)   # last column stays zero
    for o, g in ((ok_, gk), (ov_, gv)):
        q, s = g.download()
        assert np.array_equal(q, o.q_data) and np.array_equal(s.view(np.uint32), o.scales.view(np.uint32))
    gk.store_column(1, np.zeros(d, np.float32))             # an all-zero column: scale 1, q 0 (src/quant.zig:333)
    ok_.store_column(1, np.zeros(d, np.float32))
    q, s = gk.download()
    assert np.array_equal(q, ok_.q_data) and np.array_equal(s.view(np.uint32), ok_.scales.view(np.uint32))
    assert np.max(np.abs(gk.dequant_column(0) - ok_.dequant_column(0))) == 0
    gk.clear()
    q, s = gk.download()
    assert not q.any() and not s.any()
    gk.free(); gv.free()


@pytest.mark.parametrize("int8_query", [True, False])
@pytest.mark.parametrize("d,bs,seq_kv", [(64, 32, 8), (64, 32, 21), (32, 32, 7), (128, 32, 512), (128, 16, 300), (64, 64, 33),
                                         (512, 32, 40), (8, 4, 5), (96, 12, 70)])
def test_attention_decode_matches_oracle(cuda_backend, d, bs, seq_kv, int8_query):
    ok_, ov_, gk, gv = make_caches(cuda_backend, d, seq_kv, bs, 3 * d + seq_kv)
    q = fill(d, 7)
    scale = np.float32(1.0 / np.sqrt(d))
    want = oracle.attention_quantized(q, 1, ok_, 0, ov_, 0, seq_kv, scale, use_sdot=int8_query)
    got = attention_quantized(cuda_backend, q, 1, gk, 0, gv, 0, seq_kv, scale, int8_query=int8_query)
    assert np.max(np.abs(got - want)) < ATOL
    gk.free(); gv.free()


@pytest.mark.parametrize("int8_query", [True, False])
def test_attention_causal_broadcast_mask_and_slabs(cuda_backend, int8_query):  # src/quant.zig:1413-1545
    d, slab, bs = 32, 40, 32
    ok_, ov_, gk, gv = make_caches(cuda_backend, d, 3 * slab, bs, 11)
    q = fill(d, 12)
    scale = np.float32(1.0 / np.sqrt(d))
    for pos in (0, 4, 31, 32, 39):
        mask = np.where(np.arange(slab) <= pos, 0.0, -np.inf).astype(np.float32)
        for start in (0, slab, 2 * slab):
            want = oracle.attention_quantized(q, 1, ok_, start, ov_, start, slab, scale, mask=mask, mask_row_stride=1,
                                              mask_col_stride=0, use_sdot=int8_query)
            got = attention_quantized(cuda_backend, q, 1, gk, start, gv, start, slab, scale, mask=mask, mask_row_stride=1,
                                      mask_col_stride=0, int8_query=int8_query)
            assert np.max(np.abs(got - want)) < ATOL
    gk.free(); gv.free()


def test_attention_prefill_columns_strides_and_fully_masked(cuda_backend):
    d, seq_kv, seq_q, bs = 64, 45, 4, 32
    ok_, ov_, gk, gv = make_caches(cuda_backend, d, seq_kv, bs, 21)
    q_cs, d_cs = d + 5, d + 3
    q = fill(q_cs * seq_q, 22)
    mask = np.zeros((seq_q, seq_kv), np.float32)            # column qi at qi * seq_kv, row stride 1
    mask[0, 5:] = -np.inf
    mask[1, :] = -np.inf                                    # fully masked query column -> zeros
    mask[2, 33:] = -np.inf
    mask[3, ::2] = -1.5                                     # finite additive bias
    scale = np.float32(0.2)
    for int8_query in (True, False):
        want = oracle.attention_quantized(q, seq_q, ok_, 0, ov_, 0, seq_kv, scale, mask=mask, mask_row_stride=1,
                                          mask_col_stride=seq_kv, use_sdot=int8_query, q_col_stride=q_cs, dst_col_stride=d_cs)
        got = attention_quantized(cuda_backend, q, seq_q, gk, 0, gv, 0, seq_kv, scale, mask=mask, mask_row_stride=1,
                                  mask_col_stride=seq_kv, int8_query=int8_query, q_col_stride=q_cs, dst_col_stride=d_cs)
        assert np.max(np.abs(got - want)) < ATOL
        assert not got[d_cs:d_cs + d].any()                 # the fully masked column
        assert not got[d:d_cs].any()                        # cells between dst columns untouched
    gk.free(); gv.free()


def test_quantized_kv_contract_errors(cuda_backend):
    with pytest.raises(BackendError):
        QuantizedKVCache(cuda_backend, 48, 4, 32)           # d_head % block_size != 0 (src/quant.zig:659)
    with pytest.raises(BackendError):
        QuantizedKVCache(cuda_backend, 1024, 4, 32)         # d_head > 512 (src/quant.zig:944-947)
    k, v = QuantizedKVCache(cuda_backend, 64, 8, 32), QuantizedKVCache(cuda_backend, 64, 8, 32)
    with pytest.raises(BackendError):
        k.store_columns(7, np.zeros((2, 64), np.float32))   # past the last column
    with pytest.raises(BackendError):
        attention_quantized(cuda_backend, np.zeros(64, np.float32), 1, k, 4, v, 4, 5, 1.0)   # kv range outside the cache
    k.free(); v.free()
    k, v = QuantizedKVCache(cuda_backend, 512, 2, 4), QuantizedKVCache(cuda_backend, 512, 2, 4)   # 128 blocks per column
    with pytest.raises(BackendError):
        attention_quantized(cuda_backend, np.zeros(512, np.float32), 1, k, 0, v, 0, 2, 1.0, int8_query=True)
    out = attention_quantized(cuda_backend, np.ones(512, np.float32), 1, k, 0, v, 0, 2, 1.0, int8_query=False)
    assert not out.any()                                    # empty (zero) cache: values are zero
    k.free(); v.free()
