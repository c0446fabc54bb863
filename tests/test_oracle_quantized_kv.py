"""Oracle pinning for the quantized KV cache path: the reference's own tests for
QuantizedKVCache / attentionQuantized (src/quant.zig:1259-1620) re-run against the C restatement.

They are property tests against an in-test float reference with stated tolerances; inputs come from
std.Random.DefaultPrng in the reference — any RNG re-runs them, here numpy U(-1, 1) like
fillRandF32 (src/quant.zig:1253-1257).  CPU only: the oracle is the checker the GPU tests compare against."""
import numpy as np
import pytest

from oracle import oracle


def fill(n, seed):  # fillRandF32: (rng.float - 0.5) * 2
    return ((np.random.default_rng(seed).random(n, dtype=np.float32) - 0.5) * 2.0).astype(np.float32)


def streaming_reference(q, k_ref, v_ref, d_head, seq_kv, scale, mask=None):
    """The single-column streaming softmax the reference's tests use (src/quant.zig:1372-1391), in f32."""
    out = np.zeros(d_head, np.float32)
    m_val, l = -np.inf, np.float32(0)
    for s in range(seq_kv):
        add = 0.0 if mask is None else mask[s]
        if not np.isfinite(add):
            continue
        dot = np.float32(0)
        for r in range(d_head):
            dot = np.float32(dot + q[r] * k_ref[s * d_head + r])
        score = np.float32(dot * scale + add)
        new_m = max(m_val, score)
        alpha = np.float32(0) if m_val == -np.inf else np.exp(np.float32(m_val - new_m))
        w = np.exp(np.float32(score - new_m))
        out = (out * alpha + w * v_ref[s * d_head:(s + 1) * d_head]).astype(np.float32)
        l = np.float32(l * alpha + w)
        m_val = new_m
    return out * (np.float32(1) / l if l > 0 else np.float32(0))


def filled_caches(d_head, seq_kv, bs, seed_k, seed_v):
    k, v = oracle.QuantizedKVCache(d_head, seq_kv, bs), oracle.QuantizedKVCache(d_head, seq_kv, bs)
    kd, vd = fill(d_head * seq_kv, seed_k), fill(d_head * seq_kv, seed_v)
    for c in range(seq_kv):
        k.store_column(c, kd[c * d_head:(c + 1) * d_head])
        v.store_column(c, vd[c * d_head:(c + 1) * d_head])
    k_ref = np.concatenate([k.dequant_column(c) for c in range(seq_kv)])
    v_ref = np.concatenate([v.dequant_column(c) for c in range(seq_kv)])
    return k, v, k_ref, v_ref


def test_store_then_dequant_roundtrip():  # src/quant.zig:1259-1276
    cache = oracle.QuantizedKVCache(64, 4, 32)
    src = fill(64, 1)
    cache.store_column(2, src)
    assert np.max(np.abs(cache.dequant_column(2) - src)) < 0.02
    assert not cache.q_data[:2 * 64].any() and not cache.q_data[3 * 64:].any()   # other columns untouched
    # storeColumn is quantizeInput on the column (src/quant.zig:694-700)
    q, s = oracle.quantize_input(src, 32)
    assert np.array_equal(cache.q_data[2 * 64:3 * 64], q) and np.array_equal(cache.scales[4:6], s)


def test_dots_match_float_reference():  # src/quant.zig:1278-1317 (dotF32 vs float, dotI8 vs dotF32: 0.05)
    d = 64
    k_col, q_vec = fill(d, 2), fill(d, 3)
    cache = oracle.QuantizedKVCache(d, 1, 32)
    cache.store_column(0, k_col)
    ref = float(np.dot(q_vec.astype(np.float64), k_col.astype(np.float64)))
    # dotF32 / dotI8 are private to the attention loop here: read the dot back through a two-position softmax whose
    # second key is zero (score 0) and whose values are 1 and 0 — out[0] = 1 / (1 + exp(-dot))
    two_k = oracle.QuantizedKVCache(d, 2, 32)
    two_k.store_column(0, k_col)
    two_k.store_column(1, np.zeros(d, np.float32))    # score 0
    two_v = oracle.QuantizedKVCache(d, 2, 32)
    two_v.store_column(0, np.ones(d, np.float32))
    two_v.store_column(1, np.zeros(d, np.float32))
    for use_sdot in (False, True):
        out = oracle.attention_quantized(q_vec, 1, two_k, 0, two_v, 0, 2, 1.0, use_sdot=use_sdot)
        p0 = float(out[0])
        dot = np.log(p0 / (1 - p0))
        assert abs(dot - ref) < 0.05


@pytest.mark.parametrize("use_sdot", [True, False])
def test_attention_decode_matches_reference(use_sdot):  # src/quant.zig:1339-1411
    d, seq_kv = 64, 8
    k, v, k_ref, v_ref = filled_caches(d, seq_kv, 32, 10, 11)
    q = fill(d, 12)
    scale = np.float32(1.0 / np.sqrt(d))
    out = oracle.attention_quantized(q, 1, k, 0, v, 0, seq_kv, scale, use_sdot=use_sdot)
    want = streaming_reference(q, k_ref, v_ref, d, seq_kv, scale)
    assert np.max(np.abs(out - want)) < (0.01 if use_sdot else 1e-5)   # SDOT path adds the Q-quantization error


@pytest.mark.parametrize("use_sdot", [True, False])
def test_attention_causal_mask_broadcast_column(use_sdot):  # src/quant.zig:1413-1485
    d, seq_kv, pos = 32, 8, 4
    k, v, k_ref, v_ref = filled_caches(d, seq_kv, 32, 20, 21)
    q = fill(d, 22)
    mask = np.where(np.arange(seq_kv) <= pos, 0.0, -np.inf).astype(np.float32)
    scale = np.float32(1.0 / np.sqrt(d))
    out = oracle.attention_quantized(q, 1, k, 0, v, 0, seq_kv, scale, mask=mask, mask_row_stride=1, mask_col_stride=0,
                                     use_sdot=use_sdot)
    want = streaming_reference(q, k_ref, v_ref, d, seq_kv, scale, mask)
    assert np.max(np.abs(out - want)) < (0.01 if use_sdot else 1e-5)


def test_attention_col_offset_selects_slab():  # src/quant.zig:1487-1545
    d, slab = 32, 6
    k_big, v_big = oracle.QuantizedKVCache(d, 2 * slab, 32), oracle.QuantizedKVCache(d, 2 * slab, 32)
    k_small, v_small = oracle.QuantizedKVCache(d, slab, 32), oracle.QuantizedKVCache(d, slab, 32)
    k0, v0, k1, v1 = fill(d * slab, 30), fill(d * slab, 31), fill(d * slab, 32), fill(d * slab, 33)
    for c in range(slab):
        k_big.store_column(c, k0[c * d:(c + 1) * d]); v_big.store_column(c, v0[c * d:(c + 1) * d])
        k_big.store_column(slab + c, k1[c * d:(c + 1) * d]); v_big.store_column(slab + c, v1[c * d:(c + 1) * d])
        k_small.store_column(c, k1[c * d:(c + 1) * d]); v_small.store_column(c, v1[c * d:(c + 1) * d])
    q = fill(d, 34)
    scale = np.float32(1.0 / np.sqrt(d))
    big = oracle.attention_quantized(q, 1, k_big, slab, v_big, slab, slab, scale)
    small = oracle.attention_quantized(q, 1, k_small, 0, v_small, 0, slab, scale)
    assert np.max(np.abs(big - small)) < 1e-6


@pytest.mark.parametrize("seq_kv", [21, 7, 8, 16, 40])
def test_attention_tile_plus_tail_matches_single_column_reference(seq_kv):  # src/quant.zig:1547-1620 (seq_kv = 21)
    d = 64
    k, v, k_ref, v_ref = filled_caches(d, seq_kv, 32, 40, 41)
    q = fill(d, 42)
    scale = np.float32(1.0 / np.sqrt(d))
    out = oracle.attention_quantized(q, 1, k, 0, v, 0, seq_kv, scale)
    want = streaming_reference(q, k_ref, v_ref, d, seq_kv, scale)
    assert np.max(np.abs(out - want)) < 0.01
    out_f = oracle.attention_quantized(q, 1, k, 0, v, 0, seq_kv, scale, use_sdot=False)
    assert np.max(np.abs(out_f - want)) < 1e-5


def test_attention_prefill_columns_masks_and_fully_masked_rows():
    """seq_q > 1 with a [seq_kv, seq_q] causal mask (strides 1, seq_kv), column strides on q / dst, a fully masked
    query column (zeros, src/quant.zig:1075) and the d_head limits."""
    d, seq_kv, seq_q = 32, 19, 3
    k, v, k_ref, v_ref = filled_caches(d, seq_kv, 16, 50, 51)
    q = fill((d + 5) * seq_q, 52)
    mask = np.zeros((seq_q, seq_kv), np.float32)      # column qi at qi * seq_kv
    mask[0, 5:] = -np.inf
    mask[1, :] = -np.inf                                # fully masked
    mask[2, 17:] = -np.inf
    scale = np.float32(0.2)
    out = oracle.attention_quantized(q, seq_q, k, 0, v, 0, seq_kv, scale, mask=mask, mask_row_stride=1, mask_col_stride=seq_kv,
                                     use_sdot=False, q_col_stride=d + 5, dst_col_stride=d + 2)
    for qi in range(seq_q):
        got = out[qi * (d + 2):qi * (d + 2) + d]
        if qi == 1:
            assert not got.any()
            continue
        want = streaming_reference(q[qi * (d + 5):qi * (d + 5) + d], k_ref, v_ref, d, seq_kv, scale, mask[qi])
        assert np.max(np.abs(got - want)) < 1e-5
    with pytest.raises(AssertionError):
        oracle.QuantizedKVCache(48, 1, 32)            # d_head % block_size != 0 (src/quant.zig:659)
