/*
 * zgml_oracle.c — CPU restatement (plain C) of zgml's block-quantized
 * dequantize-and-dot path and of the DeviceOp reference executor.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call it, and only as the checker / timed CPU baseline.
 *
 * Parity pinning: the reference (Zig >= 0.16) cannot be built in this image
 * (no zig toolchain; its W8A8 kernel is aarch64-only inline asm,
 * src/quant.zig:347-354), so this restatement is pinned against every
 * known-answer vector the reference's own tests hold for the path (SURVEY.md
 * §8c; tests/test_oracle_golden.py).  End-to-end GGUF-direct LLaMA logits are
 * not pinned by any reference test ("parity unpinned" for configs 1/3/4/5 beyond
 * the per-op vectors).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (Zig's default float mode is
 * strict and quant.zig uses no @mulAdd, so products and sums round separately).
 *
 * Every function cites the reference file:line it follows
 * (paths relative to the reference root).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/zgml_cuda.h"

#define ZO_API __attribute__((visibility("default")))

static inline size_t zo_min(size_t a, size_t b) { return a < b ? a : b; }
static inline size_t zo_max(size_t a, size_t b) { return a > b ? a : b; }

/* f16 bits -> f32, exact (Zig: @floatCast(@as(f16, @bitCast(bytes)))). */
ZO_API float zo_f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu;
    uint32_t man = h & 0x3FFu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal: normalise */
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            man &= 0x3FFu;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7F800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

static inline float zo_clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ── QuantizedWeight.fromSlice — src/quant.zig:216-256 ─────────────────────
 * per flat block: max_abs; scale = max_abs/127 (1.0 if 0); q = trunc(clamp(v*(127/max_abs), ±127)). */
ZO_API void zo_from_slice(const float* weights, size_t rows, size_t cols, size_t block_size,
                          int8_t* data, float* scales) {
    size_t n_elems = rows * cols;
    size_t n_blocks = (n_elems + block_size - 1) / block_size;
    for (size_t b = 0; b < n_blocks; b++) {
        size_t start = b * block_size, end = zo_min(start + block_size, n_elems);
        float max_abs = 0;
        for (size_t j = start; j < end; j++) {
            float a = fabsf(weights[j]);
            if (a > max_abs) max_abs = a;
        }
        float scale = max_abs > 0 ? max_abs / 127.0f : 1.0f;
        scales[b] = scale;
        float inv_scale = max_abs > 0 ? 127.0f / max_abs : 0.0f;
        for (size_t j = start; j < end; j++) {
            float q = weights[j] * inv_scale;
            data[j] = (int8_t)zo_clampf(q, -127.0f, 127.0f); /* @intFromFloat truncates */
        }
    }
}

/* ── prepareTransposed — src/quant.zig:274-317 (twin: src/backend/reference.zig:26-70)
 * dequantize [K,N] then RE-quantize (lossy) into [N,K] with K-aligned blocks. */
ZO_API void zo_prepare_transposed(const int8_t* data, const float* scales, size_t K, size_t N,
                                  size_t bs, int8_t* t_data, float* t_scales) {
    size_t bpr = (K + bs - 1) / bs;
    for (size_t n = 0; n < N; n++) {
        for (size_t b = 0; b < bpr; b++) {
            size_t k_start = b * bs, k_end = zo_min(k_start + bs, K);
            float max_abs = 0;
            for (size_t k = k_start; k < k_end; k++) {
                size_t flat = k * N + n;
                float val = (float)data[flat] * scales[flat / bs];
                float a = fabsf(val);
                if (a > max_abs) max_abs = a;
            }
            float scale = max_abs > 0 ? max_abs / 127.0f : 1.0f;
            float inv_scale = max_abs > 0 ? 127.0f / max_abs : 0.0f;
            t_scales[n * bpr + b] = scale;
            for (size_t k = k_start; k < k_end; k++) {
                size_t flat = k * N + n;
                float val = (float)data[flat] * scales[flat / bs];
                float q = val * inv_scale;
                t_data[n * K + k] = (int8_t)zo_clampf(q, -127.0f, 127.0f);
            }
        }
    }
}

/* ── quantizeInput — src/quant.zig:320-341 ─────────────────────────────────── */
ZO_API void zo_quantize_input(const float* input, size_t K, size_t bs, int8_t* inp_q,
                              float* inp_scales) {
    size_t bpr = (K + bs - 1) / bs;
    for (size_t b = 0; b < bpr; b++) {
        size_t k_start = b * bs, k_end = zo_min(k_start + bs, K);
        float max_abs = 0;
        for (size_t k = k_start; k < k_end; k++) {
            float a = fabsf(input[k]);
            if (a > max_abs) max_abs = a;
        }
        float scale = max_abs > 0 ? max_abs / 127.0f : 1.0f;
        float inv_scale = max_abs > 0 ? 127.0f / max_abs : 0.0f;
        inp_scales[b] = scale;
        for (size_t k = k_start; k < k_end; k++)
            inp_q[k] = (int8_t)zo_clampf(input[k] * inv_scale, -127.0f, 127.0f);
    }
}

/* ── gemvRange (W8A8) — src/quant.zig:358-440 ────────────────────────────────
 * The aarch64 `sdot` (16 lanes of i8*i8 -> 4 x i32, then @reduce) is an exact
 * integer dot product, so a scalar i32 loop is bit-identical.  Per output the
 * f32 accumulation is over K-blocks ascending: acc += f32(int) * (s_x[b]*s_w[n,b]). */
ZO_API void zo_gemv_range(const int8_t* t_d, const float* t_s, const int8_t* inp_q,
                          const float* inp_scales, float* dst, size_t n_start, size_t n_end,
                          size_t K, size_t bs) {
    size_t bpr = (K + bs - 1) / bs;
    for (size_t n = n_start; n < n_end; n++) {
        float acc = 0;
        for (size_t b = 0; b < bpr; b++) {
            size_t k_start = b * bs, k_end = zo_min(k_start + bs, K);
            float combined = inp_scales[b] * t_s[n * bpr + b];
            int32_t int_acc = 0;
            for (size_t k = k_start; k < k_end; k++)
                int_acc += (int32_t)inp_q[k] * (int32_t)t_d[n * K + k];
            acc += (float)int_acc * combined;
        }
        dst[n] = acc;
    }
}

/* ── gemv — src/quant.zig:443-459 (stack limits K<=16384, K/bs<=512) ────────── */
ZO_API int zo_gemv(const int8_t* t_d, const float* t_s, const float* input, float* dst, size_t N,
                   size_t K, size_t bs) {
    size_t bpr = (K + bs - 1) / bs;
    if (K > 16384 || bpr > 512) return -1;
    int8_t inp_q[16384];
    float inp_scales[512];
    zo_quantize_input(input, K, bs, inp_q, inp_scales);
    zo_gemv_range(t_d, t_s, inp_q, inp_scales, dst, 0, N, K, bs);
    return 0;
}

/* ── GemvPool.dispatch partitioning — src/quant.zig:135-196 ───────────────────
 * n_active = min(max(1, N*K / 2^20), n_workers), n_workers <= 16; N split into
 * 4-aligned chunks; the caller runs chunk 0.  Threads are created per call here
 * (the reference keeps a condvar pool; partitioning and arithmetic are the same). */
typedef struct {
    const int8_t* t_d; const float* t_s; const int8_t* inp_q; const float* inp_scales;
    float* dst; size_t n_start, n_end, K, bs;
} zo_gemv_task;
static void* zo_gemv_worker(void* arg) {
    zo_gemv_task* t = (zo_gemv_task*)arg;
    zo_gemv_range(t->t_d, t->t_s, t->inp_q, t->inp_scales, t->dst, t->n_start, t->n_end, t->K, t->bs);
    return NULL;
}
ZO_API int zo_gemv_pool(const int8_t* t_d, const float* t_s, const float* input, float* dst,
                        size_t N, size_t K, size_t bs, size_t n_workers) {
    size_t bpr = (K + bs - 1) / bs;
    if (K > 16384 || bpr > 512) return -1;
    int8_t inp_q[16384];
    float inp_scales[512];
    zo_quantize_input(input, K, bs, inp_q, inp_scales);
    if (n_workers > 16) n_workers = 16;
    size_t useful = zo_max(1, (N * K) / (1024 * 1024));
    size_t n_active = zo_min(useful, n_workers);
    if (n_active <= 1) {
        zo_gemv_range(t_d, t_s, inp_q, inp_scales, dst, 0, N, K, bs);
        return 1;
    }
    size_t chunk = (((N + n_active - 1) / n_active) + 3) & ~(size_t)3;
    pthread_t th[16];
    zo_gemv_task tasks[16];
    size_t n_disp = 0, n_start = chunk;
    for (size_t i = 1; i < n_active; i++) {
        if (n_start >= N) break;
        tasks[i] = (zo_gemv_task){t_d, t_s, inp_q, inp_scales, dst, n_start, zo_min(n_start + chunk, N), K, bs};
        pthread_create(&th[i], NULL, zo_gemv_worker, &tasks[i]);
        n_disp++;
        n_start += chunk;
    }
    zo_gemv_range(t_d, t_s, inp_q, inp_scales, dst, 0, zo_min(chunk, N), K, bs);
    for (size_t i = 1; i <= n_disp; i++) pthread_join(th[i], NULL);
    return (int)(n_disp + 1);
}

/* ── QuantizedKVCache + attentionQuantized — src/quant.zig:633-1091 ─────────
 * Column-major Q8 cache: column c = d_head int8 at q_data[c*d_head], d_head/bs scales at
 * scales[c*bpc].  storeColumn = quantizeInput on the column (src/quant.zig:689-701). */
ZO_API void zo_kv_store_column(int8_t* q_data, float* scales, size_t d_head, size_t bs, size_t col,
                               const float* src) {
    size_t bpc = d_head / bs;
    zo_quantize_input(src, d_head, bs, q_data + col * d_head, scales + col * bpc);
}

/* dequantColumn — src/quant.zig:704-716 */
ZO_API void zo_kv_dequant_column(const int8_t* q_data, const float* scales, size_t d_head, size_t bs,
                                 size_t col, float* dst) {
    size_t bpc = d_head / bs;
    for (size_t b = 0; b < bpc; b++)
        for (size_t i = 0; i < bs; i++)
            dst[b * bs + i] = (float)q_data[col * d_head + b * bs + i] * scales[col * bpc + b];
}

/* dotI8I8 — src/quant.zig:764-798: per block an exact int32 dot, total += f32(int) * q_s[b] * k_s[b]
 * (left to right), blocks ascending. */
static float zo_dot_i8i8(const int8_t* q_i8, const float* q_s, const int8_t* k_i8, const float* k_s,
                         size_t bs, size_t nb) {
    float total = 0;
    for (size_t b = 0; b < nb; b++) {
        int32_t acc = 0;
        for (size_t j = 0; j < bs; j++) acc += (int32_t)q_i8[b * bs + j] * (int32_t)k_i8[b * bs + j];
        total += (float)acc * q_s[b] * k_s[b];
    }
    return total;
}

/* dotI8F32 — src/quant.zig:800-830: 8 vector lanes of partial sums per block, @reduce(.Add) taken as
 * lanes 0..7 in order (strict float mode), scalar tail, total += sub * scale[b]. */
static float zo_dot_i8f32(const float* f, const int8_t* q, const float* sc, size_t bs, size_t nb) {
    float total = 0;
    for (size_t b = 0; b < nb; b++) {
        float lane[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        size_t i = 0;
        for (; i + 8 <= bs; i += 8)
            for (size_t v = 0; v < 8; v++) lane[v] += f[b * bs + i + v] * (float)q[b * bs + i + v];
        float sub = lane[0];
        for (size_t v = 1; v < 8; v++) sub += lane[v];
        for (; i < bs; i++) sub += f[b * bs + i] * (float)q[b * bs + i];
        total += sub * sc[b];
    }
    return total;
}

/* accumI8F32 — src/quant.zig:832-860: acc[i] = acc[i] + (w * scale[b]) * f32(q[i]) */
static void zo_accum_i8f32(float* acc, float w, const int8_t* q, const float* sc, size_t bs, size_t nb) {
    for (size_t b = 0; b < nb; b++) {
        float ws = w * sc[b];
        for (size_t i = 0; i < bs; i++) acc[b * bs + i] = acc[b * bs + i] + ws * (float)q[b * bs + i];
    }
}

/* accumBatchI8F32 — src/quant.zig:865-908 with Bs = 8: per block ws_scaled[c] = ws[c] * scale of column
 * col_start + c; per element sum = acc, then += ws_scaled[c] * f32(q) for c = 0..7 in order. */
static void zo_accum_batch8(float* acc, const float* ws, const int8_t* q_data, const float* scales,
                            size_t col_start, size_t d_head, size_t bs, size_t nb) {
    for (size_t bl = 0; bl < nb; bl++) {
        float wsc[8];
        for (size_t c = 0; c < 8; c++) wsc[c] = ws[c] * scales[(col_start + c) * nb + bl];
        for (size_t i = 0; i < bs; i++) {
            float sum = acc[bl * bs + i];
            for (size_t c = 0; c < 8; c++) sum = sum + wsc[c] * (float)q_data[(col_start + c) * d_head + bl * bs + i];
            acc[bl * bs + i] = sum;
        }
    }
}

/* attentionQuantized — src/quant.zig:924-1091.  use_sdot = the aarch64 branch (query quantized per column with
 * the cache's block size, int8 x int8 scores); 0 = the portable branch (f32 query x int8 keys).  Tiles of 8
 * kv positions with one rescale per tile, then a one-at-a-time tail; non-finite mask entries skip the
 * position, a fully masked query column gives zeros.  Returns -1 when d_head > 512 (the reference's stack
 * buffers) or d_head % bs != 0. */
ZO_API int zo_attention_quantized(float* dst, size_t dst_cs, const float* q, size_t q_cs, size_t d_head, size_t seq_q,
                                  const int8_t* k_q, const float* k_s, size_t k_col_start, const int8_t* v_q,
                                  const float* v_s, size_t v_col_start, size_t bs, size_t seq_kv, const float* mask,
                                  size_t mask_rs, size_t mask_cs, float scale, int use_sdot) {
    if (d_head > 512 || bs == 0 || d_head % bs != 0) return -1;
    size_t nb = d_head / bs;
    if (use_sdot && nb > 32) return -1; /* q_scales_buf holds 512 / 16 entries */
    float acc[512];
    int8_t q_i8[512];
    float q_sc[512];
    const float neg_inf = -INFINITY;
    for (size_t qi = 0; qi < seq_q; qi++) {
        const float* q_col = q + qi * q_cs;
        size_t mask_base = qi * mask_cs;
        if (use_sdot) zo_quantize_input(q_col, d_head, bs, q_i8, q_sc);
        float m_val = neg_inf, l = 0;
        memset(acc, 0, d_head * sizeof(float));
        size_t s = 0;
        for (; s + 8 <= seq_kv; s += 8) {
            if (mask) {
                int any_valid = 0;
                for (size_t b = 0; b < 8; b++)
                    if (isfinite(mask[mask_base + (s + b) * mask_rs])) any_valid = 1;
                if (!any_valid) continue;
            }
            float scores[8], ws[8], tile_max = neg_inf;
            for (size_t b = 0; b < 8; b++) {
                float mask_add = mask ? mask[mask_base + (s + b) * mask_rs] : 0.0f;
                if (isfinite(mask_add)) {
                    size_t c = k_col_start + s + b;
                    float dot = use_sdot ? zo_dot_i8i8(q_i8, q_sc, k_q + c * d_head, k_s + c * nb, bs, nb)
                                         : zo_dot_i8f32(q_col, k_q + c * d_head, k_s + c * nb, bs, nb);
                    float score = dot * scale + mask_add;
                    scores[b] = score;
                    if (score > tile_max) tile_max = score;
                } else {
                    scores[b] = neg_inf;
                }
            }
            if (tile_max == neg_inf) continue;
            float new_m = (m_val == neg_inf) ? tile_max : (m_val > tile_max ? m_val : tile_max);
            float alpha = (m_val == neg_inf) ? 0.0f : expf(m_val - new_m);
            float tile_l = 0;
            for (size_t b = 0; b < 8; b++) {
                float w = expf(scores[b] - new_m);
                ws[b] = w;
                tile_l += w;
            }
            if (m_val != neg_inf && alpha != 1.0f)
                for (size_t r = 0; r < d_head; r++) acc[r] = acc[r] * alpha;
            zo_accum_batch8(acc, ws, v_q, v_s, v_col_start + s, d_head, bs, nb);
            l = l * alpha + tile_l;
            m_val = new_m;
        }
        for (; s < seq_kv; s++) {
            float mask_add = mask ? mask[mask_base + s * mask_rs] : 0.0f;
            if (!isfinite(mask_add)) continue;
            size_t c = k_col_start + s;
            float dot = use_sdot ? zo_dot_i8i8(q_i8, q_sc, k_q + c * d_head, k_s + c * nb, bs, nb)
                                 : zo_dot_i8f32(q_col, k_q + c * d_head, k_s + c * nb, bs, nb);
            float score = dot * scale + mask_add;
            if (!isfinite(score)) continue;
            float new_m = m_val > score ? m_val : score;
            float alpha = (m_val == neg_inf) ? 0.0f : expf(m_val - new_m);
            float w = expf(score - new_m);
            if (m_val != neg_inf && alpha != 1.0f)
                for (size_t r = 0; r < d_head; r++) acc[r] = acc[r] * alpha;
            size_t vc = v_col_start + s;
            zo_accum_i8f32(acc, w, v_q + vc * d_head, v_s + vc * nb, bs, nb);
            l = l * alpha + w;
            m_val = new_m;
        }
        float inv_l = l > 0 ? 1.0f / l : 0.0f;
        for (size_t r = 0; r < d_head; r++) dst[qi * dst_cs + r] = acc[r] * inv_l;
    }
    return 0;
}

/* ── QuantizedWeight.matmul (W8·f32) — src/quant.zig:475-578 ─────────────────
 * dst zeroed; loop M -> K (unrolled x4) -> N in chunks cut at the FIRST row's
 * block boundary.  Vector path (8 lanes): d += f32(q) * (scale*x), one k after
 * the other; scalar tail (<8 left in the chunk): d += (x*f32(q))*scale.
 * Note (faithful quirk): inside the k-unrolled loop the chunk boundary and the
 * single per-chunk scale of rows ki>0 are derived from their own flat index at
 * the chunk START, so when N % bs != 0 a chunk can straddle a block of row ki>0
 * and use one scale for all of it.  N % bs == 0 (every model shape) is unaffected.
 * Restricted to output columns [n_lo, n_hi) so the threaded baseline can split N;
 * n_lo must be a multiple of 8 so vector/scalar paths coincide with the full call. */
static void zo_matmul_cols(const int8_t* w_data, const float* scales, size_t bs, const float* input,
                           float* dst, size_t M, size_t N, size_t K, size_t n_lo, size_t n_hi) {
    const size_t vec_len = 8, k_unroll = 4;
    for (size_t m = 0; m < M; m++) {
        float* dst_row = dst + m * N;
        const float* inp_row = input + m * K;
        for (size_t n = n_lo; n < n_hi; n++) dst_row[n] = 0;
        size_t k = 0;
        for (; k + k_unroll <= K; k += k_unroll) {
            float inp_vals[4];
            size_t w_bases[4];
            for (size_t ki = 0; ki < k_unroll; ki++) {
                inp_vals[ki] = inp_row[k + ki];
                w_bases[ki] = (k + ki) * N;
            }
            size_t n = 0;
            while (n < N) {
                size_t flat0 = w_bases[0] + n;
                size_t block_rem = bs - (flat0 % bs);
                size_t chunk = zo_min(block_rem, N - n);
                if (n >= n_hi) break;
                if (n + chunk <= n_lo) { n += chunk; continue; }
                float combined[4];
                for (size_t ki = 0; ki < k_unroll; ki++)
                    combined[ki] = scales[(w_bases[ki] + n) / bs] * inp_vals[ki];
                size_t vec_end = (chunk / vec_len) * vec_len;
                size_t jlo = n < n_lo ? n_lo - n : 0;            /* n_lo % 8 == 0 */
                size_t jhi = n + chunk > n_hi ? n_hi - n : chunk;
                size_t j = jlo;
                for (; j + vec_len <= vec_end && j + vec_len <= jhi; j += vec_len) {
                    float d[8];
                    for (size_t l = 0; l < 8; l++) d[l] = dst_row[n + j + l];
                    for (size_t ki = 0; ki < k_unroll; ki++) {
                        const int8_t* wp = w_data + w_bases[ki] + n + j;
                        float c = combined[ki];
                        for (size_t l = 0; l < 8; l++) {
                            float p = (float)wp[l] * c;
                            d[l] = d[l] + p;
                        }
                    }
                    for (size_t l = 0; l < 8; l++) dst_row[n + j + l] = d[l];
                }
                for (; j < jhi; j++) {
                    size_t col = n + j;
                    if (j < vec_end) { /* partial vector group at a thread boundary: vector arithmetic */
                        float d = dst_row[col];
                        for (size_t ki = 0; ki < k_unroll; ki++) {
                            float p = (float)w_data[w_bases[ki] + col] * combined[ki];
                            d = d + p;
                        }
                        dst_row[col] = d;
                        continue;
                    }
                    for (size_t ki = 0; ki < k_unroll; ki++) {
                        size_t flat_ki = w_bases[ki] + col;
                        float p = inp_vals[ki] * (float)w_data[flat_ki];
                        p = p * scales[flat_ki / bs];
                        dst_row[col] = dst_row[col] + p;
                    }
                }
                n += chunk;
            }
        }
        for (; k < K; k++) {
            float inp_val = inp_row[k];
            size_t w_base = k * N;
            size_t n = 0;
            while (n < N) {
                size_t flat = w_base + n;
                float scale = scales[flat / bs];
                float combined_s = scale * inp_val;
                size_t block_rem = bs - (flat % bs);
                size_t chunk = zo_min(block_rem, N - n);
                if (n >= n_hi) break;
                if (n + chunk <= n_lo) { n += chunk; continue; }
                size_t vec_end = (chunk / vec_len) * vec_len;
                size_t jlo = n < n_lo ? n_lo - n : 0;
                size_t jhi = n + chunk > n_hi ? n_hi - n : chunk;
                for (size_t j = jlo; j < jhi; j++) {
                    size_t col = n + j;
                    float p;
                    if (j < vec_end) {
                        p = (float)w_data[flat + j] * combined_s;
                    } else {
                        p = inp_val * (float)w_data[flat + j];
                        p = p * scale;
                    }
                    dst_row[col] = dst_row[col] + p;
                }
                n += chunk;
            }
        }
    }
}

ZO_API void zo_matmul(const int8_t* w_data, const float* scales, size_t bs, const float* input,
                      float* dst, size_t M, size_t N, size_t K) {
    zo_matmul_cols(w_data, scales, bs, input, dst, M, N, K, 0, N);
}

/* matmulBias — src/quant.zig:581-589 */
ZO_API void zo_matmul_bias(const int8_t* w_data, const float* scales, size_t bs, const float* input,
                           const float* bias, float* dst, size_t M, size_t N, size_t K) {
    zo_matmul(w_data, scales, bs, input, dst, M, N, K);
    for (size_t m = 0; m < M; m++)
        for (size_t n = 0; n < N; n++) dst[m * N + n] += bias[n];
}

/* Threaded W8·f32 baseline: the reference runs src/quant.zig:475-578 on ONE
 * thread (src/inference_utils.zig:192); this splits the N columns over
 * `n_threads` with per-column arithmetic identical to zo_matmul, so the result is
 * bit-identical to the single-threaded call.  Used only as the all-cores CPU
 * baseline in bench.py. */
typedef struct {
    const int8_t* w; const float* s; size_t bs; const float* in; float* dst;
    size_t M, N, K, n_lo, n_hi;
} zo_mm_task;
static void* zo_mm_worker(void* arg) {
    zo_mm_task* t = (zo_mm_task*)arg;
    zo_matmul_cols(t->w, t->s, t->bs, t->in, t->dst, t->M, t->N, t->K, t->n_lo, t->n_hi);
    return NULL;
}
ZO_API void zo_matmul_mt(const int8_t* w_data, const float* scales, size_t bs, const float* input,
                         float* dst, size_t M, size_t N, size_t K, size_t n_threads) {
    if (n_threads <= 1 || N < 64) { zo_matmul(w_data, scales, bs, input, dst, M, N, K); return; }
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    zo_mm_task tasks[256];
    size_t chunk = (((N + n_threads - 1) / n_threads) + 31) & ~(size_t)31;
    size_t nt = 0;
    for (size_t lo = 0; lo < N; lo += chunk, nt++) {
        tasks[nt] = (zo_mm_task){w_data, scales, bs, input, dst, M, N, K, lo, zo_min(lo + chunk, N)};
        if (nt > 0) pthread_create(&th[nt], NULL, zo_mm_worker, &tasks[nt]);
    }
    zo_mm_worker(&tasks[0]);
    for (size_t i = 1; i < nt; i++) pthread_join(th[i], NULL);
}

/* ── dequantizeTo — src/quant.zig:594-618: dst[j] = f32(q[j]) * scales[j/bs] ── */
ZO_API void zo_dequantize_to(const int8_t* data, const float* scales, size_t n_elems, size_t bs,
                             float* dst) {
    for (size_t j = 0; j < n_elems; j++) dst[j] = (float)data[j] * scales[j / bs];
}

/* ── GGUF block decode — src/models/gguf_loader.zig:33-80 ─────────────────────
 * zgml nibble order (NOT ggml's): element i in byte i/2, even -> low nibble. */
ZO_API void zo_dequant_q4_0(float* dst, const uint8_t* src, size_t n_elems) {
    size_t n_blocks = (n_elems + 31) / 32;
    for (size_t b = 0; b < n_blocks; b++) {
        size_t off = b * 18;
        float scale = zo_f16_to_f32((uint16_t)(src[off] | (src[off + 1] << 8)));
        size_t elems = zo_min(32, n_elems - b * 32);
        for (size_t i = 0; i < elems; i++) {
            uint8_t byte = src[off + 2 + i / 2];
            uint8_t nib = (i % 2 == 0) ? (byte & 0x0F) : (byte >> 4);
            float sv = (float)(int8_t)((int16_t)nib - 8);
            dst[b * 32 + i] = sv * scale;
        }
    }
}
ZO_API void zo_dequant_q8_0(float* dst, const uint8_t* src, size_t n_elems) {
    size_t n_blocks = (n_elems + 31) / 32;
    for (size_t b = 0; b < n_blocks; b++) {
        size_t off = b * 34;
        float scale = zo_f16_to_f32((uint16_t)(src[off] | (src[off + 1] << 8)));
        size_t elems = zo_min(32, n_elems - b * 32);
        for (size_t i = 0; i < elems; i++) dst[b * 32 + i] = (float)(int8_t)src[off + 2 + i] * scale;
    }
}
ZO_API void zo_dequant_f16(float* dst, const uint8_t* src, size_t n_elems) {
    for (size_t i = 0; i < n_elems; i++)
        dst[i] = zo_f16_to_f32((uint16_t)(src[2 * i] | (src[2 * i + 1] << 8)));
}

/* ── quantizedWeightFromInfo — src/models/gguf_loader.zig:99-154 ───────────────
 * GGUF Q8_0 (type 8) / Q4_0 (type 2) blocks -> i8 data + f32 scales, bs = 32;
 * rows = dims[0], cols = dims[1] (no transpose).  Returns 0, or -1 if unsupported. */
ZO_API int zo_qweight_from_gguf(const uint8_t* raw, uint32_t ggml_type, size_t n_elems,
                                int8_t* data, float* scales) {
    size_t n_blocks = (n_elems + 31) / 32;
    if (ggml_type == 8) {
        for (size_t b = 0; b < n_blocks; b++) {
            size_t off = b * 34;
            scales[b] = zo_f16_to_f32((uint16_t)(raw[off] | (raw[off + 1] << 8)));
            size_t elems = zo_min(32, n_elems - b * 32);
            for (size_t i = 0; i < elems; i++) data[b * 32 + i] = (int8_t)raw[off + 2 + i];
        }
        return 0;
    }
    if (ggml_type == 2) {
        for (size_t b = 0; b < n_blocks; b++) {
            size_t off = b * 18;
            scales[b] = zo_f16_to_f32((uint16_t)(raw[off] | (raw[off + 1] << 8)));
            size_t elems = zo_min(32, n_elems - b * 32);
            for (size_t i = 0; i < elems; i++) {
                uint8_t byte = raw[off + 2 + i / 2];
                uint8_t nib = (i % 2 == 0) ? (byte & 0x0F) : (byte >> 4);
                data[b * 32 + i] = (int8_t)((int16_t)nib - 8);
            }
        }
        return 0;
    }
    return -1;
}

/* ═══ DeviceOp reference executor — src/backend/reference.zig ═════════════════ */

typedef struct { float* ptr; size_t len; } zo_buffer;
typedef struct { const int8_t* data; const float* scales; size_t block_size; } zo_qweight;

#define ZO_V 8

static float zo_gelu_tanh(float a) { /* reference.zig:265-269 scalar tail */
    float kk = 0.7978845608f * (a + 0.044715f * a * a * a);
    return 0.5f * a * (1.0f + tanhf(kk));
}
static float zo_gelu_exp(float a) { /* reference.zig:258-264 vector body */
    float k = 0.7978845608f * (a + 0.044715f * a * a * a);
    float e2k = expf(k + k);
    return 0.5f * a * (1.0f + (e2k - 1.0f) / (e2k + 1.0f));
}

/* elementwise — reference.zig:201-273.  ops outside the switch memcpy src0. */
static void zo_elementwise(const zo_buffer* bufs, const ZgOp* op) {
    float* dst = bufs[op->u.elementwise.dst].ptr + op->u.elementwise.dst_offset;
    const float* s0 = bufs[op->u.elementwise.src0].ptr + op->u.elementwise.src0_offset;
    const float* s1 = bufs[op->u.elementwise.src1].ptr + op->u.elementwise.src1_offset;
    size_t n = op->u.elementwise.n;
    switch (op->u.elementwise.op) {
        case ZG_EW_ADD: for (size_t i = 0; i < n; i++) dst[i] = s0[i] + s1[i]; break;
        case ZG_EW_MUL: for (size_t i = 0; i < n; i++) dst[i] = s0[i] * s1[i]; break;
        case ZG_EW_NEG: for (size_t i = 0; i < n; i++) dst[i] = -s0[i]; break;
        case ZG_EW_ABS: for (size_t i = 0; i < n; i++) dst[i] = fabsf(s0[i]); break;
        case ZG_EW_RELU: for (size_t i = 0; i < n; i++) dst[i] = s0[i] > 0.0f ? s0[i] : 0.0f; break;
        case ZG_EW_SQRT: for (size_t i = 0; i < n; i++) dst[i] = sqrtf(s0[i]); break;
        case ZG_EW_RECIP: for (size_t i = 0; i < n; i++) dst[i] = 1.0f / s0[i]; break;
        case ZG_EW_EXP: for (size_t i = 0; i < n; i++) dst[i] = expf(s0[i]); break;
        case ZG_EW_LOG: for (size_t i = 0; i < n; i++) dst[i] = logf(s0[i]); break;
        case ZG_EW_GELU: {
            size_t i = 0;
            for (; i + ZO_V <= n; i += ZO_V)
                for (size_t l = 0; l < ZO_V; l++) dst[i + l] = zo_gelu_exp(s0[i + l]);
            for (; i < n; i++) dst[i] = zo_gelu_tanh(s0[i]);
            break;
        }
        default: memmove(dst, s0, n * sizeof(float)); break;
    }
}

/* fusedElementwise — reference.zig:275-307; unknown ops leave v unchanged. */
static void zo_fused_elementwise(const zo_buffer* bufs, const ZgOp* op) {
    float* dst = bufs[op->u.fused_elementwise.dst].ptr + op->u.fused_elementwise.dst_offset;
    const float* src = bufs[op->u.fused_elementwise.src].ptr + op->u.fused_elementwise.src_offset;
    size_t n = op->u.fused_elementwise.n;
    for (size_t i = 0; i < n; i++) {
        float v = src[i];
        for (size_t s = 0; s < op->u.fused_elementwise.n_steps; s++) {
            const ZgFusedEwStep* st = &op->u.fused_elementwise.steps[s];
            switch (st->op) {
                case ZG_EW_NEG: v = -v; break;
                case ZG_EW_ABS: v = fabsf(v); break;
                case ZG_EW_RELU: v = v > 0.0f ? v : 0.0f; break;
                case ZG_EW_SQRT: v = sqrtf(v); break;
                case ZG_EW_RECIP: v = 1.0f / v; break;
                case ZG_EW_EXP: v = expf(v); break;
                case ZG_EW_LOG: v = logf(v); break;
                case ZG_EW_GELU: v = zo_gelu_tanh(v); break;
                case ZG_EW_ADD: {
                    const float* sp = bufs[st->secondary_buf].ptr + st->secondary_offset;
                    v = st->is_swapped ? sp[i] + v : v + sp[i];
                    break;
                }
                case ZG_EW_MUL: {
                    const float* sp = bufs[st->secondary_buf].ptr + st->secondary_offset;
                    v = st->is_swapped ? sp[i] * v : v * sp[i];
                    break;
                }
                default: break;
            }
        }
        dst[i] = v;
    }
}

/* softmax — reference.zig:309-327 */
static void zo_softmax(const zo_buffer* bufs, const ZgOp* op) {
    const float* src = bufs[op->u.softmax.src].ptr;
    float* dst = bufs[op->u.softmax.dst].ptr;
    size_t cols = op->u.softmax.cols;
    for (size_t row = 0; row < op->u.softmax.rows; row++) {
        size_t sb = op->u.softmax.src_offset + row * cols, db = op->u.softmax.dst_offset + row * cols;
        float m = -INFINITY;
        for (size_t j = 0; j < cols; j++) m = fmaxf(m, src[sb + j]);
        float sum = 0;
        for (size_t j = 0; j < cols; j++) {
            float v = expf(src[sb + j] - m);
            dst[db + j] = v;
            sum += v;
        }
        float inv = sum > 0.0f ? 1.0f / sum : 0.0f;
        for (size_t j = 0; j < cols; j++) dst[db + j] *= inv;
    }
}

/* layernorm — reference.zig:329-347 (no affine) */
static void zo_layernorm(const zo_buffer* bufs, const ZgOp* op) {
    const float* src = bufs[op->u.layernorm.src].ptr;
    float* dst = bufs[op->u.layernorm.dst].ptr;
    size_t cols = op->u.layernorm.cols;
    for (size_t row = 0; row < op->u.layernorm.rows; row++) {
        size_t base = op->u.layernorm.src_offset + row * cols, dbase = op->u.layernorm.dst_offset + row * cols;
        float mu = 0;
        for (size_t j = 0; j < cols; j++) mu += src[base + j];
        mu /= (float)cols;
        float v = 0;
        for (size_t j = 0; j < cols; j++) { float d = src[base + j] - mu; v += d * d; }
        float inv_std = 1.0f / sqrtf(v / (float)cols + op->u.layernorm.eps);
        for (size_t j = 0; j < cols; j++) dst[dbase + j] = (src[base + j] - mu) * inv_std;
    }
}

/* rmsnorm — reference.zig:349-374: 8 lane partial sums, @reduce(.Add), scalar tail. */
static void zo_rmsnorm(const zo_buffer* bufs, const ZgOp* op) {
    const float* src = bufs[op->u.rmsnorm.src].ptr;
    float* dst = bufs[op->u.rmsnorm.dst].ptr;
    size_t cols = op->u.rmsnorm.cols;
    for (size_t row = 0; row < op->u.rmsnorm.rows; row++) {
        const float* s = src + op->u.rmsnorm.src_offset + row * cols;
        float* d = dst + op->u.rmsnorm.dst_offset + row * cols;
        float acc[ZO_V] = {0};
        size_t i = 0;
        for (; i + ZO_V <= cols; i += ZO_V)
            for (size_t l = 0; l < ZO_V; l++) acc[l] += s[i + l] * s[i + l];
        float ss = 0;
        for (size_t l = 0; l < ZO_V; l++) ss += acc[l];
        for (; i < cols; i++) ss += s[i] * s[i];
        float inv_rms = 1.0f / sqrtf(ss / (float)cols + op->u.rmsnorm.eps);
        for (i = 0; i < cols; i++) d[i] = s[i] * inv_rms;
    }
}

/* reduce — reference.zig:376-389 (sum / max over contiguous groups) */
static void zo_reduce(const zo_buffer* bufs, const ZgOp* op) {
    const float* src = bufs[op->u.reduce.src].ptr;
    float* dst = bufs[op->u.reduce.dst].ptr;
    size_t rs = op->u.reduce.reduce_size;
    int is_max = op->u.reduce.op == ZG_EW_MAX;
    for (size_t i = 0; i < op->u.reduce.n_out; i++) {
        size_t sb = op->u.reduce.src_offset + i * rs;
        float val = is_max ? -INFINITY : 0.0f;
        for (size_t k = 0; k < rs; k++) {
            float v = src[sb + k];
            val = is_max ? fmaxf(val, v) : val + v;
        }
        dst[op->u.reduce.dst_offset + i] = val;
    }
}

/* repeat — reference.zig:391-433 */
static void zo_repeat(const zo_buffer* bufs, const ZgOp* op) {
    const float* src = bufs[op->u.repeat.src].ptr;
    float* dst = bufs[op->u.repeat.dst].ptr;
    size_t n = op->u.repeat.n;
    float* d = dst + op->u.repeat.dst_offset;
    const float* s = src + op->u.repeat.src_offset;
    const uint32_t* ne = op->u.repeat.src_ne;
    const uint32_t* st = op->u.repeat.src_strides;
    size_t src_n = (size_t)ne[0] * ne[1] * ne[2] * ne[3];
    if (src_n == 1) { for (size_t i = 0; i < n; i++) d[i] = s[0]; return; }
    if (src_n >= n) { memmove(d, s, n * sizeof(float)); return; }
    if (n % src_n == 0 && st[0] == 1 && (ne[1] <= 1 || st[1] == ne[0]) &&
        (ne[2] <= 1 || st[2] == ne[0] * ne[1]) && (ne[3] <= 1 || st[3] == ne[0] * ne[1] * ne[2])) {
        for (size_t off = 0; off + src_n <= n; off += src_n) memmove(d + off, s, src_n * sizeof(float));
        return;
    }
    for (size_t gid = 0; gid < n; gid++) {
        size_t idx = gid, src_idx = op->u.repeat.src_offset;
        for (int dim = 3; dim >= 0; dim--) {
            size_t coord = idx / op->u.repeat.dst_strides[dim];
            idx = idx % op->u.repeat.dst_strides[dim];
            src_idx += (coord % ne[dim]) * st[dim];
        }
        dst[op->u.repeat.dst_offset + gid] = src[src_idx];
    }
}

/* sliceAssign — reference.zig:435-455 (uses dst_offset; dst_base_offset/patch_stride
 * only matter to the caller's patchSliceAssignOffset). */
static void zo_slice_assign(const zo_buffer* bufs, const ZgOp* op) {
    const float* src = bufs[op->u.slice_assign.src].ptr;
    float* dst = bufs[op->u.slice_assign.dst].ptr;
    size_t rows = op->u.slice_assign.rows, cols = op->u.slice_assign.cols;
    size_t doff = op->u.slice_assign.dst_offset, soff = op->u.slice_assign.src_offset;
    size_t drs = op->u.slice_assign.dst_row_stride, dcs = op->u.slice_assign.dst_col_stride;
    size_t srs = op->u.slice_assign.src_row_stride, scs = op->u.slice_assign.src_col_stride;
    for (size_t col = 0; col < cols; col++)
        for (size_t row = 0; row < rows; row++)
            dst[doff + row * drs + col * dcs] = src[soff + row * srs + col * scs];
}

/* rope — reference.zig:457-478 (half-split pairing, packed cos|sin columns) */
static void zo_rope(const zo_buffer* bufs, const ZgOp* op) {
    const float* src = bufs[op->u.rope.src].ptr;
    const float* cs = bufs[op->u.rope.cos_sin].ptr;
    float* dst = bufs[op->u.rope.dst].ptr;
    size_t hd = op->u.rope.half_d;
    for (size_t col = 0; col < op->u.rope.seq_len; col++) {
        for (size_t pair = 0; pair < hd; pair++) {
            float x_lo = src[op->u.rope.src_off + pair * op->u.rope.src_rs + col * op->u.rope.src_cs];
            float x_hi = src[op->u.rope.src_off + (pair + hd) * op->u.rope.src_rs + col * op->u.rope.src_cs];
            float c = cs[op->u.rope.cs_off + pair + col * op->u.rope.cs_cs];
            float sn = cs[op->u.rope.cs_off + pair + hd + col * op->u.rope.cs_cs];
            float a0 = x_lo * c, a1 = x_hi * sn;
            float b0 = x_hi * c, b1 = x_lo * sn;
            dst[op->u.rope.dst_off + pair + col * 2 * hd] = a0 - a1;
            dst[op->u.rope.dst_off + pair + hd + col * 2 * hd] = b0 + b1;
        }
    }
}

/* matmul — reference.zig:480-497 -> forward.blasSgemm (src/tensor/forward.zig:686-753):
 * dst[m*drs+n] = sum_k A[a_off + m*a_rs + k*a_cs] * B[b_off + k*b_rs + n*b_cs], k ascending. */
static void zo_matmul_dense(const zo_buffer* bufs, const ZgOp* op) {
    const ZgMatMulGeometry* g = &op->u.matmul.geom;
    const float* A = bufs[op->u.matmul.a].ptr;
    const float* B = bufs[op->u.matmul.b].ptr;
    float* dst = bufs[op->u.matmul.dst].ptr;
    for (size_t m = 0; m < g->M; m++)
        for (size_t n = 0; n < g->N; n++) {
            float acc = 0;
            for (size_t k = 0; k < g->K; k++) {
                float p = A[g->a_offset + m * g->a_row_stride + k * g->a_col_stride] *
                          B[g->b_offset + k * g->b_row_stride + n * g->b_col_stride];
                acc = acc + p;
            }
            dst[g->dst_offset + m * g->dst_row_stride + n] = acc;
        }
}

/* qmatmul — reference.zig:499-566 (x86 path: the aarch64 W8A8 shortcut at :512-528
 * is compiled out).  Row zeroed, loop K -> N in per-row block chunks; vector path
 * d + f32(q)*(scale*x); scalar tail d += f32(q)*(scale*x) (same product order). */
static void zo_qmatmul_op_cols(const float* input, float* dst_ptr, const int8_t* w_data,
                               const float* w_scales, size_t bs, size_t M, size_t N, size_t K,
                               size_t input_offset, size_t input_row_stride, size_t dst_offset,
                               size_t dst_row_stride, size_t n_lo, size_t n_hi) {
    if (input_row_stride == 0) input_row_stride = K;
    if (dst_row_stride == 0) dst_row_stride = N;
    for (size_t row = 0; row < M; row++) {
        const float* input_row = input + input_offset + row * input_row_stride;
        float* dst_row = dst_ptr + dst_offset + row * dst_row_stride;
        for (size_t n = n_lo; n < n_hi; n++) dst_row[n] = 0;
        for (size_t k = 0; k < K; k++) {
            float input_v = input_row[k];
            size_t w_base = k * N, n = n_lo;
            while (n < n_hi) {
                size_t flat = w_base + n;
                float scale = w_scales[flat / bs] * input_v;
                size_t block_rem = bs - (flat % bs);
                size_t chunk = zo_min(block_rem, n_hi - n);
                for (size_t j = 0; j < chunk; j++) {
                    float p = (float)w_data[flat + j] * scale;
                    dst_row[n + j] = dst_row[n + j] + p;
                }
                n += chunk;
            }
        }
    }
}
ZO_API void zo_qmatmul_op(const float* input, float* dst_ptr, const int8_t* w_data,
                          const float* w_scales, size_t bs, size_t M, size_t N, size_t K,
                          size_t input_offset, size_t input_row_stride, size_t dst_offset,
                          size_t dst_row_stride) {
    zo_qmatmul_op_cols(input, dst_ptr, w_data, w_scales, bs, M, N, K, input_offset, input_row_stride, dst_offset, dst_row_stride, 0, N);
}

/* All-host-cores variant for the program executor (bench.py --impl reference only): the reference runs qmatmul on one
 * thread (src/inference_utils.zig:192); here the output columns are split over `zo_exec_threads` threads.  Every output
 * column is accumulated by exactly the same operations in the same order, so results are bit-identical to one thread. */
static size_t zo_exec_threads = 1;
ZO_API void zo_set_exec_threads(size_t n) { zo_exec_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }
typedef struct {
    const float* input; float* dst; const int8_t* w; const float* s; size_t bs, M, N, K, io, irs, doff, drs, n_lo, n_hi;
} zo_qmm_task;
static void* zo_qmm_worker(void* arg) {
    zo_qmm_task* t = (zo_qmm_task*)arg;
    zo_qmatmul_op_cols(t->input, t->dst, t->w, t->s, t->bs, t->M, t->N, t->K, t->io, t->irs, t->doff, t->drs, t->n_lo, t->n_hi);
    return NULL;
}
static void zo_qmatmul_op_mt(const float* input, float* dst_ptr, const int8_t* w_data, const float* w_scales, size_t bs, size_t M,
                             size_t N, size_t K, size_t io, size_t irs, size_t doff, size_t drs) {
    size_t nt_want = zo_exec_threads;
    if (nt_want <= 1 || N < 256 || (bs && N % bs != 0)) { zo_qmatmul_op(input, dst_ptr, w_data, w_scales, bs, M, N, K, io, irs, doff, drs); return; }
    pthread_t th[256];
    zo_qmm_task tasks[256];
    size_t chunk = (((N + nt_want - 1) / nt_want) + 31) & ~(size_t)31, nt = 0;
    for (size_t lo = 0; lo < N; lo += chunk, nt++) {
        tasks[nt] = (zo_qmm_task){input, dst_ptr, w_data, w_scales, bs, M, N, K, io, irs, doff, drs, lo, zo_min(lo + chunk, N)};
        if (nt > 0) pthread_create(&th[nt], NULL, zo_qmm_worker, &tasks[nt]);
    }
    zo_qmm_worker(&tasks[0]);
    for (size_t i = 1; i < nt; i++) pthread_join(th[i], NULL);
}

/* attention — reference.zig:568-672: per query row, online softmax over seq_kv;
 * non-finite mask or score entries are skipped; 8-lane dot when q,k unit stride. */
static void zo_attention(const zo_buffer* bufs, const ZgOp* op) {
    const float* q_ptr = bufs[op->u.attention.q].ptr;
    const float* k_ptr = bufs[op->u.attention.k].ptr;
    const float* v_ptr = bufs[op->u.attention.v].ptr;
    const float* mask_ptr = bufs[op->u.attention.mask].ptr;
    float* dst = bufs[op->u.attention.dst].ptr;
    size_t dh = op->u.attention.d_head, sq = op->u.attention.seq_q, skv = op->u.attention.seq_kv;
    size_t qrs = op->u.attention.q_rs, qcs = op->u.attention.q_cs;
    size_t krs = op->u.attention.k_rs, kcs = op->u.attention.k_cs;
    size_t vrs = op->u.attention.v_rs, vcs = op->u.attention.v_cs;
    size_t mrs = op->u.attention.mask_rs, mcs = op->u.attention.mask_cs;
    size_t drs = op->u.attention.dst_rs, dcs = op->u.attention.dst_cs;
    float acc[512];
    for (size_t qi = 0; qi < sq; qi++) {
        size_t q_off = op->u.attention.q_off + qi * qcs;
        size_t d_off = op->u.attention.dst_off + qi * dcs;
        size_t mask_q_off = op->u.attention.mask_off + qi * mcs;
        float m_val = -INFINITY, l = 0;
        for (size_t r = 0; r < dh; r++) acc[r] = 0;
        for (size_t s = 0; s < skv; s++) {
            float mask_add = op->u.attention.has_mask ? mask_ptr[mask_q_off + s * mrs] : 0.0f;
            if (!isfinite(mask_add)) continue;
            float dot = 0;
            if (qrs == 1 && krs == 1) {
                float dv[ZO_V] = {0};
                size_t r = 0, kb = op->u.attention.k_off + s * kcs;
                for (; r + ZO_V <= dh; r += ZO_V)
                    for (size_t lne = 0; lne < ZO_V; lne++) dv[lne] += q_ptr[q_off + r + lne] * k_ptr[kb + r + lne];
                for (size_t lne = 0; lne < ZO_V; lne++) dot += dv[lne];
                for (; r < dh; r++) dot += q_ptr[q_off + r] * k_ptr[kb + r];
            } else {
                for (size_t r = 0; r < dh; r++)
                    dot += q_ptr[q_off + r * qrs] * k_ptr[op->u.attention.k_off + r * krs + s * kcs];
            }
            float score = dot * op->u.attention.scale + mask_add;
            if (!isfinite(score)) continue;
            float new_m = fmaxf(m_val, score);
            float alpha = (m_val == -INFINITY) ? 0.0f : expf(m_val - new_m);
            float w = expf(score - new_m);
            l = l * alpha + w;
            m_val = new_m;
            for (size_t r = 0; r < dh; r++)
                acc[r] = acc[r] * alpha + w * v_ptr[op->u.attention.v_off + r * vrs + s * vcs];
        }
        float inv_l = l > 0 ? 1.0f / l : 0.0f;
        for (size_t r = 0; r < dh; r++) dst[d_off + r * drs] = acc[r] * inv_l;
    }
}

/* executeProgram / executeOp — reference.zig:129-176 */
ZO_API void zo_execute_ops(float** buf_ptrs, const size_t* buf_lens, size_t n_bufs,
                           const ZgQWeight* qweights, size_t n_qweights, const ZgOp* ops,
                           size_t n_ops) {
    zo_buffer* bufs = (zo_buffer*)malloc(sizeof(zo_buffer) * (n_bufs ? n_bufs : 1));
    for (size_t i = 0; i < n_bufs; i++) { bufs[i].ptr = buf_ptrs[i]; bufs[i].len = buf_lens[i]; }
    (void)n_qweights;
    for (size_t i = 0; i < n_ops; i++) {
        const ZgOp* op = &ops[i];
        switch (op->tag) {
            case ZG_OP_ELEMENTWISE: zo_elementwise(bufs, op); break;
            case ZG_OP_MATMUL: zo_matmul_dense(bufs, op); break;
            case ZG_OP_QMATMUL: {
                const ZgQWeight* w = &qweights[op->u.qmatmul.weight_idx];
                zo_qmatmul_op_mt(bufs[op->u.qmatmul.input].ptr, bufs[op->u.qmatmul.dst].ptr, w->data,
                              w->scales, w->block_size, op->u.qmatmul.M, op->u.qmatmul.N,
                              op->u.qmatmul.K, op->u.qmatmul.input_offset,
                              op->u.qmatmul.input_row_stride, op->u.qmatmul.dst_offset,
                              op->u.qmatmul.dst_row_stride);
                break;
            }
            case ZG_OP_SOFTMAX: zo_softmax(bufs, op); break;
            case ZG_OP_LAYERNORM: zo_layernorm(bufs, op); break;
            case ZG_OP_RMSNORM: zo_rmsnorm(bufs, op); break;
            case ZG_OP_REDUCE: zo_reduce(bufs, op); break;
            case ZG_OP_REPEAT: zo_repeat(bufs, op); break;
            case ZG_OP_SLICE_ASSIGN: zo_slice_assign(bufs, op); break;
            case ZG_OP_ROPE: zo_rope(bufs, op); break;
            case ZG_OP_ATTENTION: zo_attention(bufs, op); break;
            case ZG_OP_FUSED_ELEMENTWISE: zo_fused_elementwise(bufs, op); break;
            default: break;
        }
    }
    free(bufs);
}

/* OwnedBufferTable + CpuBackend.compile/execute — reference.zig:77-127, cpu.zig:55-119:
 * zero-filled buffers of max(size,1) f32, initial uploads, run ops, download. */
ZO_API int zo_run_program(const ZgProgram* program, const ZgIO* inputs, size_t n_inputs,
                          const ZgIO* outputs, size_t n_outputs) {
    size_t nb = program->n_buffers;
    float** ptrs = (float**)calloc(nb ? nb : 1, sizeof(float*));
    size_t* lens = (size_t*)calloc(nb ? nb : 1, sizeof(size_t));
    for (size_t i = 0; i < nb; i++) {
        lens[i] = program->buffer_sizes[i] > 1 ? program->buffer_sizes[i] : 1;
        ptrs[i] = (float*)calloc(lens[i], sizeof(float));
        if (!ptrs[i]) return -1;
    }
    for (size_t i = 0; i < program->n_uploads; i++) {
        const ZgIO* io = &program->initial_uploads[i];
        memcpy((uint8_t*)ptrs[io->buf_idx] + io->offset, io->host_ptr, io->size);
    }
    for (size_t i = 0; i < n_inputs; i++)
        memcpy((uint8_t*)ptrs[inputs[i].buf_idx] + inputs[i].offset, inputs[i].host_ptr, inputs[i].size);
    zo_execute_ops(ptrs, lens, nb, program->qweights, program->n_qweights, program->ops, program->n_ops);
    for (size_t i = 0; i < n_outputs; i++)
        memcpy(outputs[i].host_ptr, (uint8_t*)ptrs[outputs[i].buf_idx] + outputs[i].offset, outputs[i].size);
    for (size_t i = 0; i < nb; i++) free(ptrs[i]);
    free(ptrs);
    free(lens);
    return 0;
}

/* Persistent variant for multi-step programs (KV cache carried between steps). */
typedef struct { size_t nb; float** ptrs; size_t* lens; } zo_state;
ZO_API void* zo_state_create(const ZgProgram* program) {
    zo_state* st = (zo_state*)calloc(1, sizeof(zo_state));
    st->nb = program->n_buffers;
    st->ptrs = (float**)calloc(st->nb ? st->nb : 1, sizeof(float*));
    st->lens = (size_t*)calloc(st->nb ? st->nb : 1, sizeof(size_t));
    for (size_t i = 0; i < st->nb; i++) {
        st->lens[i] = program->buffer_sizes[i] > 1 ? program->buffer_sizes[i] : 1;
        st->ptrs[i] = (float*)calloc(st->lens[i], sizeof(float));
    }
    for (size_t i = 0; i < program->n_uploads; i++) {
        const ZgIO* io = &program->initial_uploads[i];
        memcpy((uint8_t*)st->ptrs[io->buf_idx] + io->offset, io->host_ptr, io->size);
    }
    return st;
}
ZO_API void zo_state_execute(void* state, const ZgProgram* program, const ZgOp* ops, size_t n_ops,
                             const ZgIO* inputs, size_t n_inputs, const ZgIO* outputs, size_t n_outputs) {
    zo_state* st = (zo_state*)state;
    for (size_t i = 0; i < n_inputs; i++)
        memcpy((uint8_t*)st->ptrs[inputs[i].buf_idx] + inputs[i].offset, inputs[i].host_ptr, inputs[i].size);
    zo_execute_ops(st->ptrs, st->lens, st->nb, program->qweights, program->n_qweights, ops, n_ops);
    for (size_t i = 0; i < n_outputs; i++)
        memcpy(outputs[i].host_ptr, (uint8_t*)st->ptrs[outputs[i].buf_idx] + outputs[i].offset, outputs[i].size);
}
ZO_API void zo_state_destroy(void* state) {
    zo_state* st = (zo_state*)state;
    for (size_t i = 0; i < st->nb; i++) free(st->ptrs[i]);
    free(st->ptrs);
    free(st->lens);
    free(st);
}
