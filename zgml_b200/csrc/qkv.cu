// Quantized KV cache and attention over it for sm_100a — zgml's QuantizedKVCache / attentionQuantized
// (src/quant.zig:633-1091; callers src/llama_inference.zig:330-377), SURVEY.md §8f-2.
//
//   * Cache layout as the reference's: column-major Q8, column c = d_head int8 at q[c * d_head] plus d_head / bs f32
//     scales at s[c * bpc]; one column per (kv position) inside a head's slab of n_cols.  3.6x fewer bytes than the f32
//     cache at d_head 64, bs 32 — attention is the largest single HBM stream of long-context decode.
//   * storeColumn = quantizeInput on the column: a warp per (column, block), bit-identical data and scales.
//   * attentionQuantized: one CTA per query column, 8 warps stride over tiles of 32 kv positions.  Scores: lane = kv
//     position; the lane walks its own contiguous K column (int8 query branch: dp4a block dots and
//     f32(dot) * q_s[b] * k_s[b]; f32 query branch: the reference's eight partial sums per block) — the per-position
//     dot products are bit-identical to the reference's.  Softmax over the tile with warp shuffles, one rescale per
//     tile; V: lane = four head dimensions, positions of the tile one after the other (a 128-byte coalesced read per
//     position at d_head 128), masked / out-of-range positions skipped like the reference skips them.  The warps'
//     partial states (m, l, acc) merge through shared memory in warp order.  The softmax bookkeeping therefore groups
//     32 x 8 positions where the reference groups 8: outputs agree to float rounding (tests: 2e-6 absolute on O(1) values).
//
// Requirements beyond the reference's (d_head <= 512, d_head % bs == 0): d_head % 4 == 0 and bs % 4 == 0 (word loads).
#include "zg_internal.cuh"

#include <math.h>

struct ZgCudaKVCache {
    size_t d_head = 0, n_cols = 0, bs = 0, bpc = 0;
    int8_t* q = nullptr;
    float* s = nullptr;
};

namespace {

constexpr uint32_t kAttnWarps = 8, kMaxDHead = 512;

__device__ __forceinline__ float byte_f(uint32_t w, int e) { return (float)(int)(int8_t)(w >> (8 * e)); }

// a warp per (column, block): quantizeInput on column col_start + i (src/quant.zig:689-701, 320-341); source column i
// starts at src + i * src_cs
__device__ __forceinline__ void kv_store_body(const float* __restrict__ src, size_t src_cs, uint32_t d_head, uint32_t bs, uint32_t bpc, uint32_t n_write,
                                              int8_t* __restrict__ q, float* __restrict__ s, size_t col_start, uint32_t wid, uint32_t lane) {
    if (wid >= n_write * bpc) return;
    const uint32_t i = wid / bpc, b = wid % bpc;
    const float* x = src + (size_t)i * src_cs + (size_t)b * bs;
    float mx = 0.0f;
    for (uint32_t k = lane; k < bs; k += 32) mx = fmaxf(mx, fabsf(x[k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float scale = mx > 0.0f ? __fdiv_rn(mx, 127.0f) : 1.0f;
    const float inv = mx > 0.0f ? __fdiv_rn(127.0f, mx) : 0.0f;
    const size_t col = col_start + i;
    if (lane == 0) s[col * bpc + b] = scale;
    for (uint32_t k = lane; k < bs; k += 32) {
        float v = __fmul_rn(x[k], inv);
        v = v < -127.0f ? -127.0f : (v > 127.0f ? 127.0f : v);
        q[col * d_head + (size_t)b * bs + k] = (int8_t)(int)v;
    }
}
__global__ void k_kv_store(const float* __restrict__ src, uint32_t d_head, uint32_t bs, uint32_t bpc, uint32_t n_write,
                           int8_t* __restrict__ q, float* __restrict__ s, size_t col_start) {
    kv_store_body(src, d_head, d_head, bs, bpc, n_write, q, s, col_start, (blockIdx.x * blockDim.x + threadIdx.x) >> 5, threadIdx.x & 31);
}
// the cache-backed slice_assign ops of one dependency level of a program (blockIdx.y = op): the destination column comes
// from the op's run-time patched dst_offset (f32 elements into the cache buffer the Q8 cache stands in for)
__global__ void k_kv_store_tab(const ZgKvqStore* __restrict__ tab, const uint32_t* __restrict__ dyn) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const ZgKvqStore e = tab[blockIdx.y];
    kv_store_body(e.src, e.src_cs, e.d_head, e.bs, e.bpc, e.n_write, e.q, e.s, (size_t)(dyn[e.dyn_idx] / e.d_head),
                  (blockIdx.x * blockDim.x + threadIdx.x) >> 5, threadIdx.x & 31);
}

// dotI8I8 (src/quant.zig:764-798) of one K column with the quantized query in shared memory
__device__ float dot_i8(const uint32_t* __restrict__ kw, const float* __restrict__ ks, const uint32_t* qw, const float* qs,
                        uint32_t bs, uint32_t nb) {
    float total = 0.0f;
    const uint32_t wpb = bs / 4;
    for (uint32_t b = 0; b < nb; b++) {
        int acc = 0;
        for (uint32_t j = 0; j < wpb; j++) acc = __dp4a((int)__ldg(kw + b * wpb + j), (int)qw[b * wpb + j], acc);
        total = __fadd_rn(total, __fmul_rn(__fmul_rn((float)acc, qs[b]), __ldg(ks + b)));
    }
    return total;
}

// dotI8F32 (src/quant.zig:800-830): eight partial sums per block added in lane order, scalar tail, times the block scale
__device__ float dot_f32(const uint32_t* __restrict__ kw, const float* __restrict__ ks, const float* qf, uint32_t bs, uint32_t nb) {
    float total = 0.0f;
    const uint32_t wpb = bs / 4, vec_end = bs & ~7u;
    for (uint32_t b = 0; b < nb; b++) {
        float ln[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        const float* f = qf + (size_t)b * bs;
        uint32_t i = 0;
        for (; i < vec_end; i += 8) {
            const uint32_t w0 = __ldg(kw + b * wpb + i / 4), w1 = __ldg(kw + b * wpb + i / 4 + 1);
#pragma unroll
            for (int e = 0; e < 4; e++) {
                ln[e] = __fadd_rn(ln[e], __fmul_rn(f[i + e], byte_f(w0, e)));
                ln[4 + e] = __fadd_rn(ln[4 + e], __fmul_rn(f[i + 4 + e], byte_f(w1, e)));
            }
        }
        float sub = ln[0];
#pragma unroll
        for (int v = 1; v < 8; v++) sub = __fadd_rn(sub, ln[v]);
        for (; i < bs; i += 4) {
            const uint32_t w0 = __ldg(kw + b * wpb + i / 4);
#pragma unroll
            for (int e = 0; e < 4; e++) sub = __fadd_rn(sub, __fmul_rn(f[i + e], byte_f(w0, e)));
        }
        total = __fadd_rn(total, __fmul_rn(sub, __ldg(ks + b)));
    }
    return total;
}

using AttnParams = ZgKvqAttn;

// `split` of `splits` CTAs share one query column's kv range (tiles of 32 positions); with splits > 1 the CTA leaves its
// (max, sum, unnormalised accumulator) state in part[(row * splits_max + split) * (2 + d_head)] and the last CTA to arrive at
// cnt[row] merges the states in split order (deterministic) — ops.cu k_attention_fast does the same for the f32 cache.
__device__ __forceinline__ void attention_quantized_body(const AttnParams& p, const uint32_t qi, const uint32_t split, const uint32_t splits,
                                                         const uint32_t splits_max, float* __restrict__ part, uint32_t* __restrict__ cnt, const uint32_t row) {
    __shared__ float s_q[kMaxDHead];
    __shared__ uint32_t s_qi8[kMaxDHead / 4];
    __shared__ float s_qs[kMaxDHead / 4];
    __shared__ float s_acc[kAttnWarps][kMaxDHead];
    __shared__ float s_m[kAttnWarps], s_l[kAttnWarps];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t d = p.d_head, nb = p.nb, bs = p.bs;
    const float* q_col = p.q + (size_t)qi * p.q_cs;
    for (uint32_t r = threadIdx.x; r < d; r += blockDim.x) s_q[r] = q_col[r];
    __syncthreads();
    if (p.int8_query) {   // quantizeInput on the query column (src/quant.zig:976-979), a warp per block
        int8_t* qb = reinterpret_cast<int8_t*>(s_qi8);
        for (uint32_t b = warp; b < nb; b += kAttnWarps) {
            float mx = 0.0f;
            for (uint32_t k = lane; k < bs; k += 32) mx = fmaxf(mx, fabsf(s_q[b * bs + k]));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float scale = mx > 0.0f ? __fdiv_rn(mx, 127.0f) : 1.0f;
            const float inv = mx > 0.0f ? __fdiv_rn(127.0f, mx) : 0.0f;
            if (lane == 0) s_qs[b] = scale;
            for (uint32_t k = lane; k < bs; k += 32) {
                float v = __fmul_rn(s_q[b * bs + k], inv);
                v = v < -127.0f ? -127.0f : (v > 127.0f ? 127.0f : v);
                qb[b * bs + k] = (int8_t)(int)v;
            }
        }
        __syncthreads();
    }

    const size_t mask_base = (size_t)qi * p.mask_cs;
    const uint32_t wpc = d / 4;                                   // 32-bit words per column
    float m_val = -INFINITY, l = 0.0f;
    float acc[4][4];
#pragma unroll
    for (int g = 0; g < 4; g++)
#pragma unroll
        for (int e = 0; e < 4; e++) acc[g][e] = 0.0f;

    const uint32_t n_tiles = (p.seq_kv + 31) / 32;
    const uint32_t tiles_per = (n_tiles + splits - 1) / splits, t_lo = split * tiles_per, t_hi = min(t_lo + tiles_per, n_tiles);
    for (uint32_t t = t_lo + warp; t < t_hi; t += kAttnWarps) {
        const uint32_t s = t * 32 + lane;
        bool ok = s < p.seq_kv;
        float mask_add = 0.0f;
        if (ok && p.mask) { mask_add = p.mask[mask_base + (size_t)s * p.mask_rs]; ok = isfinite(mask_add); }
        float score = -INFINITY;
        if (ok) {
            const size_t c = p.k_col_start + s;
            const uint32_t* kw = reinterpret_cast<const uint32_t*>(p.k_q + c * d);
            const float* ks = p.k_s + c * nb;
            const float dot = p.int8_query ? dot_i8(kw, ks, s_qi8, s_qs, bs, nb) : dot_f32(kw, ks, s_q, bs, nb);
            score = __fadd_rn(__fmul_rn(dot, p.scale), mask_add);
            if (!isfinite(score)) { ok = false; score = -INFINITY; }
        }
        float tile_max = score;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tile_max = fmaxf(tile_max, __shfl_xor_sync(0xffffffffu, tile_max, o));
        if (tile_max == -INFINITY) continue;                      // whole tile masked (warp-uniform)
        const float new_m = fmaxf(m_val, tile_max);
        const float alpha = (m_val == -INFINITY) ? 0.0f : expf(m_val - new_m);
        const float w = ok ? expf(score - new_m) : 0.0f;
        float tile_l = w;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tile_l += __shfl_xor_sync(0xffffffffu, tile_l, o);
#pragma unroll
        for (int g = 0; g < 4; g++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[g][e] *= alpha;
        for (uint32_t pp = 0; pp < 32; pp++) {
            const float w_p = __shfl_sync(0xffffffffu, w, pp);
            if (w_p == 0.0f) continue;                            // masked, past seq_kv, or underflowed: contributes nothing
            const size_t vc = p.v_col_start + (size_t)t * 32 + pp;
            const uint32_t* vw = reinterpret_cast<const uint32_t*>(p.v_q + vc * d);
            const float* vs = p.v_s + vc * nb;
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const uint32_t wi = lane + 32 * g;
                if (wi < wpc) {
                    const uint32_t word = __ldg(vw + wi);
                    const float ws = w_p * __ldg(vs + (4 * wi) / bs);
#pragma unroll
                    for (int e = 0; e < 4; e++) acc[g][e] += ws * byte_f(word, e);
                }
            }
        }
        l = l * alpha + tile_l;
        m_val = new_m;
    }

    // merge the warps' states in warp order
    if (lane == 0) { s_m[warp] = m_val; s_l[warp] = l; }
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const uint32_t wi = lane + 32 * g;
        if (wi < wpc) {
#pragma unroll
            for (int e = 0; e < 4; e++) s_acc[warp][4 * wi + e] = acc[g][e];
        }
    }
    __syncthreads();
    float big_m = -INFINITY;
    for (uint32_t w2 = 0; w2 < kAttnWarps; w2++) big_m = fmaxf(big_m, s_m[w2]);
    float tot_l = 0.0f;
    float f[kAttnWarps];
#pragma unroll
    for (uint32_t w2 = 0; w2 < kAttnWarps; w2++) {
        f[w2] = (s_m[w2] == -INFINITY) ? 0.0f : expf(s_m[w2] - big_m);
        tot_l += s_l[w2] * f[w2];
    }
    if (splits > 1) {
        __shared__ uint32_t s_last;
        const size_t stride = 2 + d;
        float* mine = part + ((size_t)row * splits_max + split) * stride;
        for (uint32_t r = threadIdx.x; r < d; r += blockDim.x) {
            float o = 0.0f;
#pragma unroll
            for (uint32_t w2 = 0; w2 < kAttnWarps; w2++) o += s_acc[w2][r] * f[w2];
            mine[2 + r] = o;
        }
        if (threadIdx.x == 0) { mine[0] = big_m; mine[1] = tot_l; }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t old;
            asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(cnt + row) : "memory");
            const uint32_t last = (old == splits - 1) ? 1u : 0u;
            if (last) cnt[row] = 0u;   // re-arm for the next launch
            s_last = last;
        }
        __syncthreads();
        if (!s_last) return;
        const float* all = part + (size_t)row * splits_max * stride;
        float G = -INFINITY;
        for (uint32_t sp = 0; sp < splits; sp++) G = fmaxf(G, __ldcg(all + (size_t)sp * stride));
        float Lt = 0.0f;
        for (uint32_t sp = 0; sp < splits; sp++) {
            const float ms = __ldcg(all + (size_t)sp * stride);
            if (ms != -INFINITY) Lt += __ldcg(all + (size_t)sp * stride + 1) * expf(ms - G);
        }
        const float inv_L = Lt > 0.0f ? 1.0f / Lt : 0.0f;
        for (uint32_t r = threadIdx.x; r < d; r += blockDim.x) {
            float a = 0.0f;
            for (uint32_t sp = 0; sp < splits; sp++) {
                const float ms = __ldcg(all + (size_t)sp * stride);
                if (ms != -INFINITY) a += __ldcg(all + (size_t)sp * stride + 2 + r) * expf(ms - G);
            }
            p.dst[(size_t)qi * p.dst_cs + r] = a * inv_L;
        }
        return;
    }
    const float inv_l = tot_l > 0.0f ? 1.0f / tot_l : 0.0f;        // fully masked query column: zeros (src/quant.zig:1075)
    for (uint32_t r = threadIdx.x; r < d; r += blockDim.x) {
        float o = 0.0f;
#pragma unroll
        for (uint32_t w2 = 0; w2 < kAttnWarps; w2++) o += s_acc[w2][r] * f[w2];
        p.dst[(size_t)qi * p.dst_cs + r] = o * inv_l;
    }
}
__global__ void __launch_bounds__(32 * kAttnWarps) k_attention_quantized(const AttnParams p) { attention_quantized_body(p, blockIdx.x, 0, 1, 1, nullptr, nullptr, 0); }
// the cache-backed attention ops of one dependency level of a program (blockIdx.y = op); seq_kv is the op's patched value
__global__ void __launch_bounds__(32 * kAttnWarps) k_attention_quantized_tab(const AttnParams* __restrict__ tab, const uint32_t* __restrict__ dyn,
                                                                            float* __restrict__ part, uint32_t* __restrict__ cnt, const uint32_t splits_max) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    AttnParams p = tab[blockIdx.y];
    p.seq_kv = dyn[p.dyn_idx];
    // the launch provides splits_max CTAs per query column; a short context uses fewer (>= 64 positions each), the rest leave
    const uint32_t splits = min(splits_max, max(1u, (p.seq_kv + 63u) / 64u));
    if (blockIdx.z >= splits) return;
    attention_quantized_body(p, blockIdx.x, blockIdx.z, splits, splits_max, part, cnt, blockIdx.y * gridDim.x + blockIdx.x);
}

} // namespace

// ── cache-backed program ops (backend.cu zg_cuda_program_quantize_kv) ──
bool zg_kvq_cache_arrays(const ZgCudaKVCache* c, int8_t** q, float** s, uint32_t* d_head, uint32_t* bs, uint32_t* bpc, size_t* n_cols) {
    if (!c) return false;
    *q = c->q; *s = c->s; *d_head = (uint32_t)c->d_head; *bs = (uint32_t)c->bs; *bpc = (uint32_t)c->bpc; *n_cols = c->n_cols;
    return true;
}
bool zg_kvq_launch_stores(const ZgKvqStore* d_tab, uint32_t count, uint32_t max_warps, const uint32_t* d_dyn, cudaStream_t st) {
    if (count == 0 || max_warps == 0) return true;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((max_warps * 32 + 255) / 256, count); cfg.blockDim = dim3(256); cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_zg_pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_kv_store_tab, d_tab, d_dyn);
    ZG_COUNT_LAUNCH();
    if (e != cudaSuccess) { zg_set_error("quantized KV store launch failed: %s", cudaGetErrorString(e)); return false; }
    return true;
}
bool zg_kvq_launch_attention(const ZgKvqAttn* d_tab, uint32_t count, uint32_t seq_q, const uint32_t* d_dyn, float* part, uint32_t* cnt,
                             uint32_t splits_max, cudaStream_t st) {
    if (count == 0 || seq_q == 0) return true;
    if (!part || !cnt || splits_max < 1) splits_max = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(seq_q, count, splits_max); cfg.blockDim = dim3(32 * kAttnWarps); cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_zg_pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_attention_quantized_tab, d_tab, d_dyn, part, cnt, splits_max);
    ZG_COUNT_LAUNCH();
    if (e != cudaSuccess) { zg_set_error("quantized attention launch failed: %s", cudaGetErrorString(e)); return false; }
    return true;
}

extern "C" ZgCudaKVCache* zg_cuda_kvcache_create(ZgCudaCtx* ctx, size_t d_head, size_t n_cols, size_t block_size) {
    if (!ctx || d_head == 0 || n_cols == 0 || block_size == 0 || d_head % block_size != 0) {
        zg_set_error("kvcache_create: d_head must be a positive multiple of block_size (src/quant.zig:659)");
        return nullptr;
    }
    if (d_head > kMaxDHead || d_head % 4 != 0 || block_size % 4 != 0) {
        zg_set_error("kvcache_create: needs d_head <= 512 (src/quant.zig:944-947), d_head %% 4 == 0 and block_size %% 4 == 0");
        return nullptr;
    }
    cudaSetDevice(ctx->device);
    ZgCudaKVCache* c = new ZgCudaKVCache();
    c->d_head = d_head; c->n_cols = n_cols; c->bs = block_size; c->bpc = d_head / block_size;
    if (cudaMalloc(&c->q, d_head * n_cols) != cudaSuccess || cudaMalloc(&c->s, c->bpc * n_cols * sizeof(float)) != cudaSuccess) {
        zg_set_error("kvcache_create: cudaMalloc failed");
        cudaFree(c->q); cudaFree(c->s); delete c;
        return nullptr;
    }
    cudaMemsetAsync(c->q, 0, d_head * n_cols, ctx->stream);
    cudaMemsetAsync(c->s, 0, c->bpc * n_cols * sizeof(float), ctx->stream);
    return c;
}

extern "C" void zg_cuda_kvcache_free(ZgCudaCtx* ctx, ZgCudaKVCache* c) {
    if (!c) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    cudaFree(c->q); cudaFree(c->s);
    delete c;
}

extern "C" int zg_cuda_kvcache_clear(ZgCudaCtx* ctx, ZgCudaKVCache* c) {
    if (!ctx || !c) { zg_set_error("kvcache_clear: bad arguments"); return -1; }
    cudaSetDevice(ctx->device);
    cudaMemsetAsync(c->q, 0, c->d_head * c->n_cols, ctx->stream);
    cudaMemsetAsync(c->s, 0, c->bpc * c->n_cols * sizeof(float), ctx->stream);
    return 0;
}

extern "C" int zg_cuda_kvcache_store_device(ZgCudaCtx* ctx, ZgCudaKVCache* c, size_t col_start, size_t n_write, const float* d_src) {
    if (!ctx || !c || !d_src) { zg_set_error("kvcache_store: bad arguments"); return -1; }
    if (col_start + n_write > c->n_cols) { zg_set_error("kvcache_store: columns %zu..%zu outside the cache (%zu)", col_start, col_start + n_write, c->n_cols); return -1; }
    if (n_write == 0) return 0;
    cudaSetDevice(ctx->device);
    const size_t warps = n_write * c->bpc;
    k_kv_store<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, ctx->stream>>>(d_src, (uint32_t)c->d_head, (uint32_t)c->bs, (uint32_t)c->bpc,
                                                                            (uint32_t)n_write, c->q, c->s, col_start);
    ZG_COUNT_LAUNCH();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { zg_set_error("kvcache_store: launch failed: %s", cudaGetErrorString(e)); return -1; }
    return 0;
}

extern "C" int zg_cuda_kvcache_store_host(ZgCudaCtx* ctx, ZgCudaKVCache* c, size_t col_start, size_t n_write, const float* h_src) {
    if (!ctx || !c || !h_src) { zg_set_error("kvcache_store_host: bad arguments"); return -1; }
    if (n_write == 0) return 0;
    cudaSetDevice(ctx->device);
    float* d_src = nullptr;
    const size_t bytes = n_write * c->d_head * sizeof(float);
    if (cudaMalloc(&d_src, bytes) != cudaSuccess) { zg_set_error("kvcache_store_host: cudaMalloc failed"); return -1; }
    cudaMemcpyAsync(d_src, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream);
    int rc = zg_cuda_kvcache_store_device(ctx, c, col_start, n_write, d_src);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && rc == 0) { zg_set_error("kvcache_store_host: kernel failed"); rc = -1; }
    cudaFree(d_src);
    return rc;
}

extern "C" int zg_cuda_kvcache_download(ZgCudaCtx* ctx, const ZgCudaKVCache* c, int8_t* h_q, float* h_scales) {
    if (!ctx || !c) { zg_set_error("kvcache_download: bad arguments"); return -1; }
    cudaSetDevice(ctx->device);
    if (h_q) cudaMemcpyAsync(h_q, c->q, c->d_head * c->n_cols, cudaMemcpyDeviceToHost, ctx->stream);
    if (h_scales) cudaMemcpyAsync(h_scales, c->s, c->bpc * c->n_cols * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { zg_set_error("kvcache_download: copy failed"); return -1; }
    return 0;
}

extern "C" int zg_cuda_attention_quantized_device(ZgCudaCtx* ctx, float* d_dst, size_t dst_col_stride, const float* d_q, size_t q_col_stride,
                                                  size_t d_head, size_t seq_q, const ZgCudaKVCache* k_cache, size_t k_col_start,
                                                  const ZgCudaKVCache* v_cache, size_t v_col_start, size_t seq_kv, const float* d_mask,
                                                  size_t mask_row_stride, size_t mask_col_stride, float scale, int int8_query) {
    if (!ctx || !d_dst || !d_q || !k_cache || !v_cache) { zg_set_error("attention_quantized: bad arguments"); return -1; }
    if (k_cache->d_head != d_head || v_cache->d_head != d_head || k_cache->bs != v_cache->bs) {
        zg_set_error("attention_quantized: cache d_head / block size mismatch (src/quant.zig:941)"); return -1;
    }
    if (k_col_start + seq_kv > k_cache->n_cols || v_col_start + seq_kv > v_cache->n_cols) {
        zg_set_error("attention_quantized: kv range outside the cache (src/quant.zig:942-943)"); return -1;
    }
    if (int8_query && k_cache->bpc > 32) { zg_set_error("attention_quantized: int8 query needs <= 32 blocks per column (src/quant.zig:945)"); return -1; }
    if (seq_q == 0) return 0;
    cudaSetDevice(ctx->device);
    AttnParams p;
    p.dst = d_dst; p.dst_cs = dst_col_stride; p.q = d_q; p.q_cs = q_col_stride;
    p.d_head = (uint32_t)d_head; p.seq_kv = (uint32_t)seq_kv; p.bs = (uint32_t)k_cache->bs; p.nb = (uint32_t)k_cache->bpc;
    p.k_q = k_cache->q; p.k_s = k_cache->s; p.k_col_start = k_col_start;
    p.v_q = v_cache->q; p.v_s = v_cache->s; p.v_col_start = v_col_start;
    p.mask = d_mask; p.mask_rs = mask_row_stride; p.mask_cs = mask_col_stride;
    p.scale = scale; p.int8_query = int8_query ? 1 : 0; p.dyn_idx = 0;
    k_attention_quantized<<<(unsigned)seq_q, 32 * kAttnWarps, 0, ctx->stream>>>(p);
    ZG_COUNT_LAUNCH();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { zg_set_error("attention_quantized: launch failed: %s", cudaGetErrorString(e)); return -1; }
    return 0;
}

extern "C" int zg_cuda_attention_quantized_host(ZgCudaCtx* ctx, float* h_dst, size_t dst_col_stride, const float* h_q, size_t q_col_stride,
                                                size_t d_head, size_t seq_q, const ZgCudaKVCache* k_cache, size_t k_col_start,
                                                const ZgCudaKVCache* v_cache, size_t v_col_start, size_t seq_kv, const float* h_mask,
                                                size_t mask_row_stride, size_t mask_col_stride, float scale, int int8_query) {
    if (!ctx || !h_dst || !h_q) { zg_set_error("attention_quantized_host: bad arguments"); return -1; }
    if (seq_q == 0) return 0;
    cudaSetDevice(ctx->device);
    const size_t q_n = (seq_q - 1) * q_col_stride + d_head, dst_n = (seq_q - 1) * dst_col_stride + d_head;
    const size_t mask_n = h_mask ? ((seq_kv ? seq_kv - 1 : 0) * mask_row_stride + (seq_q - 1) * mask_col_stride + 1) : 0;
    float* d_q = nullptr; float* d_dst = nullptr; float* d_mask = nullptr;
    if (cudaMalloc(&d_q, q_n * sizeof(float)) != cudaSuccess || cudaMalloc(&d_dst, dst_n * sizeof(float)) != cudaSuccess ||
        (mask_n && cudaMalloc(&d_mask, mask_n * sizeof(float)) != cudaSuccess)) {
        zg_set_error("attention_quantized_host: cudaMalloc failed");
        cudaFree(d_q); cudaFree(d_dst); cudaFree(d_mask);
        return -1;
    }
    cudaMemcpyAsync(d_q, h_q, q_n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_dst, h_dst, dst_n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);   // cells between columns keep their contents
    if (mask_n) cudaMemcpyAsync(d_mask, h_mask, mask_n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
    int rc = zg_cuda_attention_quantized_device(ctx, d_dst, dst_col_stride, d_q, q_col_stride, d_head, seq_q, k_cache, k_col_start, v_cache,
                                                v_col_start, seq_kv, d_mask, mask_row_stride, mask_col_stride, scale, int8_query);
    if (rc == 0) cudaMemcpyAsync(h_dst, d_dst, dst_n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && rc == 0) { zg_set_error("attention_quantized_host: kernel failed: %s", cudaGetErrorString(cudaGetLastError())); rc = -1; }
    cudaFree(d_q); cudaFree(d_dst); cudaFree(d_mask);
    return rc;
}
