// Dense LM-head matvec with 16-bit weights (f16 or bf16) and an argmax-safe bound (opt-in, zg_cuda_program_promote_dense).
//
// The tied LM head of the SmolLM / Llama models is a dense f32 matmul of one activation row against token_embed^T
// (src/models/llama.zig:162-165): [vocab, d_model] f32, the largest single byte stream of a SmolLM decode step.  The
// reference's WGPU backend promotes exactly these operands — matmul B buffers that come from an initial upload — to f16,
// unconditionally and lossily (src/backend/wgpu.zig:1068-1106).  Here the promotion is opt-in and keeps the argmax:
//
//   load time   w^[n, :] = f16 / bf16 (w[n, :]) (round to nearest even), err[n] = ||w[n, :] - w^[n, :]||_2 (+ accumulation slack)
//   k_head_16     y^[n] = sum_k x[k] w^[n, k] (fp32 accumulate)  -> dst[n];  |y[n] - y^[n]| <= ||x||_2 err[n] =: b[n]
//                 (Cauchy-Schwarz on the rounding error vector); per CTA (8 columns) the largest lower bound y^ - b and the
//                 largest upper bound y^ + b go to scratch.  The CTA that arrives last (atomic counter) takes
//                 L = max lower bound <= max_n y[n]; every CTA whose upper bound reaches L may hold the argmax: its 8 columns
//                 are recomputed from the f32 rows with the exact kernel's arithmetic (k_matmul_kmajor) and overwrite
//                 dst.  Every other column has y[n] <= y^[n] + b[n] < L <= max y: it cannot win.  One launch.
//
// So argmax(dst) is the f32 path's argmax and the winning logits are the f32 path's bits; the other logits carry the 16-bit
// rounding of the weights.  Two formats: f16 (what the reference's WGPU backend uses; 11 significant bits: measured <= 2e-4 of
// the output scale, inside the 1e-3 budget of whole-program logits; |w| > 65504 becomes inf and those columns are simply
// always recomputed) and bf16 (f32's range, 8 significant bits: 3e-4 ... 2e-3 of the output scale, over the budget on some
// models — argmax-safe all the same).
// HBM bytes per token: 2 N K + 12 N instead of 4 N K.
#include "zg_internal.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace {

__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// one warp per weight row: bf16 copy + the norm of the rounding error
template <int FMT>
__global__ void k_dense_to_16(const float* __restrict__ w, size_t w_off, size_t row_stride, uint32_t N, uint32_t K,
                              uint16_t* __restrict__ out, float* __restrict__ err) {
    const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (n >= N) return;
    const float* row = w + w_off + (size_t)n * row_stride;
    float e2 = 0.0f, w2 = 0.0f;
    for (uint32_t k = lane; k < K; k += 32) {
        const float v = row[k];
        float back;
        if constexpr (FMT == ZG_DENSE_BF16) { const __nv_bfloat16 h = __float2bfloat16_rn(v); out[(size_t)n * K + k] = __bfloat16_as_ushort(h); back = __bfloat162float(h); }
        else { const __half h = __float2half_rn(v); out[(size_t)n * K + k] = __half_as_ushort(h); back = __half2float(h); }
        const float d = v - back;   // |w| beyond the f16 range: inf -> err = inf -> the column is always recomputed exactly
        e2 = fmaf(d, d, e2);
        w2 = fmaf(v, v, w2);
    }
    e2 = warp_sum(e2); w2 = warp_sum(w2);
    // rounding-error norm, inflated for its own fp32 evaluation, plus slack for the fp32 accumulation of BOTH dots
    // (lane-strided sums of K / 32 terms + a 5-step tree: each within ~(K / 32 + 8) 2^-23 of sum |x_k w_k| <= ||x|| ||w||)
    if (lane == 0) err[n] = sqrtf(e2) * 1.001f + sqrtf(w2) * (float)(K / 32 + 8) * 1.2e-7f;
}

// exact dot of column n with k_matmul_kmajor's arithmetic (same lane stride, same fma order), by one warp
__device__ __forceinline__ float exact_dot(const float* __restrict__ x, const float* __restrict__ row, size_t row_stride, size_t w_off, uint32_t K, uint32_t lane) {
    float acc = 0.0f;
    if (((row_stride & 3) == 0) && ((w_off & 3) == 0) && ((K & 3) == 0)) {
        const float4* a4 = reinterpret_cast<const float4*>(x);
        const float4* b4 = reinterpret_cast<const float4*>(row);
        for (uint32_t k = lane; k < K / 4; k += 32) {
            const float4 a = a4[k], b = __ldcs(b4 + k);
            acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc);
            acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
        }
    } else {
        for (uint32_t k = lane; k < K; k += 32) acc = fmaf(x[k], row[k], acc);
    }
    return warp_sum(acc);
}

struct HeadParams {
    const float* x; const uint16_t* w16; const float* err; const float* w; size_t w_off, row_stride;
    uint32_t N, K, n_part;
    float* dst; float* part_lo; float* part_hi; uint32_t* counter;
};

// ONE launch.  The activations are staged once per CTA in shared memory as two float4 arrays (XA[k8] = x[8 k8 .. +3],
// XB[k8] = x[8 k8 + 4 .. +7]: consecutive lanes read consecutive 16 bytes, conflict-free); a warp owns 8 consecutive columns
// and walks them one after the other with up to 8 16-byte weight loads in flight per lane.  Per warp the largest lower bound
// y^ - b and the largest upper bound y^ + b of its 8 columns go to scratch; the CTA that arrives last reduces them, finds the
// 8-column groups that could still hold the maximum and recomputes those columns exactly.
template <int FMT, int U, int COLS>
__global__ void __launch_bounds__(256)
k_head_16(const HeadParams p) {
    pdl_enter();
    extern __shared__ __align__(16) float4 s_x[];    // XA[K8] | XB[K8]
    __shared__ float s_red[8];
    __shared__ uint32_t s_last, s_ncand, s_cand[1024];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t K = p.K, K8 = K / 8, grp = blockIdx.x * 8 + warp;   // group of 8 columns
    float xx = 0.0f;
    {
        const float4* x4 = reinterpret_cast<const float4*>(p.x);
        for (uint32_t i = threadIdx.x; i < 2 * K8; i += 256) {
            const float4 v = x4[i];
            s_x[(i & 1) * K8 + (i >> 1)] = v;
            xx = fmaf(v.x, v.x, xx); xx = fmaf(v.y, v.y, xx); xx = fmaf(v.z, v.z, xx); xx = fmaf(v.w, v.w, xx);
        }
        xx = warp_sum(xx);
        if (lane == 0) s_red[warp] = xx;
        __syncthreads();
        xx = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) xx += s_red[i];
        __syncthreads();
    }
    const float xn = sqrtf(xx) * 1.001f;
    const float4* XA = s_x;
    const float4* XB = s_x + K8;
    float lo = -INFINITY, hi = -INFINITY;
    if (grp * 8 < p.N) {
#pragma unroll 1
        for (uint32_t c0 = 0; c0 < 8; c0 += COLS) {     // COLS columns at a time: short rows need several columns' loads in flight
            float acc[COLS];
#pragma unroll
            for (int cc = 0; cc < COLS; cc++) acc[cc] = 0.0f;
            for (uint32_t base = 0; base < K8; base += 32 * U) {
                uint4 q[COLS][U];
#pragma unroll
                for (int cc = 0; cc < COLS; cc++) {
                    const uint32_t n = min(grp * 8 + c0 + cc, p.N - 1);
                    const uint4* wr = reinterpret_cast<const uint4*>(p.w16 + (size_t)n * K);
#pragma unroll
                    for (int u = 0; u < U; u++) { const uint32_t k8 = base + u * 32 + lane; q[cc][u] = k8 < K8 ? __ldcs(wr + k8) : make_uint4(0, 0, 0, 0); }
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const uint32_t k8 = base + u * 32 + lane;
                    if (k8 < K8) {
                        const float4 a = XA[k8], b = XB[k8];
#pragma unroll
                        for (int cc = 0; cc < COLS; cc++) {
                            const uint4 qq = q[cc][u];
                            float2 w0, w1, w2, w3;
                            if constexpr (FMT == ZG_DENSE_BF16) {
                                w0 = make_float2(__uint_as_float(qq.x << 16), __uint_as_float(qq.x & 0xFFFF0000u)); w1 = make_float2(__uint_as_float(qq.y << 16), __uint_as_float(qq.y & 0xFFFF0000u));
                                w2 = make_float2(__uint_as_float(qq.z << 16), __uint_as_float(qq.z & 0xFFFF0000u)); w3 = make_float2(__uint_as_float(qq.w << 16), __uint_as_float(qq.w & 0xFFFF0000u));
                            } else {
                                w0 = __half22float2(*reinterpret_cast<const __half2*>(&qq.x)); w1 = __half22float2(*reinterpret_cast<const __half2*>(&qq.y));
                                w2 = __half22float2(*reinterpret_cast<const __half2*>(&qq.z)); w3 = __half22float2(*reinterpret_cast<const __half2*>(&qq.w));
                            }
                            float t = acc[cc];
                            t = fmaf(a.x, w0.x, t); t = fmaf(a.y, w0.y, t); t = fmaf(a.z, w1.x, t); t = fmaf(a.w, w1.y, t);
                            t = fmaf(b.x, w2.x, t); t = fmaf(b.y, w2.y, t); t = fmaf(b.z, w3.x, t); t = fmaf(b.w, w3.y, t);
                            acc[cc] = t;
                        }
                    }
                }
            }
#pragma unroll
            for (int cc = 0; cc < COLS; cc++) {
                const uint32_t n = grp * 8 + c0 + cc;
                const float v = warp_sum(acc[cc]);
                if (n < p.N) {
                    if (lane == 0) p.dst[n] = v;
                    const float bnd = xn * p.err[n];
                    float l2 = v - bnd, h2 = v + bnd;   // non-finite activations / weights: the group is marked and recomputed
                    if (!(bnd == bnd) || !(v == v) || fabsf(bnd) > 3.0e38f || fabsf(v) > 3.0e38f) { l2 = -INFINITY; h2 = INFINITY; }
                    lo = fmaxf(lo, l2); hi = fmaxf(hi, h2);
                }
            }
        }
        if (lane == 0) { p.part_lo[grp] = lo; p.part_hi[grp] = hi; }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t old = atomicAdd(p.counter, 1u);
        s_last = (old == gridDim.x - 1) ? 1u : 0u;
        if (s_last) { *p.counter = 0u; s_ncand = 0u; }   // re-arm for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // L = max of the lower bounds <= max_n y[n]
    float L = -INFINITY;
    for (uint32_t i = threadIdx.x; i < p.n_part; i += 256) L = fmaxf(L, __ldcg(p.part_lo + i));
    L = warp_max(L);
    if (lane == 0) s_red[warp] = L;
    __syncthreads();
    L = s_red[0];
#pragma unroll
    for (int i = 1; i < 8; i++) L = fmaxf(L, s_red[i]);
    // groups with a column whose upper bound reaches L: their 8 columns are recomputed exactly (a superset of the candidates)
    for (uint32_t i = threadIdx.x; i < p.n_part; i += 256)
        if (!(__ldcg(p.part_hi + i) < L)) { const uint32_t s2 = atomicAdd(&s_ncand, 1u); if (s2 < 1024) s_cand[s2] = i; }
    __syncthreads();
    const uint32_t nc = s_ncand;
    if (nc <= 1024) {
        for (uint32_t c = 0; c < nc; c++) {
            const uint32_t col = s_cand[c] * 8 + warp;
            if (col < p.N) {
                const float v = exact_dot(p.x, p.w + p.w_off + (size_t)col * p.row_stride, p.row_stride, p.w_off, K, lane);
                if (lane == 0) p.dst[col] = v;
            }
        }
    } else {   // nothing could be excluded (degenerate or non-finite inputs): the exact result everywhere
        for (uint32_t col = warp; col < p.N; col += 8) {
            const float v = exact_dot(p.x, p.w + p.w_off + (size_t)col * p.row_stride, p.row_stride, p.w_off, K, lane);
            if (lane == 0) p.dst[col] = v;
        }
    }
}

template <int FMT>
cudaError_t launch_head(const HeadParams& p, uint32_t grid, cudaStream_t st, bool pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = p.K * 4; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    if (p.K <= 768) return cudaLaunchKernelEx(&cfg, k_head_16<FMT, 3, 4>, p);    // short rows (SmolLM-135M: 72 loads per row): 4 columns at a time
    if (p.K <= 1024) return cudaLaunchKernelEx(&cfg, k_head_16<FMT, 4, 2>, p);
    return cudaLaunchKernelEx(&cfg, k_head_16<FMT, 8, 1>, p);
}

}  // namespace

ZgDenseHead* zg_dense_head_create(ZgCudaCtx* ctx, int format, const float* d_w, size_t w_off, size_t row_stride, uint32_t N, uint32_t K) {
    ZgDenseHead* h = new ZgDenseHead();
    h->format = format; h->N = N; h->K = K; h->n_part = (N + 7) / 8;
    if (cudaMalloc(&h->w16, (size_t)N * K * 2) != cudaSuccess || cudaMalloc(&h->err, (size_t)N * 4) != cudaSuccess ||
        cudaMalloc(&h->part_max, (size_t)h->n_part * 8) != cudaSuccess || cudaMalloc(&h->glob, 8) != cudaSuccess ||
        cudaMemset(h->glob, 0, 8) != cudaSuccess) {
        zg_set_error("promote_dense: out of device memory for a %u x %u 16-bit copy", N, K);
        zg_dense_head_free(h);
        return nullptr;
    }
    if (!zg_dense_head_refresh(ctx, h, d_w, w_off, row_stride, ctx->stream)) { zg_dense_head_free(h); return nullptr; }
    return h;
}

bool zg_dense_head_refresh(ZgCudaCtx*, ZgDenseHead* h, const float* d_w, size_t w_off, size_t row_stride, cudaStream_t st) {
    if (h->format == ZG_DENSE_BF16) k_dense_to_16<ZG_DENSE_BF16><<<(h->N * 32 + 255) / 256, 256, 0, st>>>(d_w, w_off, row_stride, h->N, h->K, (uint16_t*)h->w16, h->err);
    else k_dense_to_16<ZG_DENSE_F16><<<(h->N * 32 + 255) / 256, 256, 0, st>>>(d_w, w_off, row_stride, h->N, h->K, (uint16_t*)h->w16, h->err);
    ZG_COUNT_LAUNCH();
    if (cudaGetLastError() != cudaSuccess) { zg_set_error("promote_dense: conversion launch failed"); return false; }
    return true;
}

void zg_dense_head_free(ZgDenseHead* h) {
    if (!h) return;
    cudaFree(h->w16); cudaFree(h->err); cudaFree(h->part_max); cudaFree(h->glob);
    delete h;
}

// dst[0..N) = x[0..K) . w[n, :] with the 16-bit copy + exact recompute of the argmax candidates, ONE launch
bool zg_dense_head_launch(ZgCudaCtx* ctx, const ZgDenseHead* h, const float* d_x, const float* d_w, size_t w_off, size_t row_stride,
                          float* d_dst, cudaStream_t st) {
    HeadParams p;
    p.x = d_x; p.w16 = (const uint16_t*)h->w16; p.err = h->err; p.w = d_w; p.w_off = w_off; p.row_stride = row_stride;
    p.N = h->N; p.K = h->K; p.n_part = h->n_part;
    p.dst = d_dst; p.part_lo = h->part_max; p.part_hi = h->part_max + h->n_part; p.counter = (uint32_t*)h->glob;
    const uint32_t grid = (h->n_part + 7) / 8;   // a CTA = 8 warps = 8 groups of 8 columns
    const cudaError_t e = h->format == ZG_DENSE_BF16 ? launch_head<ZG_DENSE_BF16>(p, grid, st, ctx->pdl) : launch_head<ZG_DENSE_F16>(p, grid, st, ctx->pdl);
    ZG_COUNT_LAUNCH();
    if (e != cudaSuccess) { zg_set_error("16-bit head launch failed: %s", cudaGetErrorString(e)); return false; }
    return true;
}
