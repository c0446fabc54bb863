// W8A8 decode path of zgml's QuantizedWeight for sm_100a: prepareTransposed (src/quant.zig:274-317, twin
// src/backend/reference.zig:26-70), quantizeInput (src/quant.zig:320-341) and gemv = quantizeInput + gemvRange
// (src/quant.zig:358-459) — the loop `session.quantize()` models run per linear layer on the reference's aarch64 builds.
//
// Everything here is integer / IEEE-exact work, so the results are BIT-IDENTICAL to the reference's:
//   * transposed weights: dequantize [K, N] (f32(q) * scale), re-quantize per output row n and K-aligned block with the
//     truncating rule (scale = max_abs / 127, q = trunc(clamp(v * (127 / max_abs), +-127))) — separate IEEE operations.
//   * activations: the same rule per K-block.
//   * gemv: per (n, block) an exact int32 dot product (dp4a; the reference's sdot lanes are order-free), then
//     acc += f32(dot) * (s_x[b] * s_w[n, b]) as separate multiply / add, blocks ASCENDING per output — the float
//     accumulation order of gemvRange.  A warp owns one output row: lanes stream the int8 row with 128-bit loads, eight
//     in flight per lane (one row = one contiguous run of K bytes + K/bs scales: HBM-bound, 1.125 B per weight at
//     bs = 32; the quantized activations come through L1), block terms meet in shared memory and one lane adds them in
//     block order while the SM's other warps stream.
#include "zg_internal.cuh"

#include <stdlib.h>

namespace {

constexpr uint32_t kGemvWarps = 8;
constexpr uint32_t kMaxK = 16384, kMaxBlocks = 512;   // the reference's stack buffers (src/quant.zig:452-453)

// truncating block quantization shared by prepareTransposed and quantizeInput
__device__ __forceinline__ void block_scale(float max_abs, float* scale, float* inv) {
    *scale = max_abs > 0.0f ? __fdiv_rn(max_abs, 127.0f) : 1.0f;
    *inv = max_abs > 0.0f ? __fdiv_rn(127.0f, max_abs) : 0.0f;
}
__device__ __forceinline__ int8_t quant_one(float v, float inv) {
    float q = __fmul_rn(v, inv);
    q = q < -127.0f ? -127.0f : (q > 127.0f ? 127.0f : q);
    return (int8_t)(int)q;   // float -> int truncates toward zero, like @intFromFloat
}

// One warp per K-block of the activation vector.
__global__ void k_quantize_input(const float* __restrict__ x, uint32_t K, uint32_t bs, int8_t* __restrict__ xq, float* __restrict__ xs) {
    // the gemv that follows is launched programmatically dependent: it may become resident now and stream its weight
    // rows while this kernel runs; it reads xq / xs only after its griddepcontrol.wait
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t start = b * bs;
    if (start >= K) return;
    const uint32_t end = min(start + bs, K);
    float mx = 0.0f;
    for (uint32_t k = start + lane; k < end; k += 32) mx = fmaxf(mx, fabsf(x[k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float scale, inv;
    block_scale(mx, &scale, &inv);
    if (lane == 0) xs[b] = scale;
    for (uint32_t k = start + lane; k < end; k += 32) xq[k] = quant_one(x[k], inv);
}

// Thread (n, b): the block's values deq[k, n] sit N floats apart, so a warp (32 consecutive n) reads 128 contiguous
// bytes per k.  Load-time only.
__global__ void k_transpose_requant(const float* __restrict__ deq, uint32_t K, uint32_t N, uint32_t bs, uint32_t bpr,
                                    int8_t* __restrict__ t_data, float* __restrict__ t_scales) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    for (uint32_t b = blockIdx.y; b < bpr; b += gridDim.y) {
        const uint32_t k0 = b * bs, k1 = min(k0 + bs, K);
        float mx = 0.0f;
        for (uint32_t k = k0; k < k1; k++) mx = fmaxf(mx, fabsf(deq[(size_t)k * N + n]));
        float scale, inv;
        block_scale(mx, &scale, &inv);
        t_scales[(size_t)n * bpr + b] = scale;
        int8_t* row = t_data + (size_t)n * K;
        for (uint32_t k = k0; k < k1; k++) row[k] = quant_one(deq[(size_t)k * N + n], inv);
    }
}

__device__ __forceinline__ int dot16(const uint4& a, const uint4& b) {
    int d = __dp4a((int)a.x, (int)b.x, 0);
    d = __dp4a((int)a.y, (int)b.y, d);
    d = __dp4a((int)a.z, (int)b.z, d);
    return __dp4a((int)a.w, (int)b.w, d);
}

// Fast path: K % 16 == 0, bs = 16 * GS with GS a power of two <= 32.  GS consecutive lanes hold one block.
template <int GS>
__global__ void __launch_bounds__(32 * kGemvWarps, 4)
k_gemv_w8a8(const int8_t* __restrict__ t_d, const float* __restrict__ t_s, const int8_t* __restrict__ xq,
            const float* __restrict__ xs, float* __restrict__ dst, uint32_t N, uint32_t K, uint32_t bpr) {
    extern __shared__ float terms[];                                         // [warp][bpr]
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t n = blockIdx.x * kGemvWarps + warp;
    if (n >= N) return;
    const uint32_t chunks = K / 16;
    float* my_terms = terms + (size_t)warp * bpr;
    const uint4* row = reinterpret_cast<const uint4*>(t_d + (size_t)n * K);
    const uint4* x4 = reinterpret_cast<const uint4*>(xq);
    const float* srow = t_s + (size_t)n * bpr;
    const bool leader = (lane & (GS - 1)) == 0;
    constexpr int U = 8;                                                     // 128-bit weight loads in flight per lane
    for (uint32_t c0 = 0; c0 < chunks; c0 += 32 * U) {
        uint4 w[U];
        float sw[U];
#pragma unroll
        for (int u = 0; u < U; u++) {                                        // immutable weights + their scales first
            const uint32_t c = c0 + u * 32 + lane;
            w[u] = c < chunks ? __ldg(row + c) : make_uint4(0, 0, 0, 0);
            sw[u] = (c < chunks && leader) ? __ldg(srow + c / GS) : 0.0f;
        }
        // quantized activations: written by the kernel before this one — complete and visible only after the wait
        if (c0 == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint32_t c = c0 + u * 32 + lane;
            const uint4 xv = c < chunks ? __ldg(x4 + c) : make_uint4(0, 0, 0, 0);
            int d = dot16(w[u], xv);
#pragma unroll
            for (int o = 1; o < GS; o <<= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
            if (c < chunks && leader) {
                const uint32_t b = c / GS;
                my_terms[b] = __fmul_rn((float)d, __fmul_rn(__ldg(xs + b), sw[u]));   // f32(int) * (s_x[b] * s_w[n, b])
            }
        }
    }
    __syncwarp();
    if (lane == 0) {
        float acc = 0.0f;
        for (uint32_t b = 0; b < bpr; b++) acc = __fadd_rn(acc, my_terms[b]);   // blocks ascending, like gemvRange
        dst[n] = acc;
    }
}

// Measured and removed (round 2): a single-kernel form that quantizes x inside every CTA (ZG_W8A8_FUSED) was bit-identical
// but ran at 1.39 / 1.88 TB/s against 2.23 / 3.61 TB/s for the two-kernel form below (every CTA repeats the activation
// quantization, and its 16 dependent rounds sit in front of the first weight byte).

// Measured and rejected (round 2, same-box A/B): two output rows per warp (half as many CTAs per gemv so that the next gemv's
// CTAs fit beside it) with the quantize kernel also launched programmatically dependent: 2.56 / 3.47 TB/s at 4096 x 4096 /
// 4096 x 14336 against 2.23 / 3.61 for this form — a wash; the pair of dependent launches per gemv (about 4 us of latency
// for 2.9 us of HBM time at 4096 x 4096) bounds it either way.

// Any block size / K: one thread per output, the reference loop as written.
__global__ void k_gemv_w8a8_generic(const int8_t* __restrict__ t_d, const float* __restrict__ t_s, const int8_t* __restrict__ xq,
                                    const float* __restrict__ xs, float* __restrict__ dst, uint32_t N, uint32_t K, uint32_t bs,
                                    uint32_t bpr) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float acc = 0.0f;
    for (uint32_t b = 0; b < bpr; b++) {
        const uint32_t k0 = b * bs, k1 = min(k0 + bs, K);
        int ia = 0;
        for (uint32_t k = k0; k < k1; k++) ia += (int)xq[k] * (int)t_d[(size_t)n * K + k];
        acc = __fadd_rn(acc, __fmul_rn((float)ia, __fmul_rn(xs[b], t_s[(size_t)n * bpr + b])));
    }
    dst[n] = acc;
}

bool check_limits(size_t K, size_t bs, const char* who) {
    if (bs == 0 || K == 0) { zg_set_error("%s: empty input", who); return false; }
    if (K > kMaxK || (K + bs - 1) / bs > kMaxBlocks) {
        zg_set_error("%s: K = %zu / %zu blocks exceed the reference's limits (K <= 16384, <= 512 blocks; src/quant.zig:452-453)", who, K, (K + bs - 1) / bs);
        return false;
    }
    return true;
}

bool launch_quantize(const float* d_x, size_t K, size_t bs, int8_t* d_q, float* d_s, cudaStream_t st) {
    const uint32_t bpr = (uint32_t)((K + bs - 1) / bs);
    k_quantize_input<<<(bpr * 32 + 255) / 256, 256, 0, st>>>(d_x, (uint32_t)K, (uint32_t)bs, d_q, d_s);
    ZG_COUNT_LAUNCH();
    return cudaGetLastError() == cudaSuccess;
}

// Programmatically dependent on the quantizeInput kernel launched just before it on the same stream.
template <typename... KArgs>
bool launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, KArgs... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
    ZG_COUNT_LAUNCH();
    return e == cudaSuccess;
}

template <int GS>
bool launch_fast(const ZgCudaQWeight* w, float* d_dst, uint32_t bpr, size_t smem, cudaStream_t st) {
    return launch_pdl(k_gemv_w8a8<GS>, (unsigned)((w->N + kGemvWarps - 1) / kGemvWarps), 32 * kGemvWarps, smem, st,
                      (const int8_t*)w->t_data, (const float*)w->t_scales, (const int8_t*)w->x_q, (const float*)w->x_s, d_dst,
                      (uint32_t)w->N, (uint32_t)w->K, bpr);
}

} // namespace

extern "C" int zg_cuda_qweight_prepare_transposed(ZgCudaCtx* ctx, ZgCudaQWeight* w, int8_t* h_t_data, float* h_t_scales) {
    if (!ctx || !w) { zg_set_error("prepare_transposed: bad arguments"); return -1; }
    if (w->K == 0 || w->N == 0) { zg_set_error("prepare_transposed: empty weight"); return -1; }
    if (w->K > 0xFFFFFFFFull || w->N > 0xFFFFFFFFull || w->bs > 0xFFFFFFFFull) { zg_set_error("prepare_transposed: dimensions out of range"); return -1; }
    cudaSetDevice(ctx->device);
    const size_t K = w->K, N = w->N, bs = w->bs, bpr = (K + bs - 1) / bs;
    if (!w->t_data) {
        float* d_deq = nullptr;
        int8_t* t_data = nullptr; float* t_scales = nullptr; int8_t* x_q = nullptr; float* x_s = nullptr;
        if (cudaMalloc(&d_deq, K * N * sizeof(float)) != cudaSuccess || cudaMalloc(&t_data, N * K) != cudaSuccess ||
            cudaMalloc(&t_scales, N * bpr * sizeof(float)) != cudaSuccess || cudaMalloc(&x_q, (K + 15) / 16 * 16) != cudaSuccess ||
            cudaMalloc(&x_s, bpr * sizeof(float)) != cudaSuccess) {
            zg_set_error("prepare_transposed: cudaMalloc failed");
            cudaFree(d_deq); cudaFree(t_data); cudaFree(t_scales); cudaFree(x_q); cudaFree(x_s);
            return -1;
        }
        bool ok = zg_qweight_dequant_to_device(ctx, w, d_deq);
        if (ok) {
            dim3 grid((unsigned)((N + 127) / 128), (unsigned)(bpr < 1024 ? bpr : 1024));
            k_transpose_requant<<<grid, 128, 0, ctx->stream>>>(d_deq, (uint32_t)K, (uint32_t)N, (uint32_t)bs, (uint32_t)bpr, t_data, t_scales);
            ZG_COUNT_LAUNCH();
            ok = cudaStreamSynchronize(ctx->stream) == cudaSuccess;
        }
        cudaFree(d_deq);
        if (!ok) {
            zg_set_error("prepare_transposed: kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
            cudaFree(t_data); cudaFree(t_scales); cudaFree(x_q); cudaFree(x_s);
            return -1;
        }
        w->t_data = t_data; w->t_scales = t_scales; w->x_q = x_q; w->x_s = x_s;
        w->device_bytes += N * K + N * bpr * sizeof(float);
    }
    if (h_t_data) cudaMemcpyAsync(h_t_data, w->t_data, N * K, cudaMemcpyDeviceToHost, ctx->stream);
    if (h_t_scales) cudaMemcpyAsync(h_t_scales, w->t_scales, N * bpr * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { zg_set_error("prepare_transposed: copy failed"); return -1; }
    return 0;
}

extern "C" int zg_cuda_quantize_input_host(ZgCudaCtx* ctx, const float* h_input, size_t K, size_t block_size, int8_t* h_q, float* h_scales) {
    if (!ctx || !h_input || !h_q || !h_scales) { zg_set_error("quantize_input: bad arguments"); return -1; }
    if (!check_limits(K, block_size, "quantize_input")) return -1;
    cudaSetDevice(ctx->device);
    const size_t bpr = (K + block_size - 1) / block_size;
    float* d_x = nullptr; int8_t* d_q = nullptr; float* d_s = nullptr;
    if (cudaMalloc(&d_x, K * sizeof(float)) != cudaSuccess || cudaMalloc(&d_q, K) != cudaSuccess || cudaMalloc(&d_s, bpr * sizeof(float)) != cudaSuccess) {
        zg_set_error("quantize_input: cudaMalloc failed");
        cudaFree(d_x); cudaFree(d_q); cudaFree(d_s);
        return -1;
    }
    cudaMemcpyAsync(d_x, h_input, K * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
    bool ok = launch_quantize(d_x, K, block_size, d_q, d_s, ctx->stream);
    cudaMemcpyAsync(h_q, d_q, K, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(h_scales, d_s, bpr * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    ok = cudaStreamSynchronize(ctx->stream) == cudaSuccess && ok;
    cudaFree(d_x); cudaFree(d_q); cudaFree(d_s);
    if (!ok) { zg_set_error("quantize_input: kernel failed"); return -1; }
    return 0;
}

extern "C" int zg_cuda_gemv_w8a8_device(ZgCudaCtx* ctx, const ZgCudaQWeight* w, const float* d_input, float* d_dst) {
    if (!ctx || !w || !d_input || !d_dst) { zg_set_error("gemv_w8a8: bad arguments"); return -1; }
    if (!w->t_data) { zg_set_error("gemv_w8a8: the weight has no transposed form (call zg_cuda_qweight_prepare_transposed first; src/quant.zig:444-445)"); return -1; }
    if (!check_limits(w->K, w->bs, "gemv_w8a8")) return -1;
    cudaSetDevice(ctx->device);
    const size_t K = w->K, bs = w->bs;
    const uint32_t bpr = (uint32_t)((K + bs - 1) / bs);
    bool ok = true;
    if (!launch_quantize(d_input, K, bs, w->x_q, w->x_s, ctx->stream)) { zg_set_error("gemv_w8a8: quantize launch failed"); return -1; }
    const size_t gs = bs / 16;
    const bool fast = K % 16 == 0 && bs % 16 == 0 && gs <= 32 && (gs & (gs - 1)) == 0;
    if (fast) {
        const size_t smem = (size_t)bpr * sizeof(float) * kGemvWarps;
        switch (gs) {
            case 1: ok = launch_fast<1>(w, d_dst, bpr, smem, ctx->stream); break;
            case 2: ok = launch_fast<2>(w, d_dst, bpr, smem, ctx->stream); break;
            case 4: ok = launch_fast<4>(w, d_dst, bpr, smem, ctx->stream); break;
            case 8: ok = launch_fast<8>(w, d_dst, bpr, smem, ctx->stream); break;
            case 16: ok = launch_fast<16>(w, d_dst, bpr, smem, ctx->stream); break;
            default: ok = launch_fast<32>(w, d_dst, bpr, smem, ctx->stream); break;
        }
    } else {
        ok = launch_pdl(k_gemv_w8a8_generic, (unsigned)((w->N + 127) / 128), 128u, (size_t)0, ctx->stream, (const int8_t*)w->t_data,
                        (const float*)w->t_scales, (const int8_t*)w->x_q, (const float*)w->x_s, d_dst, (uint32_t)w->N, (uint32_t)K,
                        (uint32_t)bs, bpr);
    }
    cudaError_t e = cudaGetLastError();
    if (!ok || e != cudaSuccess) { zg_set_error("gemv_w8a8: launch failed: %s", cudaGetErrorString(e)); return -1; }
    return 0;
}

extern "C" int zg_cuda_gemv_w8a8_host(ZgCudaCtx* ctx, const ZgCudaQWeight* w, const float* h_input, float* h_dst) {
    if (!ctx || !w || !h_input || !h_dst) { zg_set_error("gemv_w8a8_host: bad arguments"); return -1; }
    cudaSetDevice(ctx->device);
    float* d_x = nullptr; float* d_y = nullptr;
    if (cudaMalloc(&d_x, (w->K ? w->K : 1) * sizeof(float)) != cudaSuccess || cudaMalloc(&d_y, (w->N ? w->N : 1) * sizeof(float)) != cudaSuccess) {
        zg_set_error("gemv_w8a8_host: cudaMalloc failed");
        cudaFree(d_x); cudaFree(d_y);
        return -1;
    }
    cudaMemcpyAsync(d_x, h_input, w->K * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
    int rc = zg_cuda_gemv_w8a8_device(ctx, w, d_x, d_y);
    if (rc == 0) {
        cudaMemcpyAsync(h_dst, d_y, w->N * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { zg_set_error("gemv_w8a8_host: %s", cudaGetErrorString(e)); rc = -1; }
    } else {
        cudaStreamSynchronize(ctx->stream);
    }
    cudaFree(d_x); cudaFree(d_y);
    return rc;
}
