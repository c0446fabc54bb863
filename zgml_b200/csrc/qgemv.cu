// Decode-path quantized matvec / small-batch matmul for sm_100a.
//
//   dst[m, n] = sum_k x[m, k] * ( f32(q[k, n]) * s[(k*N + n) / 32] )
//
// i.e. zgml's W8·f32 algorithm (QuantizedWeight.matmul, src/quant.zig:475-578;
// DeviceOp.qmatmul, src/backend/reference.zig:499-566) on the packed GPU-resident
// records of zg_internal.cuh.  HBM-bound: one CTA owns (one 64-column tile) x (a
// contiguous run of k-chunk records); thread 0 streams that span into shared
// memory with cp.async.bulk (TMA bulk copy, mbarrier complete_tx) one pipeline
// stage per 4 (int8) / 8 (int4) records, all stages in flight from the first
// instruction; 256 threads unpack bytes/nibbles in registers (PRMT / LOP3 magic
// number -> f32), scale by s*x and accumulate in fp32; warp-shuffle + shared
// memory reduction; split-K partials are combined deterministically by the last
// CTA of each column tile (no atomics on the output, no pre-zeroing).
#include "zg_internal.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxSteps = 32;

struct QGemvParams {
    const uint8_t* recs;
    uint32_t rec_bytes, q_bytes, n_kc, n_tiles;
    uint32_t K, N, M;
    const float* x;
    uint32_t x_rs;
    float* out;
    uint32_t out_rs;
    uint32_t n_splits, r_max;
    float* partials;
    uint32_t* counters;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// u8 (= q + 128) in byte `sel` of w -> f32(q): PRMT builds bits 0x4B0000uu = 2^23 + u.
template <int SEL>
__device__ __forceinline__ float u8_to_f32(uint32_t w) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650 + SEL)) - 8388736.0f;
}

// FMT: ZG_QFMT_I8_F32 / ZG_QFMT_I8_F16 / ZG_QFMT_I4_F16.  MB: activation rows per CTA.
template <int FMT, int MB>
__global__ void __launch_bounds__(kThreads, ((FMT == ZG_QFMT_I4_F16 ? 32 : 16) * MB <= 16) ? 4
                                            : (((FMT == ZG_QFMT_I4_F16 ? 32 : 16) * MB <= 32) ? 3 : 2))
qgemv_kernel(const QGemvParams p) {
    constexpr bool kI4 = (FMT == ZG_QFMT_I4_F16);
    constexpr int RS = kI4 ? 8 : 4;          // records per pipeline stage
    constexpr int TPR = kThreads / RS;       // threads per record: 64 / 32
    constexpr int NACC = kI4 ? 32 : 16;      // output columns per thread

    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t t = threadIdx.x;
    const uint32_t tile = blockIdx.x / p.n_splits;
    const uint32_t split = blockIdx.x % p.n_splits;
    const uint32_t m0 = blockIdx.y * MB;
    const uint32_t kc0 = (uint32_t)(((uint64_t)split * p.n_kc) / p.n_splits);
    const uint32_t kc1 = (uint32_t)(((uint64_t)(split + 1) * p.n_kc) / p.n_splits);
    const uint32_t R = kc1 - kc0;
    const uint32_t n_steps = (R + RS - 1) / RS;

    uint8_t* rec_s = smem;
    float* x_s = reinterpret_cast<float*>(smem + (size_t)p.r_max * p.rec_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(x_s + (size_t)MB * p.r_max * ZG_KC);
    __shared__ uint32_t s_is_last;

    if (t == 0) {
        for (uint32_t s = 0; s < n_steps; s++) mbar_init(smem_u32(&bars[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint8_t* src = p.recs + ((size_t)tile * p.n_kc + kc0) * p.rec_bytes;
        for (uint32_t s = 0; s < n_steps; s++) {
            uint32_t nrec = min((uint32_t)RS, R - s * RS);
            uint32_t bytes = nrec * p.rec_bytes;
            uint32_t bar = smem_u32(&bars[s]);
            mbar_expect_tx(bar, bytes);
            bulk_g2s(smem_u32(rec_s + (size_t)s * RS * p.rec_bytes), src + (size_t)s * RS * p.rec_bytes, bytes, bar);
        }
    }
    // Stage the activation slice(s) while the weights stream in.
    {
        const uint32_t kbase = kc0 * ZG_KC, klen = R * ZG_KC;
#pragma unroll
        for (int m = 0; m < MB; m++) {
            const bool row_ok = (m0 + m) < p.M;
            const float* xr = p.x + (size_t)(m0 + m) * p.x_rs;
            for (uint32_t k = t; k < klen; k += kThreads) {
                uint32_t kg = kbase + k;
                x_s[m * klen + k] = (row_ok && kg < p.K) ? __ldg(xr + kg) : 0.0f;
            }
        }
    }
    __syncthreads(); // x_s + barrier inits visible

    float acc[MB][NACC];
#pragma unroll
    for (int m = 0; m < MB; m++)
#pragma unroll
        for (int j = 0; j < NACC; j++) acc[m][j] = 0.0f;

    const uint32_t rin = t / TPR;  // record within the stage
    const uint32_t tt = t % TPR;
    const uint32_t rq = kI4 ? (tt >> 1) : (tt >> 2);
    const uint32_t nb = kI4 ? (tt & 1) : ((tt & 3) >> 1);
    const uint32_t klen = R * ZG_KC;

    for (uint32_t s = 0; s < n_steps; s++) {
        mbar_wait(smem_u32(&bars[s]), 0);
        const uint32_t rec = s * RS + rin;
        if (rec >= R) continue;
        const uint8_t* base = rec_s + (size_t)rec * p.rec_bytes;
        uint4 q[4];
#pragma unroll
        for (int i = 0; i < 4; i++) q[i] = *reinterpret_cast<const uint4*>(base + (i * TPR + tt) * 16);
        float sc[4];
        if constexpr (FMT == ZG_QFMT_I8_F32) {
            float4 sv = *reinterpret_cast<const float4*>(base + p.q_bytes + (rq * 8 + nb * 4) * 4);
            sc[0] = sv.x; sc[1] = sv.y; sc[2] = sv.z; sc[3] = sv.w;
        } else {
            uint2 sv = *reinterpret_cast<const uint2*>(base + p.q_bytes + (rq * 8 + nb * 4) * 2);
            float2 a = __half22float2(*reinterpret_cast<const __half2*>(&sv.x));
            float2 b = __half22float2(*reinterpret_cast<const __half2*>(&sv.y));
            sc[0] = a.x; sc[1] = a.y; sc[2] = b.x; sc[3] = b.y;
        }
        float c[MB][4];
#pragma unroll
        for (int m = 0; m < MB; m++) {
            float4 xv = *reinterpret_cast<const float4*>(x_s + m * klen + rec * ZG_KC + rq * 4);
            c[m][0] = sc[0] * xv.x; c[m][1] = sc[1] * xv.y; c[m][2] = sc[2] * xv.z; c[m][3] = sc[3] * xv.w;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t w4[4] = {q[i].x, q[i].y, q[i].z, q[i].w};
            if constexpr (!kI4) {
#pragma unroll
                for (int wi = 0; wi < 4; wi++) {
                    float f0 = u8_to_f32<0>(w4[wi]), f1 = u8_to_f32<1>(w4[wi]);
                    float f2 = u8_to_f32<2>(w4[wi]), f3 = u8_to_f32<3>(w4[wi]);
#pragma unroll
                    for (int m = 0; m < MB; m++) {
                        acc[m][wi * 4 + 0] = fmaf(f0, c[m][i], acc[m][wi * 4 + 0]);
                        acc[m][wi * 4 + 1] = fmaf(f1, c[m][i], acc[m][wi * 4 + 1]);
                        acc[m][wi * 4 + 2] = fmaf(f2, c[m][i], acc[m][wi * 4 + 2]);
                        acc[m][wi * 4 + 3] = fmaf(f3, c[m][i], acc[m][wi * 4 + 3]);
                    }
                }
            } else {
                // nibble at bit 4j of a 16-bit half-word -> 2^23 + 16^j * (q+8); subtracting
                // 2^23 + 8*16^j leaves 16^j * q exactly, and c is pre-divided by 16^j.
                float cj[MB][4];
#pragma unroll
                for (int m = 0; m < MB; m++) {
                    cj[m][0] = c[m][i]; cj[m][1] = c[m][i] * 0.0625f;
                    cj[m][2] = c[m][i] * 0.00390625f; cj[m][3] = c[m][i] * 0.000244140625f;
                }
#pragma unroll
                for (int wi = 0; wi < 4; wi++) {
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const uint32_t v = h ? (w4[wi] >> 16) : w4[wi];
                        float f0 = __uint_as_float((v & 0x0000000Fu) | 0x4B000000u) - 8388616.0f;
                        float f1 = __uint_as_float((v & 0x000000F0u) | 0x4B000000u) - 8388736.0f;
                        float f2 = __uint_as_float((v & 0x00000F00u) | 0x4B000000u) - 8390656.0f;
                        float f3 = __uint_as_float((v & 0x0000F000u) | 0x4B000000u) - 8421376.0f;
                        const int e_lo0 = wi * 4 + h * 2, e_lo1 = e_lo0 + 1; // bytes 2h, 2h+1 of word wi
#pragma unroll
                        for (int m = 0; m < MB; m++) {
                            acc[m][e_lo0] = fmaf(f0, cj[m][0], acc[m][e_lo0]);
                            acc[m][16 + e_lo0] = fmaf(f1, cj[m][1], acc[m][16 + e_lo0]);
                            acc[m][e_lo1] = fmaf(f2, cj[m][2], acc[m][e_lo1]);
                            acc[m][16 + e_lo1] = fmaf(f3, cj[m][3], acc[m][16 + e_lo1]);
                        }
                    }
                }
            }
        }
    }

    // ── reduce over the threads that share a column group ──
    constexpr int kGroups = kI4 ? 2 : 4; // distinct column groups per record
#pragma unroll
    for (int m = 0; m < MB; m++)
#pragma unroll
        for (int j = 0; j < NACC; j++) {
            float v = acc[m][j];
#pragma unroll
            for (int off = kGroups; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            acc[m][j] = v;
        }
    __syncthreads(); // everyone is done reading rec_s; reuse it as scratch
    float* red = reinterpret_cast<float*>(smem); // [8 warps][MB][64]
    const uint32_t lane = t & 31, warp = t >> 5;
    if (lane < kGroups) {
#pragma unroll
        for (int m = 0; m < MB; m++)
#pragma unroll
            for (int j = 0; j < NACC; j++) red[(warp * MB + m) * ZG_TN + lane * NACC + j] = acc[m][j];
    }
    __syncthreads();
    const uint32_t Np = p.n_tiles * ZG_TN;
    for (uint32_t idx = t; idx < MB * ZG_TN; idx += kThreads) {
        const uint32_t m = idx / ZG_TN, col = idx % ZG_TN;
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < kThreads / 32; w++) v += red[(w * MB + m) * ZG_TN + col];
        const uint32_t n = tile * ZG_TN + col;
        if (m0 + m < p.M) {
            if (p.n_splits == 1) {
                if (n < p.N) p.out[(size_t)(m0 + m) * p.out_rs + n] = v;
            } else {
                p.partials[((size_t)split * p.M + (m0 + m)) * Np + n] = v;
            }
        }
    }
    if (p.n_splits == 1) return;

    __threadfence();
    __syncthreads();
    if (t == 0) {
        uint32_t old = atomicAdd(&p.counters[blockIdx.y * p.n_tiles + tile], 1u);
        s_is_last = (old == p.n_splits - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_is_last) return;
    __threadfence();
    for (uint32_t idx = t; idx < MB * ZG_TN; idx += kThreads) {
        const uint32_t m = idx / ZG_TN, col = idx % ZG_TN;
        const uint32_t n = tile * ZG_TN + col;
        if (m0 + m >= p.M || n >= p.N) continue;
        float v = 0.0f;
        for (uint32_t sp = 0; sp < p.n_splits; sp++) // fixed order: deterministic
            v += __ldcg(&p.partials[((size_t)sp * p.M + (m0 + m)) * Np + n]);
        p.out[(size_t)(m0 + m) * p.out_rs + n] = v;
    }
    if (t == 0) p.counters[blockIdx.y * p.n_tiles + tile] = 0u; // re-arm for the next launch
}

// Generic block size / ragged N: one thread per output column, exact scale lookup
// per element ((k*N+n)/bs) like src/backend/reference.zig:540-563.
__global__ void qmatmul_generic_kernel(const int8_t* __restrict__ data, const float* __restrict__ scales,
                                       uint32_t bs, const float* __restrict__ x, uint32_t x_rs,
                                       float* __restrict__ out, uint32_t out_rs, uint32_t M, uint32_t N,
                                       uint32_t K) {
    uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t m = blockIdx.y;
    if (n >= N || m >= M) return;
    const float* xr = x + (size_t)m * x_rs;
    float acc = 0.0f;
    for (uint32_t k = 0; k < K; k++) {
        size_t flat = (size_t)k * N + n;
        float c = scales[flat / bs] * xr[k];
        acc = fmaf((float)data[flat], c, acc);
    }
    out[(size_t)m * out_rs + n] = acc;
}

template <int FMT, int MB>
bool set_smem_attr() {
    cudaError_t e = cudaFuncSetAttribute(qgemv_kernel<FMT, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { zg_set_error("cudaFuncSetAttribute(qgemv) failed: %s", cudaGetErrorString(e)); return false; }
    return true;
}

template <int FMT, int MB>
bool launch_fast(ZgCudaCtx* ctx, const ZgCudaQWeight* w, const ZgGemvPlan& plan, QGemvParams& p, cudaStream_t st) {
    dim3 grid(plan.grid, (p.M + MB - 1) / MB);
    qgemv_kernel<FMT, MB><<<grid, kThreads, plan.smem_bytes, st>>>(p);
    ZG_COUNT_LAUNCH();
    (void)ctx; (void)w;
    return true;
}

} // namespace

ZgGemvPlan zg_qgemv_plan(const ZgCudaCtx* ctx, const ZgCudaQWeight* w, uint32_t M) {
    ZgGemvPlan pl;
    const bool i4 = (w->fmt == ZG_QFMT_I4_F16);
    const uint32_t RS = i4 ? 8 : 4;
    pl.m_block = i4 ? (M >= 2 ? 2 : 1) : (M >= 4 ? 4 : (M >= 2 ? 2 : 1));
    const uint32_t sm = (uint32_t)ctx->sm_count;
    const uint32_t smem_cap = 96 * 1024;
    uint64_t best_cost = ~0ull;
    uint32_t best_splits = 1;
    for (uint32_t ns = 1; ns <= w->n_kc && ns <= 256; ns++) {
        uint32_t R = (w->n_kc + ns - 1) / ns;
        size_t smem = (size_t)R * w->rec_bytes + (size_t)pl.m_block * R * ZG_KC * 4 + kMaxSteps * 8;
        if (smem > smem_cap && ns < w->n_kc) continue;
        if ((R + RS - 1) / RS > kMaxSteps) continue;
        uint64_t ctas = (uint64_t)w->n_tiles * ns;
        uint64_t per_sm = (ctas + sm - 1) / sm;
        uint32_t steps = (R + RS - 1) / RS;
        uint64_t cost = per_sm * ((uint64_t)steps * RS + 3);
        if (cost < best_cost) { best_cost = cost; best_splits = ns; }
    }
    pl.n_splits = best_splits;
    pl.rec_per_cta = (w->n_kc + best_splits - 1) / best_splits;
    size_t red_bytes = (size_t)(kThreads / 32) * pl.m_block * ZG_TN * 4;
    size_t rec_area = (size_t)pl.rec_per_cta * w->rec_bytes;
    if (rec_area < red_bytes) rec_area = red_bytes; // scratch reuse needs room (tiny K)
    // keep x_s 16-byte aligned behind the record area
    pl.smem_bytes = (uint32_t)(rec_area + (size_t)pl.m_block * pl.rec_per_cta * ZG_KC * 4 + kMaxSteps * 8);
    pl.grid = w->n_tiles * pl.n_splits;
    return pl;
}

// Opt every instantiation into >48 KB dynamic shared memory once per context,
// outside any stream capture.
bool zg_qgemv_init(ZgCudaCtx*) {
    return set_smem_attr<ZG_QFMT_I8_F32, 1>() && set_smem_attr<ZG_QFMT_I8_F32, 2>() && set_smem_attr<ZG_QFMT_I8_F32, 4>() &&
           set_smem_attr<ZG_QFMT_I8_F16, 1>() && set_smem_attr<ZG_QFMT_I8_F16, 2>() && set_smem_attr<ZG_QFMT_I8_F16, 4>() &&
           set_smem_attr<ZG_QFMT_I4_F16, 1>() && set_smem_attr<ZG_QFMT_I4_F16, 2>();
}

void zg_qgemv_ws_need(const ZgCudaCtx* ctx, const ZgCudaQWeight* w, uint32_t M, size_t* partial_elems,
                      size_t* counters) {
    *partial_elems = 0; *counters = 0;
    if (w->fmt == ZG_QFMT_GENERIC || M == 0) return;
    ZgGemvPlan plan = zg_qgemv_plan(ctx, w, M);
    if (plan.n_splits <= 1) return;
    *partial_elems = (size_t)plan.n_splits * M * w->n_tiles * ZG_TN;
    *counters = (size_t)((M + plan.m_block - 1) / plan.m_block) * w->n_tiles;
}

bool zg_qmatmul_launch(ZgCudaCtx* ctx, const ZgCudaQWeight* w, const float* d_in, float* d_out,
                       uint32_t M, uint32_t in_rs, uint32_t out_rs, const ZgGemvWs* ws, cudaStream_t st) {
    if (M == 0 || w->N == 0) return true;
    if (in_rs == 0) in_rs = (uint32_t)w->K;   // src/backend/reference.zig:509
    if (out_rs == 0) out_rs = (uint32_t)w->N; // src/backend/reference.zig:510
    if (w->fmt == ZG_QFMT_GENERIC) {
        dim3 grid((unsigned)((w->N + 127) / 128), M);
        qmatmul_generic_kernel<<<grid, 128, 0, st>>>(w->g_data, w->g_scales, (uint32_t)w->bs, d_in, in_rs,
                                                      d_out, out_rs, M, (uint32_t)w->N, (uint32_t)w->K);
        ZG_COUNT_LAUNCH();
        return true;
    }
    ZgGemvPlan plan = zg_qgemv_plan(ctx, w, M);
    QGemvParams p;
    p.recs = w->recs; p.rec_bytes = w->rec_bytes; p.q_bytes = w->q_bytes;
    p.n_kc = w->n_kc; p.n_tiles = w->n_tiles;
    p.K = (uint32_t)w->K; p.N = (uint32_t)w->N; p.M = M;
    p.x = d_in; p.x_rs = in_rs; p.out = d_out; p.out_rs = out_rs;
    p.n_splits = plan.n_splits; p.r_max = plan.rec_per_cta;
    if (plan.n_splits > 1) {
        size_t pe = 0, nc = 0;
        zg_qgemv_ws_need(ctx, w, M, &pe, &nc);
        if (!ws || ws->partials_elems < pe || ws->counters_n < nc) {
            zg_set_error("internal: split-K workspace too small (%zu/%zu needed)", pe, nc);
            return false;
        }
        p.partials = ws->partials; p.counters = ws->counters;
    } else {
        p.partials = nullptr; p.counters = nullptr;
    }
    switch (w->fmt) {
        case ZG_QFMT_I8_F32:
            if (plan.m_block == 4) return launch_fast<ZG_QFMT_I8_F32, 4>(ctx, w, plan, p, st);
            if (plan.m_block == 2) return launch_fast<ZG_QFMT_I8_F32, 2>(ctx, w, plan, p, st);
            return launch_fast<ZG_QFMT_I8_F32, 1>(ctx, w, plan, p, st);
        case ZG_QFMT_I8_F16:
            if (plan.m_block == 4) return launch_fast<ZG_QFMT_I8_F16, 4>(ctx, w, plan, p, st);
            if (plan.m_block == 2) return launch_fast<ZG_QFMT_I8_F16, 2>(ctx, w, plan, p, st);
            return launch_fast<ZG_QFMT_I8_F16, 1>(ctx, w, plan, p, st);
        case ZG_QFMT_I4_F16:
            if (plan.m_block == 2) return launch_fast<ZG_QFMT_I4_F16, 2>(ctx, w, plan, p, st);
            return launch_fast<ZG_QFMT_I4_F16, 1>(ctx, w, plan, p, st);
        default: break;
    }
    zg_set_error("qmatmul: unknown weight format %d", w->fmt);
    return false;
}
