// Decode-path quantized matvec / small-batch matmul for sm_100a.
//
//   dst[m, n] = sum_k ( x[m, k] * s[(k*N + n) / 32] ) * q[k, n]
//
// i.e. zgml's W8·f32 algorithm (QuantizedWeight.matmul, src/quant.zig:475-578;
// DeviceOp.qmatmul, src/backend/reference.zig:499-566) on the packed GPU-resident
// records of zg_internal.cuh.  The kernel is HBM-bound by construction:
//
//  * Every weight byte is read exactly once with fully coalesced 128-bit loads
//    (512 contiguous bytes per warp instruction) that land DIRECTLY in the register
//    layout of an mma.sync.m16n8k32 A fragment: no shared-memory staging, no
//    transposition and (int8) no unpack.  Each warp keeps U records (U x 0.6-1.1 KB)
//    in flight in a register ring; the first U are issued before
//    griddepcontrol.wait, i.e. while the previous kernel of the stream still runs
//    (programmatic dependent launch: weights are immutable, activations are not).
//  * The weights never pass through a dequantize step.  c[k] = x[k] * s[k, nb]
//    (the reference's first rounding, `scale * input_v`) is computed once per
//    32 weights, converted to 23-bit fixed point in a per-(row, quant-block, k-range)
//    scale E2 ~ 2 max|x| max(s) with one FFMA (F = c / E2 + 1.5 -> the three low
//    bytes of F are base-256 digits), and the digits multiply the raw int8 / int4
//    weights on the integer tensor-core path (IMMA).  A column of ones in B yields
//    sum_k q[k, n], which removes the digit bias.  Integer accumulation is exact; the
//    only rounding beyond the reference's is the 2^-23 fixed-point grid of c relative
//    to max|x| * max(s) of the warp's k-range.
//  * int4 weights stay packed: byte = 16 u[n+8] + u[n] (u = q + 8) is fed to the
//    MMA as one u8 operand row next to the row u[n] = byte & 15; the two results
//    separate exactly ((X - Y) / 16), and the +8 bias is removed with per-digit sums.
//  * Work split: CTA = (group of P column groups) x (one of S k-splits); its warps
//    split the CTA's k-range.  Per column group the warps' partial sums are combined
//    in shared memory in fixed order; with S > 1 the CTA partials meet in a global
//    scratch and the last-arriving CTA sums them in split order -> deterministic.
#include "zg_internal.cuh"

#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <type_traits>

ZG_TRACE_DECL
void zg_trace_set_gemv(unsigned long long* d_buf) { cudaMemcpyToSymbol(c_zg_trace, &d_buf, sizeof(d_buf)); }

namespace {

constexpr int kMaxWarps = 8;
constexpr int kThreads = kMaxWarps * 32;
constexpr uint32_t kLcap = 16;               // records of activations a warp may stage (2 KB per row)

struct QGemvParams {
    const uint8_t* recs;
    const float* smax;
    uint32_t n_kc, n_nb;
    uint32_t K, N, M;
    const float* x;
    uint32_t x_rs;
    uint32_t xs_stride;    // floats per staged activation row of a warp (shared memory: [warp][row][xs_stride])
    float* out;
    uint32_t out_rs;
    uint32_t P, S, NS;
    uint32_t cl;           // 1: the S k-splits of a column group are the CTAs of one thread-block cluster; partial sums meet in distributed shared memory
    float* partials;
    uint32_t* counters;
};
struct QGemvParamsPro : QGemvParams { ZgGemvPrologue pro; };   // only the prologue-fusion instantiations carry the recipe

// Up to kZgGemvBatch independent matvecs of one shape / format share a launch (blockIdx.y selects the op): q|k|v,
// gate|up, or the copies of a microbenchmark.  The parameter block stays in the constant bank.
template <bool HAS_PRO>
struct QGemvBatchT { std::conditional_t<HAS_PRO, QGemvParamsPro, QGemvParams> p[kZgGemvBatch]; };
using QGemvBatch = QGemvBatchT<true>;   // host-side form; narrowed to QGemvBatchT<false> for the plain kernels

// D(16x8, s32) += A(16x32: weights) * B(32x8, u8: digits of x*s)
__device__ __forceinline__ void imma_s8u8(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void imma_u8u8(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
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
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_cluster_f32(const float* local, uint32_t rank) {
    uint32_t remote; float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"((uint32_t)__cvta_generic_to_shared(local)), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ void sts8(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

constexpr uint32_t kPlaneRow = 144;   // bytes per (record, activation row) digit planes: 4 x 32 B + 16 B bank skew

// FMT: ZG_QFMT_I8_F32 / ZG_QFMT_I8_F16 / ZG_QFMT_I4_F16.  MP: pairs of activation rows (M <= 2*MP).
// XR: activation rows staged per warp (1 for the decode matvec M == 1, else 2*MP).
// PRO: 0 = activations are read; 1 / 2 = produced in the prologue (ZgGemvPrologue; M <= 2 only)
template <int FMT, int MP, int XR, int PRO = 0>
__global__ void __launch_bounds__(kThreads, MP == 1 ? 3 : (MP == 2 ? 2 : 1))
qgemv_kernel(const __grid_constant__ QGemvBatchT<PRO != 0> bt) {
    const auto& p = bt.p[blockIdx.y];
    constexpr bool kI4 = (FMT == ZG_QFMT_I4_F16);
    constexpr bool kF32 = (FMT == ZG_QFMT_I8_F32);
    constexpr int MR = 2 * MP;
    constexpr uint32_t QB = kI4 ? 512u : 1024u;
    constexpr uint32_t SB = kF32 ? 32u : 16u;  // scale bytes per t
    constexpr uint32_t RB = QB + 4 * SB;       // record bytes

    __shared__ float part[2][kMaxWarps][MR][ZG_TN];
    __shared__ float xpart[2][MR][ZG_TN];   // cluster split-K: this CTA's partial of the current column group, read by the split-0 CTA
    __shared__ uint32_t s_last;
    // dynamic: [warp][row][xs_stride] scaled activations | [warp][slot][G records] weight ring |
    //          [warp][G][MR][kPlaneRow] digit planes | [warp][slot] mbarriers
    extern __shared__ __align__(128) uint8_t dsm[];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    const uint32_t g = lane >> 2, t = lane & 3, j = g & 3;
    const uint32_t split = blockIdx.x % p.S, grp = blockIdx.x / p.S;
    const uint32_t nb_begin = grp * p.P;
    const uint32_t nb_end = min(nb_begin + p.P, p.n_nb);
    const uint32_t ks = (uint32_t)(((uint64_t)split * p.n_kc) / p.S);
    const uint32_t ke = (uint32_t)(((uint64_t)(split + 1) * p.n_kc) / p.S);
    const uint32_t k0 = ks + (warp * (ke - ks)) / W, k1 = ks + ((warp + 1) * (ke - ks)) / W;
    const uint32_t L = k1 - k0;                    // records per column group for this warp (<= lcap)
    constexpr uint32_t G = kI4 ? 4u : 2u;          // records per ring slot (one TMA bulk copy)
    const uint32_t NS = p.NS;                      // ring slots
    const uint32_t n_chunk = (L + G - 1) / G;      // chunks (= bulk copies) per column group
    const uint32_t total_chunks = n_chunk * (nb_end - nb_begin);
    constexpr uint32_t slot_bytes = G * RB;

    const uint32_t dsm_u32 = smem_u32(dsm);
    const uint32_t xs_bytes = W * XR * p.xs_stride * 4, ring_bytes = W * NS * slot_bytes, plane_bytes = G * XR * kPlaneRow;
    float* xs_w = reinterpret_cast<float*>(dsm) + (size_t)warp * XR * p.xs_stride;
    const uint32_t xs_u32 = dsm_u32 + warp * XR * p.xs_stride * 4;
    const uint32_t ring = dsm_u32 + xs_bytes + warp * NS * slot_bytes;
    const uint32_t planes = dsm_u32 + xs_bytes + ring_bytes + warp * plane_bytes;
    const uint32_t bars = dsm_u32 + xs_bytes + ring_bytes + W * plane_bytes + warp * NS * 8;

    // Programmatic dependent launch: the next kernel in the stream may begin (and prefetch ITS weights, which
    // nobody writes) as soon as SM resources free up.  Everything mutable (x, out, partials, counters) is
    // only touched after griddepcontrol.wait, i.e. after the previous kernel has fully completed.
    ZG_TRACE_BEGIN(10)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // ── per-warp ring of NS slots in shared memory, filled by TMA bulk copies (one per chunk of <= G
    //    consecutive records, issued by lane 0, completion on the slot's mbarrier).  The first NS chunks are
    //    requested right away. ──
    // producer cursor (lane 0): next chunk to request = chunk pf_c of column group offset pf_nb, into slot pf_slot
    uint32_t pf_c = 0, pf_slot = 0, pf_left = total_chunks;
    const uint8_t* pf_src = p.recs + ((size_t)nb_begin * p.n_kc + k0) * RB;   // this warp's run in the current pf column group
    auto issue_chunk = [&]() {   // lane 0 only
        const uint32_t cnt = min(G, L - pf_c * G);
        const uint32_t bar = bars + pf_slot * 8;
        mbar_expect_tx(bar, cnt * RB);
        bulk_g2s(ring + pf_slot * slot_bytes, pf_src + (size_t)pf_c * G * RB, cnt * RB, bar);
        pf_left--;
        if (++pf_c == n_chunk) { pf_c = 0; pf_src += (size_t)p.n_kc * RB; }
        if (++pf_slot == NS) pf_slot = 0;
    };
    if (lane == 0) {
        for (uint32_t s = 0; s < NS; s++) mbar_init(bars + s * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (uint32_t i = 0; i < NS && pf_left; i++) issue_chunk();
    }
    // Measured and rejected (round 2): pulling the REST of the warp's weight run into L2 here with cp.async.bulk.prefetch.L2
    // (the kernel sits resident 5-11 us before its activations exist) made every configuration slower — 70B shard of 8:
    // 61.2 -> 62.1 us/layer, unsharded 175 -> 198, GEMV microbenchmark 5.79 -> 5.24 TB/s: the wait is not idle HBM time.
    // the constant ones plane (digit index 3) of every (record, row): B column of ones -> sum_k q
    for (uint32_t i = lane; i < G * XR * 8; i += 32) {
        const uint32_t rm = i >> 3, w4 = i & 7;
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(planes + rm * kPlaneRow + 96 + w4 * 4), "r"(0x01010101u) : "memory");
    }
    __syncwarp();
    float sm_next = __ldg(p.smax + nb_begin);      // power of two >= every scale of the column group

    asm volatile("griddepcontrol.wait;" ::: "memory");
    ZG_TRACE_MARK(1)

    // ── stage x'[m, k] = x[m, k] * 0.499 / (max|x| * smax) of this warp's k-range in shared memory (so that
    //    |s * x'| <= 0.499).  Non-finite activations poison the partial sums (NaN out, like the reference);
    //    k >= K reads as zero; only rows < M are staged (the other B columns alias row M - 1). ──
    float inv_rms[XR];
    if constexpr (PRO != 0) {
#pragma unroll
        for (int m = 0; m < XR; m++) inv_rms[m] = 0.0f;
    }
    if constexpr (PRO == 1) {
        // rmsnorm scale of every staged row: all 256 threads walk the whole row (K floats, L2-resident), fixed reduction
        // order -> every CTA of every consumer matvec computes the identical value
        __shared__ float s_red[kMaxWarps][XR];
#pragma unroll
        for (int m = 0; m < XR; m++) {
            float ss = 0.0f;
            if (XR == 1 || (uint32_t)m < p.M) {
                const float* ar = p.pro.a + (size_t)m * p.K;
                const float* br = p.pro.b ? p.pro.b + (size_t)m * p.K : nullptr;
                if ((p.K & 3u) == 0) {
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
                    for (uint32_t jq = tid; jq < (p.K >> 2); jq += kThreads) {
                        float4 v = reinterpret_cast<const float4*>(ar)[jq];
                        if (br) { const float4 t = reinterpret_cast<const float4*>(br)[jq]; v.x = __fadd_rn(v.x, t.x); v.y = __fadd_rn(v.y, t.y); v.z = __fadd_rn(v.z, t.z); v.w = __fadd_rn(v.w, t.w); }
                        s0 += v.x * v.x; s1 += v.y * v.y; s2 += v.z * v.z; s3 += v.w * v.w;
                    }
                    ss = (s0 + s1) + (s2 + s3);
                } else {
                    for (uint32_t jq = tid; jq < p.K; jq += kThreads) { float v = ar[jq]; if (br) v = __fadd_rn(v, br[jq]); ss += v * v; }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            if (lane == 0) s_red[warp][m] = ss;
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < XR; m++) {
            float tot = 0.0f;
            for (uint32_t w2 = 0; w2 < W; w2++) tot += s_red[w2][m];
            inv_rms[m] = 1.0f / sqrtf(tot / (float)p.K + p.pro.eps);
        }
    }
    bool pro_write = false;   // column-group block 0: its S splits x 8 warps cover every k once
    if constexpr (PRO != 0) pro_write = p.pro.write != 0 && grp == 0;
    // activation k of row m: read (PRO 0) or produced by the absorbed ops, whose outputs the writer CTAs also store
    auto load_x = [&](uint32_t m, uint32_t kidx) -> float {
        if constexpr (PRO == 0) {
            return p.x[(size_t)m * p.x_rs + kidx];
        } else if constexpr (PRO == 1) {
            const size_t f = (size_t)m * p.K + kidx;
            float sv = p.pro.a[f];
            if (p.pro.b) sv = __fadd_rn(sv, p.pro.b[f]);
            const float bare = __fmul_rn(sv, inv_rms[m]), gm = p.pro.gamma[kidx], xv = __fmul_rn(bare, gm);
            if (pro_write) {
                if (p.pro.o_sum) p.pro.o_sum[f] = sv;
                p.pro.o_mid[f] = bare; p.pro.o_grep[f] = gm; p.pro.o_x[f] = xv;
            }
            return xv;
        } else {
            const size_t f = (size_t)m * p.K + kidx;
            float v = p.pro.a[f];
            for (uint32_t st = 0; st < p.pro.n_steps; st++) {
                const uint32_t sop = p.pro.steps[st].op, sw = p.pro.steps[st].is_swapped;
                if (sop == ZG_EW_ADD) { const float t2 = p.pro.steps[st].sec[f]; v = sw ? __fadd_rn(t2, v) : __fadd_rn(v, t2); }
                else if (sop == ZG_EW_MUL) { const float t2 = p.pro.steps[st].sec[f]; v = sw ? __fmul_rn(t2, v) : __fmul_rn(v, t2); }
                else if (sop == ZG_EW_NEG) v = -v;
                else if (sop == ZG_EW_EXP) v = expf(v);
                else if (sop == ZG_EW_RECIP) v = 1.0f / v;
                else if (sop == ZG_EW_ABS) v = fabsf(v);
                else if (sop == ZG_EW_RELU) v = fmaxf(v, 0.0f);
                else if (sop == ZG_EW_SQRT) v = sqrtf(v);
                else if (sop == ZG_EW_LOG) v = logf(v);
            }
            const float xv = __fmul_rn(v, p.pro.b[f]);
            if (pro_write) { p.pro.o_mid[f] = v; p.pro.o_x[f] = xv; }
            return xv;
        }
    };
    float xm[MR];
    {
        const uint32_t kb = k0 * ZG_KR;
        const float rsm = 1.0f / sm_next;
#pragma unroll
        for (int m = 0; m < MR; m++) xm[m] = 0.0f;
#pragma unroll
        for (int m = 0; m < XR; m++) {
            if (XR == 1 || (uint32_t)m < p.M) {
                const float* xr = p.x + (size_t)m * p.x_rs + kb + lane;
                float* xd = xs_w + (size_t)m * p.xs_stride + lane;
                float v[kLcap];
                float mx = 0.0f;
#pragma unroll
                for (int i = 0; i < (int)kLcap; i++) {
                    if constexpr (PRO == 0) v[i] = ((uint32_t)i < L && kb + lane + 32 * i < p.K) ? xr[32 * i] : 0.0f;
                    else v[i] = ((uint32_t)i < L && kb + lane + 32 * i < p.K) ? load_x((uint32_t)m, kb + lane + 32 * i) : 0.0f;
                    const float aa = fabsf(v[i]);
                    mx = (aa <= 3.0e38f) ? fmaxf(mx, aa) : INFINITY;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                xm[m] = mx;
                const float f = (mx <= 3.0e38f && mx >= 1.0e-30f) ? (0.499f / mx) * rsm : 0.0f;   // tiny rows flush to zero
#pragma unroll
                for (int i = 0; i < (int)kLcap; i++)
                    if ((uint32_t)i < L) xd[32 * i] = v[i] * f;
            }
        }
        __syncwarp();
    }

    // MMA role of this lane: B column g = digit j (j == 3: the ones plane) of activation row 2*mp + (g >> 2)
    uint32_t brow[MP];
#pragma unroll
    for (int mp = 0; mp < MP; mp++)
        brow[mp] = planes + (XR == 1 ? 0u : min((uint32_t)(2 * mp) + (g >> 2), p.M - 1)) * kPlaneRow + j * 32 + 4 * t;
    const uint32_t q_off = lane * 16;
    // digit-generation role of this lane: k row `lane` of each record
    const uint32_t sc_off = QB + (kF32 ? 4u : 2u) * (8 * ((lane & 15) >> 2) + 4 * (lane >> 4) + (lane & 3));  // zg_scale_row_slot(lane)

    int acc[MP][2][4];
    uint32_t dsum[MP];
    uint32_t buf = 0, slot = 0, parity = 0, slot_u32 = ring, bar_u32 = bars;
    const uint32_t xs_row_bytes = p.xs_stride * 4;
    float sm_prev = sm_next;

    for (uint32_t nb = nb_begin; nb < nb_end; nb++) {
        const float sm = sm_next;
        if (nb + 1 < nb_end) sm_next = __ldg(p.smax + nb + 1);
        if (sm != sm_prev) {
            // re-normalise the staged activations for this column group's scale: exact (powers of two)
            const float ratio = sm_prev / sm;
            for (uint32_t m = 0; m < p.M && m < (uint32_t)XR; m++) {
                float* xd = xs_w + (size_t)m * p.xs_stride + lane;
                for (uint32_t i = 0; i < L; i++) xd[32 * i] *= ratio;
            }
            __syncwarp();
            sm_prev = sm;
        }
#pragma unroll
        for (int mp = 0; mp < MP; mp++) {
            dsum[mp] = 0;
#pragma unroll
            for (int ct = 0; ct < 2; ct++)
#pragma unroll
                for (int i = 0; i < 4; i++) acc[mp][ct][i] = 0;
        }

        uint32_t xa = xs_u32 + lane * 4;   // this lane's (k = lane) staged activation of the chunk's first record
        for (uint32_t c = 0; c < n_chunk; c++, xa += G * ZG_KR * 4) {
            const uint32_t cnt = min(G, L - c * G);
            mbar_wait(bar_u32, parity);
            auto chunk = [&](auto full_tag) {
                constexpr bool FULL = decltype(full_tag)::value;
                // ── digit generation, lane = k: F = s * x' + 1.5 in (1, 2); the three low bytes of F are the
                //    base-256 digits of c / E2 + 0.5 -> byte planes [record][row][digit][k] ──
#pragma unroll
                for (uint32_t r = 0; r < G; r++) {
                    if (FULL || r < cnt) {
                        float sc;
                        if constexpr (kF32) {
                            sc = __uint_as_float(lds32(slot_u32 + sc_off + r * RB));
                        } else {
                            unsigned short h;
                            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(slot_u32 + sc_off + r * RB));
                            sc = __half2float(__ushort_as_half(h));
                        }
#pragma unroll
                        for (int m = 0; m < XR; m++) {
                            if (XR == 1 || (uint32_t)m < p.M) {
                                const uint32_t F = __float_as_uint(fmaf(sc, __uint_as_float(lds32(xa + r * ZG_KR * 4 + m * xs_row_bytes)), 1.5f));
                                const uint32_t pa = planes + lane + (r * XR + m) * kPlaneRow;
                                sts8(pa, F);
                                sts8(pa + 32, F >> 8);
                                sts8(pa + 64, F >> 16);
                            }
                        }
                    }
                }
                __syncwarp();
                // ── MMA: the ring's shared-memory bytes ARE the A fragments, the planes ARE the B fragments ──
#pragma unroll
                for (uint32_t r = 0; r < G; r++) {
                    if (FULL || r < cnt) {
                        const uint32_t qa = slot_u32 + q_off + r * RB;
                        uint32_t a[2][4];
                        if constexpr (!kI4) {
                            const uint4 q0 = lds128(qa), q1 = lds128(qa + 512);
                            a[0][0] = q0.x; a[0][1] = q0.y; a[0][2] = q0.z; a[0][3] = q0.w;
                            a[1][0] = q1.x; a[1][1] = q1.y; a[1][2] = q1.z; a[1][3] = q1.w;
                        } else {
                            // row g: 16 u[n + 8] + u[n], row g + 8: u[n]
                            const uint4 q0 = lds128(qa);
                            a[0][0] = q0.x; a[0][1] = q0.x & 0x0F0F0F0Fu; a[0][2] = q0.y; a[0][3] = q0.y & 0x0F0F0F0Fu;
                            a[1][0] = q0.z; a[1][1] = q0.z & 0x0F0F0F0Fu; a[1][2] = q0.w; a[1][3] = q0.w & 0x0F0F0F0Fu;
                        }
#pragma unroll
                        for (int mp = 0; mp < MP; mp++) {
                            const uint32_t b0 = lds32(brow[mp] + r * XR * kPlaneRow), b1 = lds32(brow[mp] + r * XR * kPlaneRow + 16);
                            if constexpr (kI4) {
                                dsum[mp] = __dp4a(b0, 0x01010101u, __dp4a(b1, 0x01010101u, dsum[mp]));
                                imma_u8u8(acc[mp][0], a[0][0], a[0][1], a[0][2], a[0][3], b0, b1);
                                imma_u8u8(acc[mp][1], a[1][0], a[1][1], a[1][2], a[1][3], b0, b1);
                            } else {
                                imma_s8u8(acc[mp][0], a[0][0], a[0][1], a[0][2], a[0][3], b0, b1);
                                imma_s8u8(acc[mp][1], a[1][0], a[1][1], a[1][2], a[1][3], b0, b1);
                            }
                        }
                    }
                }
            };
            if (cnt == G) chunk(std::true_type{}); else chunk(std::false_type{});
            // ── every lane is done with the slot and the planes: lane 0 requests the chunk NS ahead ──
            __syncwarp();
            if (lane == 0 && pf_left) issue_chunk();
            slot_u32 += slot_bytes; bar_u32 += 8;
            if (++slot == NS) { slot = 0; slot_u32 = ring; bar_u32 = bars; parity ^= 1; }
        }
        // ── flush: integer sums -> float partials of this warp in shared memory ──
        {
            const uint32_t kcnt = L * ZG_KR;   // rows fed to the MMA
            float* dstp = &part[buf][warp][0][0];
#pragma unroll
            for (int mp = 0; mp < MP; mp++) {
                const float xw = (t >> 1) ? xm[2 * mp + 1] : xm[2 * mp];   // row this lane writes
                // 2^-23 * E2, E2 = max|x| * smax / 0.499
                const float esc = (xw <= 3.0e38f) ? ((xw >= 1.0e-30f) ? (xw * (2.004008016f * 1.1920928955078125e-07f)) * sm : 0.0f)
                                                  : __int_as_float(0x7fc00000);
                long long dS = 0;
                if constexpr (kI4) {
                    uint32_t ds = dsum[mp];
                    ds += __shfl_xor_sync(0xffffffffu, ds, 1);
                    ds += __shfl_xor_sync(0xffffffffu, ds, 2);
                    const uint32_t gsrc = 4 * (t >> 1);   // digit columns of the row this lane writes
                    const uint32_t D0 = __shfl_sync(0xffffffffu, ds, 4 * (gsrc + 0));
                    const uint32_t D1 = __shfl_sync(0xffffffffu, ds, 4 * (gsrc + 1));
                    const uint32_t D2 = __shfl_sync(0xffffffffu, ds, 4 * (gsrc + 2));
                    dS = (long long)D0 + ((long long)D1 << 8) + ((long long)D2 << 16);
                }
#pragma unroll
                for (int ct = 0; ct < 2; ct++) {
                    int pz[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) pz[i] = __shfl_xor_sync(0xffffffffu, acc[mp][ct][i], 1);
                    if ((t & 1) == 0) {
                        const int m = 2 * mp + (int)(t >> 1);
                        const int* o = acc[mp][ct];
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            long long T;
                            if constexpr (!kI4) {
                                // own columns 2t, 2t+1 = digits 0, 1; partner's = digit 2 and sum_k q; row g + 8h
                                T = (long long)o[2 * h] + ((long long)o[2 * h + 1] << 8) + ((long long)pz[2 * h] << 16) -
                                    12582912LL * (long long)pz[2 * h + 1];
                            } else {
                                // Y (row g + 8) = sum u[n] d ; X (row g) = sum (16 u[n+8] + u[n]) d
                                long long u0, u1, u2, us;
                                if (h == 0) { u0 = o[2]; u1 = o[3]; u2 = pz[2]; us = pz[3]; }
                                else { u0 = (o[0] - o[2]) >> 4; u1 = (o[1] - o[3]) >> 4; u2 = (pz[0] - pz[2]) >> 4; us = (pz[1] - pz[3]) >> 4; }
                                // sum (u - 8)(Mk - 3*2^22) = sum u Mk - 8 sum Mk - 3*2^22 (sum u - 8 kcnt)
                                T = u0 + (u1 << 8) + (u2 << 16) - 8 * dS - 12582912LL * (us - 8LL * (long long)kcnt);
                            }
                            dstp[m * ZG_TN + ct * 16 + g + 8 * h] = __ll2float_rn(T) * esc;
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ── combine the warps in fixed order; S == 1: final result, else partial of this split ──
        if (tid < MR * ZG_TN) {
            const uint32_t m = tid >> 5, col = tid & 31;
            float v = 0.0f;
            for (uint32_t w = 0; w < W; w++) v += part[buf][w][m][col];
            if (p.S == 1) {
                if (m < p.M) p.out[(size_t)m * p.out_rs + nb * ZG_TN + col] = v;
            } else if (p.cl) {
                xpart[buf][m][col] = v;
            } else {
                p.partials[(((size_t)nb * p.S + split) * MR + m) * ZG_TN + col] = v;
            }
        }
        if (p.S > 1 && p.cl) {
            // the S splits of this column group are the CTAs of one cluster (rank = split): one cluster barrier publishes the
            // partials, the split-0 CTA adds them in split order through distributed shared memory — no global scratch, no atomic.
            // The next group uses the other buffer, and rank 0 reaches that group's barrier only after this read.
            cluster_sync_all();
            if (split == 0 && tid < MR * ZG_TN) {
                const uint32_t m = tid >> 5, col = tid & 31;
                float v = 0.0f;
                for (uint32_t s2 = 0; s2 < p.S; s2++) v += ld_cluster_f32(&xpart[buf][m][col], s2);
                if (m < p.M) p.out[(size_t)m * p.out_rs + nb * ZG_TN + col] = v;
            }
        } else if (p.S > 1) {
            __syncthreads();
            if (tid == 0) {
                // release: publishes the whole CTA's partials (ordered before this thread by the barrier above);
                // acquire: the last arriver sees every earlier contributor's partials
                uint32_t old;
                asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(p.counters + nb) : "memory");
                const uint32_t last = (old == p.S - 1) ? 1u : 0u;
                if (last) p.counters[nb] = 0u;  // re-arm for the next launch
                s_last = last;
            }
            __syncthreads();
            if (s_last && tid < MR * ZG_TN) {
                const uint32_t m = tid >> 5, col = tid & 31;
                float v = 0.0f;
                for (uint32_t s2 = 0; s2 < p.S; s2++)
                    v += __ldcg(p.partials + (((size_t)nb * p.S + s2) * MR + m) * ZG_TN + col);
                if (m < p.M) p.out[(size_t)m * p.out_rs + nb * ZG_TN + col] = v;
            }
        }
        buf ^= 1;
    }
    if (p.S > 1 && p.cl) cluster_sync_all();   // nobody leaves while the split-0 CTA may still read its partials
    ZG_TRACE_MARK(2)
}

// ── gate | up pair with the activation epilogue (M == 1) ─────────────────────────────────────────────────────────────
// The decode matvec above, specialised to one activation row, walking its column groups twice per group — first the
// `a` weight (gate), then the `b` weight (up) — through the same TMA ring and the same staged activations.  When the second
// sum of a column group is final (S == 1: right away; S > 1: in the CTA that arrives last for `b` — every CTA passes `a`
// before `b`, so `a`'s final values are visible to it), the epilogue evaluates mid = steps(gate), dst = mid * up.
struct QGemvPair { QGemvParams a, b; ZgGemvEpilogue epi; };
struct QGemvPairBatch { QGemvPair p[2]; };

__device__ __forceinline__ float epi_unary(uint32_t op, float v) {
    switch (op) {
        case ZG_EW_NEG: return -v;
        case ZG_EW_ABS: return fabsf(v);
        case ZG_EW_RELU: return fmaxf(v, 0.0f);
        case ZG_EW_SQRT: return sqrtf(v);
        case ZG_EW_RECIP: return 1.0f / v;
        case ZG_EW_EXP: return expf(v);
        case ZG_EW_LOG: return logf(v);
        default: return v;
    }
}

template <int FMT>
__global__ void __launch_bounds__(kThreads, 3)
qgemv_pair_kernel(const __grid_constant__ QGemvPairBatch bt) {
    const QGemvPair& pp = bt.p[blockIdx.y];
    const QGemvParams& pa = pp.a;
    const QGemvParams& pb = pp.b;
    constexpr bool kI4 = (FMT == ZG_QFMT_I4_F16);
    constexpr bool kF32 = (FMT == ZG_QFMT_I8_F32);
    constexpr uint32_t QB = kI4 ? 512u : 1024u;
    constexpr uint32_t SB = kF32 ? 32u : 16u;
    constexpr uint32_t RB = QB + 4 * SB;
    constexpr uint32_t G = kI4 ? 4u : 2u;

    __shared__ float part[2][kMaxWarps][ZG_TN];
    __shared__ float xpart[2][ZG_TN];   // cluster split-K (see qgemv_kernel)
    __shared__ uint32_t s_last;
    extern __shared__ __align__(128) uint8_t dsm[];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    const uint32_t g = lane >> 2, t = lane & 3, j = g & 3;
    const uint32_t split = blockIdx.x % pa.S, grp = blockIdx.x / pa.S;
    const uint32_t nb_begin = grp * pa.P;
    const uint32_t nb_end = min(nb_begin + pa.P, pa.n_nb);
    const uint32_t ks = (uint32_t)(((uint64_t)split * pa.n_kc) / pa.S);
    const uint32_t ke = (uint32_t)(((uint64_t)(split + 1) * pa.n_kc) / pa.S);
    const uint32_t k0 = ks + (warp * (ke - ks)) / W, k1 = ks + ((warp + 1) * (ke - ks)) / W;
    const uint32_t L = k1 - k0;
    const uint32_t NS = pa.NS;
    const uint32_t n_chunk = (L + G - 1) / G;
    const uint32_t n_virtual = 2 * (nb_end - nb_begin);          // (column group, weight) pairs this CTA walks: a, b, a, b, ...
    const uint32_t total_chunks = n_chunk * n_virtual;
    constexpr uint32_t slot_bytes = G * RB;

    const uint32_t dsm_u32 = smem_u32(dsm);
    const uint32_t xs_bytes = W * pa.xs_stride * 4, ring_bytes = W * NS * slot_bytes, plane_bytes = G * kPlaneRow;
    float* xs_w = reinterpret_cast<float*>(dsm) + (size_t)warp * pa.xs_stride;
    const uint32_t xs_u32 = dsm_u32 + warp * pa.xs_stride * 4;
    const uint32_t ring = dsm_u32 + xs_bytes + warp * NS * slot_bytes;
    const uint32_t planes = dsm_u32 + xs_bytes + ring_bytes + warp * plane_bytes;
    const uint32_t bars = dsm_u32 + xs_bytes + ring_bytes + W * plane_bytes + warp * NS * 8;

    ZG_TRACE_BEGIN(10)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    uint32_t pf_c = 0, pf_v = 0, pf_slot = 0, pf_left = total_chunks;
    auto issue_chunk = [&]() {   // lane 0 only
        const uint8_t* base = ((pf_v & 1) ? pb.recs : pa.recs) + ((size_t)(nb_begin + (pf_v >> 1)) * pa.n_kc + k0) * RB;
        const uint32_t cnt = min(G, L - pf_c * G);
        const uint32_t bar = bars + pf_slot * 8;
        mbar_expect_tx(bar, cnt * RB);
        bulk_g2s(ring + pf_slot * slot_bytes, base + (size_t)pf_c * G * RB, cnt * RB, bar);
        pf_left--;
        if (++pf_c == n_chunk) { pf_c = 0; pf_v++; }
        if (++pf_slot == NS) pf_slot = 0;
    };
    if (lane == 0) {
        for (uint32_t s = 0; s < NS; s++) mbar_init(bars + s * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (uint32_t i = 0; i < NS && pf_left; i++) issue_chunk();
    }
    for (uint32_t i = lane; i < G * 8; i += 32) {
        const uint32_t rm = i >> 3, w4 = i & 7;
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(planes + rm * kPlaneRow + 96 + w4 * 4), "r"(0x01010101u) : "memory");
    }
    __syncwarp();
    float sm_prev = __ldg(pa.smax + nb_begin);

    asm volatile("griddepcontrol.wait;" ::: "memory");
    ZG_TRACE_MARK(1)

    float xm = 0.0f;
    {
        const uint32_t kb = k0 * ZG_KR;
        const float rsm = 1.0f / sm_prev;
        const float* xr = pa.x + kb + lane;
        float* xd = xs_w + lane;
        float v[kLcap];
        float mx = 0.0f;
#pragma unroll
        for (int i = 0; i < (int)kLcap; i++) {
            v[i] = ((uint32_t)i < L && kb + lane + 32 * i < pa.K) ? xr[32 * i] : 0.0f;
            const float aa = fabsf(v[i]);
            mx = (aa <= 3.0e38f) ? fmaxf(mx, aa) : INFINITY;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        xm = mx;
        const float f = (mx <= 3.0e38f && mx >= 1.0e-30f) ? (0.499f / mx) * rsm : 0.0f;
#pragma unroll
        for (int i = 0; i < (int)kLcap; i++)
            if ((uint32_t)i < L) xd[32 * i] = v[i] * f;
        __syncwarp();
    }
    const uint32_t brow = planes + j * 32 + 4 * t;
    const uint32_t q_off = lane * 16;
    const uint32_t sc_off = QB + (kF32 ? 4u : 2u) * (8 * ((lane & 15) >> 2) + 4 * (lane >> 4) + (lane & 3));

    uint32_t buf = 0, slot = 0, parity = 0, slot_u32 = ring, bar_u32 = bars;
    float v_a = 0.0f;   // tid < 32: the finished `a` sum of the current column group (S == 1)

    for (uint32_t vi = 0; vi < n_virtual; vi++) {
        const uint32_t nb = nb_begin + (vi >> 1);
        const bool is_b = (vi & 1) != 0;
        const QGemvParams& cur = is_b ? pb : pa;
        const float sm = __ldg(cur.smax + nb);
        if (sm != sm_prev) {
            const float ratio = sm_prev / sm;
            float* xd = xs_w + lane;
            for (uint32_t i = 0; i < L; i++) xd[32 * i] *= ratio;
            __syncwarp();
            sm_prev = sm;
        }
        int acc[2][4];
        uint32_t dsum = 0;
#pragma unroll
        for (int ct = 0; ct < 2; ct++)
#pragma unroll
            for (int i = 0; i < 4; i++) acc[ct][i] = 0;
        uint32_t xa = xs_u32 + lane * 4;
        for (uint32_t c = 0; c < n_chunk; c++, xa += G * ZG_KR * 4) {
            const uint32_t cnt = min(G, L - c * G);
            mbar_wait(bar_u32, parity);
#pragma unroll
            for (uint32_t r = 0; r < G; r++) {
                if (r < cnt) {
                    float sc;
                    if constexpr (kF32) {
                        sc = __uint_as_float(lds32(slot_u32 + sc_off + r * RB));
                    } else {
                        unsigned short h;
                        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(slot_u32 + sc_off + r * RB));
                        sc = __half2float(__ushort_as_half(h));
                    }
                    const uint32_t F = __float_as_uint(fmaf(sc, __uint_as_float(lds32(xa + r * ZG_KR * 4)), 1.5f));
                    const uint32_t pa2 = planes + lane + r * kPlaneRow;
                    sts8(pa2, F);
                    sts8(pa2 + 32, F >> 8);
                    sts8(pa2 + 64, F >> 16);
                }
            }
            __syncwarp();
#pragma unroll
            for (uint32_t r = 0; r < G; r++) {
                if (r < cnt) {
                    const uint32_t qa = slot_u32 + q_off + r * RB;
                    const uint32_t b0 = lds32(brow + r * kPlaneRow), b1 = lds32(brow + r * kPlaneRow + 16);
                    if constexpr (!kI4) {
                        const uint4 q0 = lds128(qa), q1 = lds128(qa + 512);
                        imma_s8u8(acc[0], q0.x, q0.y, q0.z, q0.w, b0, b1);
                        imma_s8u8(acc[1], q1.x, q1.y, q1.z, q1.w, b0, b1);
                    } else {
                        const uint4 q0 = lds128(qa);
                        dsum = __dp4a(b0, 0x01010101u, __dp4a(b1, 0x01010101u, dsum));
                        imma_u8u8(acc[0], q0.x, q0.x & 0x0F0F0F0Fu, q0.y, q0.y & 0x0F0F0F0Fu, b0, b1);
                        imma_u8u8(acc[1], q0.z, q0.z & 0x0F0F0F0Fu, q0.w, q0.w & 0x0F0F0F0Fu, b0, b1);
                    }
                }
            }
            __syncwarp();
            if (lane == 0 && pf_left) issue_chunk();
            slot_u32 += slot_bytes; bar_u32 += 8;
            if (++slot == NS) { slot = 0; slot_u32 = ring; bar_u32 = bars; parity ^= 1; }
        }
        {   // flush: integer sums -> this warp's float partial
            const uint32_t kcnt = L * ZG_KR;
            float* dstp = &part[buf][warp][0];
            const float esc = (xm <= 3.0e38f) ? ((xm >= 1.0e-30f) ? (xm * (2.004008016f * 1.1920928955078125e-07f)) * sm : 0.0f)
                                              : __int_as_float(0x7fc00000);
            long long dS = 0;
            if constexpr (kI4) {
                uint32_t ds = dsum;
                ds += __shfl_xor_sync(0xffffffffu, ds, 1);
                ds += __shfl_xor_sync(0xffffffffu, ds, 2);
                const uint32_t D0 = __shfl_sync(0xffffffffu, ds, 0);
                const uint32_t D1 = __shfl_sync(0xffffffffu, ds, 4);
                const uint32_t D2 = __shfl_sync(0xffffffffu, ds, 8);
                dS = (long long)D0 + ((long long)D1 << 8) + ((long long)D2 << 16);
            }
#pragma unroll
            for (int ct = 0; ct < 2; ct++) {
                int pz[4];
#pragma unroll
                for (int i = 0; i < 4; i++) pz[i] = __shfl_xor_sync(0xffffffffu, acc[ct][i], 1);
                if (t == 0) {
                    const int* o = acc[ct];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        long long T;
                        if constexpr (!kI4) {
                            T = (long long)o[2 * h] + ((long long)o[2 * h + 1] << 8) + ((long long)pz[2 * h] << 16) - 12582912LL * (long long)pz[2 * h + 1];
                        } else {
                            long long u0, u1, u2, us;
                            if (h == 0) { u0 = o[2]; u1 = o[3]; u2 = pz[2]; us = pz[3]; }
                            else { u0 = (o[0] - o[2]) >> 4; u1 = (o[1] - o[3]) >> 4; u2 = (pz[0] - pz[2]) >> 4; us = (pz[1] - pz[3]) >> 4; }
                            T = u0 + (u1 << 8) + (u2 << 16) - 8 * dS - 12582912LL * (us - 8LL * (long long)kcnt);
                        }
                        dstp[ct * 16 + g + 8 * h] = __ll2float_rn(T) * esc;
                    }
                }
            }
        }
        __syncthreads();
        float fin = 0.0f;
        bool have = false;
        const uint32_t n = nb * ZG_TN + tid;
        if (tid < ZG_TN) {
            float v = 0.0f;
            for (uint32_t w = 0; w < W; w++) v += part[buf][w][tid];
            if (cur.S == 1) { fin = v; have = true; }
            else if (cur.cl) xpart[buf][tid] = v;
            else cur.partials[(((size_t)nb * cur.S + split) * 2) * ZG_TN + tid] = v;
        }
        if (cur.S > 1 && cur.cl) {
            cluster_sync_all();
            if (split == 0 && tid < ZG_TN) {
                float v = 0.0f;
                for (uint32_t s2 = 0; s2 < cur.S; s2++) v += ld_cluster_f32(&xpart[buf][tid], s2);
                fin = v; have = true;
            }
        } else if (cur.S > 1) {
            __syncthreads();
            if (tid == 0) {
                uint32_t old;
                asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(cur.counters + nb) : "memory");
                const uint32_t last = (old == cur.S - 1) ? 1u : 0u;
                if (last) cur.counters[nb] = 0u;
                s_last = last;
            }
            __syncthreads();
            if (s_last && tid < ZG_TN) {
                float v = 0.0f;
                for (uint32_t s2 = 0; s2 < cur.S; s2++) v += __ldcg(cur.partials + (((size_t)nb * cur.S + s2) * 2) * ZG_TN + tid);
                fin = v; have = true;
            }
        }
        if (have) {
            cur.out[n] = fin;
            if (!is_b) v_a = fin;
            else {
                const float gt = (pa.S == 1 || pa.cl) ? v_a : __ldcg(pa.out + n), up = fin;   // cluster split-K: the same (split-0) CTA finishes both
                float v = gt;
                for (uint32_t st = 0; st < pp.epi.n_steps; st++) {
                    const uint32_t sop = pp.epi.steps[st].op, sw = pp.epi.steps[st].is_swapped;
                    if (sop == ZG_EW_ADD || sop == ZG_EW_MUL) {
                        const uint32_t kd = pp.epi.sec_kind[st];
                        const float o2 = kd == 1 ? gt : (kd == 2 ? up : pp.epi.steps[st].sec[n]);
                        if (sop == ZG_EW_ADD) v = sw ? __fadd_rn(o2, v) : __fadd_rn(v, o2);
                        else v = sw ? __fmul_rn(o2, v) : __fmul_rn(v, o2);
                    } else v = epi_unary(sop, v);
                }
                pp.epi.o_mid[n] = v;
                pp.epi.o_dst[n] = __fmul_rn(v, up);
            }
        }
        buf ^= 1;
    }
    if (pa.S > 1 && pa.cl) cluster_sync_all();   // nobody leaves while the split-0 CTA may still read its partials
    ZG_TRACE_MARK(2)
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

template <int FMT, int MP, int XR, int PRO = 0>
bool launch_fast(const ZgGemvPlan& plan, const QGemvBatch& p, uint32_t count, cudaStream_t st, bool pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(plan.grid, count);
    cfg.blockDim = dim3(plan.threads);
    cfg.dynamicSmemBytes = plan.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    unsigned na = 0;
    if (pdl) { attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; na++; }
    if (p.p[0].cl) {   // the k-splits of a column group = one cluster (consecutive blockIdx.x)
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = plan.S; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1; na++;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    cudaError_t e;
    if constexpr (PRO != 0) {
        e = cudaLaunchKernelEx(&cfg, qgemv_kernel<FMT, MP, XR, PRO>, p);
    } else {
        QGemvBatchT<false> plain;
        for (uint32_t i = 0; i < kZgGemvBatch; i++) plain.p[i] = static_cast<const QGemvParams&>(p.p[i]);
        e = cudaLaunchKernelEx(&cfg, qgemv_kernel<FMT, MP, XR, 0>, plain);
    }
    ZG_COUNT_LAUNCH();
    if (e != cudaSuccess) {
        zg_set_error("qgemv launch failed: %s (grid %u, %u B shared)", cudaGetErrorString(e), plan.grid, plan.smem_bytes);
        return false;
    }
    return true;
}

template <int FMT>
bool launch_fmt(const ZgGemvPlan& plan, const QGemvBatch& p, uint32_t count, cudaStream_t st, bool pdl) {
    const uint32_t pk = p.p[0].pro.kind;
    if (pk != 0) {   // prologue recipes exist for the decode variants only (M <= 2)
        if (plan.mp != 1) { zg_set_error("internal: matvec prologue with more than 2 rows"); return false; }
        if (pk == 1) return p.p[0].M == 1 ? launch_fast<FMT, 1, 1, 1>(plan, p, count, st, pdl) : launch_fast<FMT, 1, 2, 1>(plan, p, count, st, pdl);
        return p.p[0].M == 1 ? launch_fast<FMT, 1, 1, 2>(plan, p, count, st, pdl) : launch_fast<FMT, 1, 2, 2>(plan, p, count, st, pdl);
    }
    switch (plan.mp) {
        case 1: return p.p[0].M == 1 ? launch_fast<FMT, 1, 1>(plan, p, count, st, pdl) : launch_fast<FMT, 1, 2>(plan, p, count, st, pdl);
        case 2: return launch_fast<FMT, 2, 4>(plan, p, count, st, pdl);
        case 4: return launch_fast<FMT, 4, 8>(plan, p, count, st, pdl);
        default: zg_set_error("qmatmul: bad plan (row pairs %u)", plan.mp); return false;
    }
}

} // namespace

// Work split for one launch of up to 8 activation rows.
ZgGemvPlan zg_qgemv_plan(const ZgCudaCtx* ctx, const ZgCudaQWeight* w, uint32_t M, uint32_t count) {
    ZgGemvPlan pl;
    const uint32_t rows = M > 8 ? 8 : M;
    pl.mp = rows <= 2 ? 1 : (rows <= 4 ? 2 : 4);
    pl.threads = kThreads;
    const uint32_t occ = pl.mp == 1 ? 3 : (pl.mp == 2 ? 2 : 1);
    const uint32_t target = (uint32_t)ctx->sm_count * occ;   // CTAs resident at once
    const uint32_t warps = kThreads / 32;
    // k-splits: each warp stages its k-range of the activations in shared memory -> at most lcap_max records per
    // warp.  Beyond that K is split across CTAs only until there is one CTA per SM (>= 4 records per warp): CTAs
    // that live long amortise their fixed latencies (activation staging, flush, reduction) and leave room for the
    // next kernel's CTAs, which is what keeps HBM busy across kernel boundaries.
    const uint32_t lcap_max = kLcap / pl.mp;   // staged activations stay <= 32 KB per CTA
    uint32_t S = (w->n_kc + warps * lcap_max - 1) / (warps * lcap_max);
    if (w->n_nb * S * 2 <= (uint32_t)ctx->sm_count) {
        uint32_t fill = (uint32_t)ctx->sm_count / w->n_nb;
        const uint32_t by_work = w->n_kc / (warps * 4);
        if (fill > by_work) fill = by_work;
        if (fill > S) S = fill;
    }
    if (ctx->tune_s && (uint32_t)ctx->tune_s > S) S = (uint32_t)ctx->tune_s;
    if (ctx->tune_smax && (uint32_t)ctx->tune_smax < S && (w->n_kc + ctx->tune_smax * warps * lcap_max - 1) / (ctx->tune_smax * warps * lcap_max) <= 1) S = (uint32_t)ctx->tune_smax;
    if (S > w->n_kc) S = w->n_kc;
    if (S < 1) S = 1;
    // Wave quantisation: `count` matvecs share the launch, and a launch of slightly more CTAs than the GPU holds at once
    // (Llama-3-70B gate|up sharded over 8 GPUs: 2 x 112 column groups x 2 splits = 448 CTAs for 444 slots) costs an
    // extra round.  ZG_GEMV_WAVE=1 tries a few more k-splits and takes the cheapest in rounds x (records per warp + fixed
    // per-CTA cost).  MEASURED AND REJECTED as a default (round 2, same box A/B): 2 % slower on the 70B decode step both
    // unsharded (252.4 -> 247.2 tok/s at 24 layers) and for an 8-way shard (58.9 -> 60.5 us/layer) — every extra split adds
    // a partial-sum round trip through global scratch that costs more than the emptier last round saves.
    static const bool wave_aware = [] { const char* e = getenv("ZG_GEMV_WAVE"); return e && e[0] == '1'; }();
    if (wave_aware && S >= 2 && !ctx->tune_s && count >= 1 && (uint64_t)w->n_nb * S * count > target) {
        const double c0 = 6.0;
        double best = 1e30; uint32_t bS = S;
        for (uint32_t s2 = S; s2 <= S + 4 && s2 <= w->n_kc; s2++) {
            const uint64_t ctas = (uint64_t)w->n_nb * s2 * count;
            const double rounds = (double)((ctas + target - 1) / target);
            const double per_warp = (double)(((w->n_kc + s2 - 1) / s2 + warps - 1) / warps);
            const double cost = rounds * (per_warp + c0);
            if (cost < best - 1e-9) { best = cost; bS = s2; }
        }
        S = bS;
    }
    uint32_t P = ((uint64_t)w->n_nb * S + target - 1) / target;
    if (P < 1) P = 1;
    if (ctx->tune_p) P = (uint32_t)ctx->tune_p;
    pl.P = P; pl.S = S;
    const uint32_t len_max = (w->n_kc + S - 1) / S;
    pl.lcap = (len_max + warps - 1) / warps;
    // weight ring per warp: NS slots of G records (one TMA bulk copy each), ~4.5 KB per warp
    const uint32_t rb = w->rec_bytes;
    pl.G = w->fmt == ZG_QFMT_I4_F16 ? 4 : 2;   // compile-time constant of the kernel
    // 2 slots keep a decode CTA under 70 KB of shared memory: THREE CTAs per SM stay resident (the limit the 80 registers
    // allow), which is worth more than a deeper ring - short-lived CTAs are latency-bound, bytes in flight come from CTA count
    pl.NS = ctx->tune_u ? (uint32_t)ctx->tune_u : 2;
    pl.xs_stride = pl.lcap * ZG_KR;
    const uint32_t xrows = rows == 1 ? 1 : 2 * pl.mp;
    const uint32_t chunks = ((pl.lcap + pl.G - 1) / pl.G) * P;   // chunks a warp ever requests
    if (pl.NS > chunks) pl.NS = chunks < 1 ? 1 : chunks;
    auto smem_of = [&]() {
        return warps * xrows * pl.xs_stride * 4 + warps * pl.NS * pl.G * rb + warps * pl.G * xrows * kPlaneRow + warps * pl.NS * 8;
    };
    while (smem_of() > 200 * 1024 && pl.NS > 1) pl.NS--;
    pl.smem_bytes = smem_of();
    pl.grid = ((w->n_nb + P - 1) / P) * S;
    return pl;
}

template <int FMT, int MP, int XR, int PRO = 0>
bool set_smem_attr() {
    cudaError_t e = cudaFuncSetAttribute(qgemv_kernel<FMT, MP, XR, PRO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { zg_set_error("cudaFuncSetAttribute(qgemv) failed: %s", cudaGetErrorString(e)); return false; }
    return true;
}
template <int FMT>
bool set_smem_attrs() {
    return set_smem_attr<FMT, 1, 1>() && set_smem_attr<FMT, 1, 2>() && set_smem_attr<FMT, 2, 4>() && set_smem_attr<FMT, 4, 8>() &&
           set_smem_attr<FMT, 1, 1, 1>() && set_smem_attr<FMT, 1, 2, 1>() && set_smem_attr<FMT, 1, 1, 2>() && set_smem_attr<FMT, 1, 2, 2>();
}

bool zg_qgemv_init(ZgCudaCtx* ctx) {
    if (const char* e = getenv("ZG_GEMV_S")) ctx->tune_s = atoi(e);
    if (const char* e = getenv("ZG_GEMV_P")) ctx->tune_p = atoi(e);
    if (const char* e = getenv("ZG_GEMV_SMAX")) ctx->tune_smax = atoi(e);
    if (const char* e = getenv("ZG_GEMV_NS")) ctx->tune_u = atoi(e);
    if (const char* e = getenv("ZG_GEMV_CLUSTER")) ctx->gemv_cluster = (e[0] != '0');   // 0: split-K through global scratch + arrival counters
    if (const char* e = getenv("ZG_GEMV_G")) ctx->tune_g = atoi(e);
    if (const char* e = getenv("ZG_GEMV_ROWS")) { ctx->tune_rows = atoi(e); if (ctx->tune_rows < 1 || ctx->tune_rows > 8) ctx->tune_rows = 0; }
    if (ctx->tune_u < 2 || ctx->tune_u > 16) ctx->tune_u = 0;
    if (ctx->tune_g < 1 || ctx->tune_g > 16) ctx->tune_g = 0;
    // opt every instantiation into its dynamic shared memory once per context, outside any stream capture
    return set_smem_attrs<ZG_QFMT_I8_F32>() && set_smem_attrs<ZG_QFMT_I8_F16>() && set_smem_attrs<ZG_QFMT_I4_F16>() && zg_qgemv_stream_init(ctx);
}

void zg_qgemv_ws_need(const ZgCudaCtx* ctx, const ZgCudaQWeight* w, uint32_t M, size_t* partial_elems,
                      size_t* counters) {
    *partial_elems = 0; *counters = 0;
    if (w->fmt == ZG_QFMT_GENERIC || M == 0 || M > 8) return;
    uint32_t S = 1, mp = 1;
    for (uint32_t count = 1; count <= kZgGemvBatch; count++) {   // the split count depends on how many matvecs share the launch: size for the largest
        const ZgGemvPlan plan = zg_qgemv_plan(ctx, w, M, count);
        if (plan.S > S) S = plan.S;
        mp = plan.mp;
    }
    if (S > 1) {
        *partial_elems = (size_t)w->n_nb * S * (2 * mp) * ZG_TN;
        *counters = w->n_nb;
    }
    if (M == 1)   // the streamed form cuts column groups at CTA boundaries: one partial slot per piece
        for (uint32_t count = 1; count <= kZgGemvBatch; count++) {
            const ZgGemvStreamPlan sp = zg_qgemv_stream_plan(ctx, w, count);
            if (!sp.use) continue;
            *partial_elems = std::max(*partial_elems, (size_t)w->n_nb * sp.slots * ZG_TN);
            *counters = w->n_nb;
        }
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
    if (M > 8) {   // prefill: dense contraction on the tcgen05 tensor cores (qgemm.cu)
        const size_t need = zg_qgemm_scratch_elems(w, M);
        if (!ws || ws->gemm_scratch_elems < need) { zg_set_error("internal: GEMM scratch too small (%zu needed)", need); return false; }
        return zg_qgemm_launch(ctx, w, d_in, d_out, M, in_rs, out_rs, ws->gemm_scratch, st);
    }
    if (M <= 8) return zg_qgemv_launch_batch(ctx, 1, &w, &d_in, &d_out, M, &in_rs, &out_rs, ws, st);
    // M > 8 in the generic-format-free matvec path never happens (the GEMM takes it); kept for direct callers
    for (uint32_t m0 = 0; m0 < M; m0 += 8) {
        const uint32_t rows = (M - m0) > 8 ? 8 : (M - m0);
        const float* xin = d_in + (size_t)m0 * in_rs;
        float* xout = d_out + (size_t)m0 * out_rs;
        if (!zg_qgemv_launch_batch(ctx, 1, &w, &xin, &xout, rows, &in_rs, &out_rs, ws, st)) return false;
    }
    return true;
}

// `count` (<= kZgGemvBatch) matvecs with identical weight shape, format and row count M <= 8 in one launch per pass of
// `rows_per_pass` activation rows (ZG_GEMV_ROWS: the wide 8-row variant keeps one CTA per SM and short k-ranges; fewer
// rows per pass re-stream the weights but run the well-occupied narrow variants).
static bool launch_batch_rows(ZgCudaCtx* ctx, uint32_t count, const ZgCudaQWeight* const* ws_w, const float* const* d_in,
                              float* const* d_out, uint32_t M, const uint32_t* in_rs, const uint32_t* out_rs,
                              const ZgGemvWs* ws, cudaStream_t st, const ZgGemvPrologue* pro) {
    const ZgCudaQWeight* w0 = ws_w[0];
    if (M == 1 && !(pro && pro[0].kind)) {
        const ZgGemvStreamPlan sp = zg_qgemv_stream_plan(ctx, w0, count);
        if (sp.use) return zg_qgemv_stream_launch(ctx, sp, count, ws_w, d_in, d_out, ws, st);
    }
    ZgGemvPlan plan = zg_qgemv_plan(ctx, w0, M, count);
    QGemvBatch bt;
    memset(&bt, 0, sizeof(bt));
    for (uint32_t i = 0; i < count; i++) {
        const ZgCudaQWeight* w = ws_w[i];
        if (w->fmt != w0->fmt || w->K != w0->K || w->N != w0->N || w->fmt == ZG_QFMT_GENERIC) { zg_set_error("internal: mixed matvec batch"); return false; }
        size_t pe = 0, nc = 0;
        zg_qgemv_ws_need(ctx, w, M, &pe, &nc);
        if (pe && (ws[i].partials_elems < pe || ws[i].counters_n < nc)) {
            zg_set_error("internal: split workspace too small (%zu/%zu needed)", pe, nc);
            return false;
        }
        QGemvParamsPro& p = bt.p[i];
        p.recs = w->recs; p.smax = w->smax;
        p.n_kc = w->n_kc; p.n_nb = w->n_nb;
        p.K = (uint32_t)w->K; p.N = (uint32_t)w->N; p.M = M;
        p.x = d_in[i]; p.x_rs = in_rs[i] ? in_rs[i] : (uint32_t)w->K;
        p.xs_stride = plan.xs_stride;
        p.out = d_out[i]; p.out_rs = out_rs[i] ? out_rs[i] : (uint32_t)w->N;
        p.P = plan.P; p.S = plan.S; p.NS = plan.NS;
        p.cl = (ctx->gemv_cluster && (plan.S == 2 || plan.S == 4 || plan.S == 8)) ? 1u : 0u;
        p.partials = ws[i].partials; p.counters = ws[i].counters;
        if (pro) {
            if (pro[i].kind != pro[0].kind || (pro[i].kind && (M > 2 || p.x_rs != p.K))) { zg_set_error("internal: mixed or unsupported matvec prologue"); return false; }
            p.pro = pro[i];
        }
    }
    switch (w0->fmt) {
        case ZG_QFMT_I8_F32: return launch_fmt<ZG_QFMT_I8_F32>(plan, bt, count, st, ctx->pdl);
        case ZG_QFMT_I8_F16: return launch_fmt<ZG_QFMT_I8_F16>(plan, bt, count, st, ctx->pdl);
        case ZG_QFMT_I4_F16: return launch_fmt<ZG_QFMT_I4_F16>(plan, bt, count, st, ctx->pdl);
        default: zg_set_error("qmatmul: unknown weight format %d", w0->fmt); return false;
    }
}

bool zg_qgemv_launch_batch(ZgCudaCtx* ctx, uint32_t count, const ZgCudaQWeight* const* ws_w, const float* const* d_in,
                           float* const* d_out, uint32_t M, const uint32_t* in_rs, const uint32_t* out_rs,
                           const ZgGemvWs* ws, cudaStream_t st, const ZgGemvPrologue* pro) {
    if (count == 0 || M == 0) return true;
    if (count > kZgGemvBatch || M > 8) { zg_set_error("internal: bad matvec batch (%u ops, %u rows)", count, M); return false; }
    const ZgGemvWs none[kZgGemvBatch] = {};
    if (!ws) ws = none;
    const uint32_t rpp = ctx->tune_rows ? (uint32_t)ctx->tune_rows : 4u;   // measured: two 4-row passes beat one 8-row pass (ZG_GEMV_ROWS)
    if (M <= rpp) return launch_batch_rows(ctx, count, ws_w, d_in, d_out, M, in_rs, out_rs, ws, st, pro);
    if (pro && pro[0].kind) { zg_set_error("internal: matvec prologue on a multi-pass launch"); return false; }
    for (uint32_t m0 = 0; m0 < M; m0 += rpp) {
        const float* xin[kZgGemvBatch]; float* xout[kZgGemvBatch];
        for (uint32_t i = 0; i < count; i++) {
            xin[i] = d_in[i] + (size_t)m0 * (in_rs[i] ? in_rs[i] : (uint32_t)ws_w[i]->K);
            xout[i] = d_out[i] + (size_t)m0 * (out_rs[i] ? out_rs[i] : (uint32_t)ws_w[i]->N);
        }
        if (!launch_batch_rows(ctx, count, ws_w, xin, xout, std::min(rpp, M - m0), in_rs, out_rs, ws, st, nullptr)) return false;
    }
    return true;
}


// gate | up pair with the activation epilogue: see qgemv_pair_kernel.  M == 1, identical fast-format shapes, one input vector.
template <int FMT>
static bool launch_pair_fmt(const ZgGemvPlan& plan, const QGemvPairBatch& bt, cudaStream_t st, bool pdl) {
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(qgemv_pair_kernel<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) { zg_set_error("cudaFuncSetAttribute(qgemv pair) failed"); return false; }
        attr_done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(plan.grid, 1);
    cfg.blockDim = dim3(plan.threads);
    cfg.dynamicSmemBytes = plan.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    unsigned na = 0;
    if (pdl) { attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; na++; }
    if (bt.p[0].a.cl) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = plan.S; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1; na++;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, qgemv_pair_kernel<FMT>, bt);
    ZG_COUNT_LAUNCH();
    if (e != cudaSuccess) { zg_set_error("qgemv pair launch failed: %s", cudaGetErrorString(e)); return false; }
    return true;
}

bool zg_qgemv_launch_pair(ZgCudaCtx* ctx, const ZgCudaQWeight* wa, const ZgCudaQWeight* wb, const float* d_in, float* d_out_a, float* d_out_b,
                          const ZgGemvWs* ws_a, const ZgGemvWs* ws_b, const ZgGemvEpilogue& epi, cudaStream_t st) {
    if (wa->fmt != wb->fmt || wa->K != wb->K || wa->N != wb->N || wa->fmt == ZG_QFMT_GENERIC) { zg_set_error("internal: mismatched matvec pair"); return false; }
    // the pair is planned like a batch of two: same column groups and k-splits, each CTA walks both weights
    ZgGemvPlan plan = zg_qgemv_plan(ctx, wa, 1, 2);
    {   // each warp requests twice as many chunks: the ring depth may use them
        const uint32_t chunks = ((plan.lcap + plan.G - 1) / plan.G) * plan.P * 2;
        static const uint32_t pair_ns = [] { const char* e = getenv("ZG_GEMV_PAIR_NS"); const int v = e ? atoi(e) : 2; return (uint32_t)(v >= 2 && v <= 8 ? v : 2); }();
        const uint32_t want = ctx->tune_u ? (uint32_t)ctx->tune_u : pair_ns;
        if (plan.NS < want && plan.NS < chunks) {
            plan.NS = std::min(want, chunks);
            const uint32_t warps = kThreads / 32;
            plan.smem_bytes = warps * plan.xs_stride * 4 + warps * plan.NS * plan.G * wa->rec_bytes + warps * plan.G * kPlaneRow + warps * plan.NS * 8;
        }
    }
    QGemvPairBatch bt;
    memset(&bt, 0, sizeof(bt));
    const ZgCudaQWeight* ws_w[2] = {wa, wb};
    float* outs[2] = {d_out_a, d_out_b};
    const ZgGemvWs* wss[2] = {ws_a, ws_b};
    for (int i = 0; i < 2; i++) {
        QGemvParams& q = i ? bt.p[0].b : bt.p[0].a;
        const ZgCudaQWeight* w = ws_w[i];
        size_t pe = 0, nc = 0;
        zg_qgemv_ws_need(ctx, w, 1, &pe, &nc);
        if (plan.S > 1 && (!wss[i] || wss[i]->partials_elems < pe || wss[i]->counters_n < nc)) { zg_set_error("internal: pair split workspace too small"); return false; }
        q.recs = w->recs; q.smax = w->smax; q.n_kc = w->n_kc; q.n_nb = w->n_nb;
        q.K = (uint32_t)w->K; q.N = (uint32_t)w->N; q.M = 1;
        q.x = d_in; q.x_rs = (uint32_t)w->K; q.xs_stride = plan.xs_stride;
        q.out = outs[i]; q.out_rs = (uint32_t)w->N;
        q.P = plan.P; q.S = plan.S; q.NS = plan.NS;
        q.cl = (ctx->gemv_cluster && (plan.S == 2 || plan.S == 4 || plan.S == 8)) ? 1u : 0u;
        q.partials = wss[i] ? wss[i]->partials : nullptr; q.counters = wss[i] ? wss[i]->counters : nullptr;
    }
    bt.p[0].epi = epi;
    switch (wa->fmt) {
        case ZG_QFMT_I8_F32: return launch_pair_fmt<ZG_QFMT_I8_F32>(plan, bt, st, ctx->pdl);
        case ZG_QFMT_I8_F16: return launch_pair_fmt<ZG_QFMT_I8_F16>(plan, bt, st, ctx->pdl);
        case ZG_QFMT_I4_F16: return launch_pair_fmt<ZG_QFMT_I4_F16>(plan, bt, st, ctx->pdl);
        default: return false;
    }
}
