// Prefill-path quantized matmul for sm_100a: dst[M, N] = x[M, K] * dequant(W)[K, N], M > 8.
//
// zgml's W8·f32 algorithm (QuantizedWeight.matmul, src/quant.zig:475-578; DeviceOp.qmatmul,
// src/backend/reference.zig:499-566) as a dense contraction on the 5th-generation tensor cores:
//
//   * tcgen05.mma.cta_group::1.kind::f16 on BF16 operands, 128 x 256 output tile per CTA, fp32 accumulators in TMEM
//     (128 lanes x 256 columns), issued by one thread; K advances 64 per pipeline stage (4 MMAs of K = 16 per term).
//   * A = activations: split once into two BF16 terms x = x_hi + x_lo (each round-to-nearest) in a dense scratch,
//     then TMA (cp.async.bulk.tensor.2d, 128B swizzle) straight into the canonical K-major shared-memory layout.
//   * B = weights: sixteen dequantize warps read the packed records (zg_internal.cuh) with 128-bit loads, form
//     w = f32(q) * s exactly like dequantizeTo (src/quant.zig:594-618), split w = w_hi + w_lo in BF16 — all in
//     registers, before waiting for the stage — then store 8-byte runs (four consecutive k) into the same swizzled
//     K-major layout with a lane / record assignment that makes the 64-bit stores bank-conflict free.
//     Weights are dequantized once per 128 activation rows, in shared memory only.
//   * 2-stage mbarrier pipeline (96 KB per stage: hi and lo tiles of A and B): TMA warp / dequant warps -> MMA warp
//     -> (tcgen05.commit) -> stage free; the dequant warps turn into the epilogue: tcgen05.ld 32x32b, a 32 x 32
//     transpose through the idle stage memory, 128-byte row-segment stores.
//
// Numerics: "3xBF16" — both operands are split hi + lo (16 significant bits, residual 2^-18), D += hi*hi + hi*lo +
// lo*hi with fp32 accumulation in TMEM: ~5e-6 relative on outputs (the dropped lo*lo and the residuals are 2^-17 per
// product and average out over K), far inside the 1e-3 budget and twice the tensor throughput of the TF32 form
// (BF16 MMAs run at the full dense rate).  A single BF16 term (4e-3) would not survive 200+ chained linears.
// The exact fixed-point matvec (qgemv.cu) stays the path for M <= 8.
#include "zg_internal.cuh"

#include <cuda.h>
#include <stdlib.h>

namespace {

constexpr uint32_t BK = 64;                                        // BK bf16 = one 128-byte swizzle row = two records of 32 k
// MH = accumulator halves along M.  MH = 1 (default): 128 x 256 tile.  MH = 2 (ZG_GEMM_MH=2, experiment): 256 x 128 tile,
// two 128-row accumulators in TMEM fed from ONE dequantized weight tile — the dequantize instructions per flop halve
// (ncu: 84 M -> 47 M warp instructions at 2048 x 4096 x 4096) but the time does not change (0-4 % slower): the dequantize
// warps are not the limiter.  Kept because it is parity-tested and the right B-side shape for a future cta_group::2 tile.
__host__ __device__ constexpr uint32_t tbm(int MH) { return 128u * MH; }
__host__ __device__ constexpr uint32_t tbn(int MH) { return MH == 1 ? 256u : 128u; }
__host__ __device__ constexpr uint32_t tile_a(int MH) { return tbm(MH) * BK * 2; }   // one BF16 A tile: 16 / 32 KB
__host__ __device__ constexpr uint32_t tile_b(int MH) { return tbn(MH) * BK * 2; }   // one BF16 B tile: 32 / 16 KB
constexpr uint32_t kDqWarps = 16;                                  // dequantize warps: enough resident warps to hide ALU latency (8 left the MMA waiting)
constexpr uint32_t kGemmThreads = 64 + 32 * kDqWarps;              // warp 0: TMA, warp 1: MMA + TMEM, warps 2-17: dequant + epilogue
constexpr uint32_t kTmemCols = 256;
// NT = BF16 terms per operand.  NT = 2 (default): x = x_hi + x_lo, w = w_hi + w_lo, D += hi*hi + hi*lo + lo*hi
// ("3xBF16": ~5e-6 relative); NT = 1: one rounded term each (~4e-3 relative, 3x fewer MMAs; throughput experiments only).
// Ring slots: 2 of 96 KB (NT = 2), 4 of 48 KB (NT = 1).  Measured and rejected: separate rings with a third slot for the
// TMA-fed activation tile (3 x 32 KB + 2 x 64 KB, own empty barriers): 3-5 % slower — the kernel is not waiting on TMA.
__host__ __device__ constexpr uint32_t stages_of(int NT) { return NT == 1 ? 4u : 2u; }
__host__ __device__ constexpr uint32_t smem_of(int NT) { return stages_of(NT) * NT * (tile_a(1) + tile_b(1)) + 1024 /* alignment slack */ + 256 /* barriers */; }   // same for MH = 2

struct QGemmParams {
    const uint8_t* recs;
    uint32_t n_kc, n_nb;
    uint32_t M, N;
    float* out;
    uint32_t out_rs;
    uint32_t out_vec4;   // destination rows are 16-byte aligned: the epilogue may use 128-bit stores
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t c0, uint32_t c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// {lo16, hi16} = {bf16_rn(a), bf16_rn(b)}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ float bf16_lo_f32(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16_hi_f32(uint32_t packed) { return __uint_as_float(packed & 0xFFFF0000u); }
// K-major, 128-byte swizzle, 8-row groups 1024 B apart (SBO); LBO unused for swizzled K-major layouts.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address
    d |= (uint64_t)1 << 16;                               // leading byte offset (ignored)
    d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset
    d |= (uint64_t)1 << 46;                               // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
    return d;
}

// hi = bf16_rn(x), lo = bf16_rn(x - hi) into dense [Mp][Kp] bf16 planes (plane 1 only when NT == 2); zero padding.
// Two k per thread (one packed 32-bit store per plane).
__global__ void k_split_bf16(const float* __restrict__ x, uint32_t x_rs, uint32_t* __restrict__ xr, uint32_t kp2,
                             uint32_t M, uint32_t K, size_t plane_words, int NT) {
    const uint32_t m = blockIdx.y;
    for (uint32_t k2 = blockIdx.x * blockDim.x + threadIdx.x; k2 < kp2; k2 += gridDim.x * blockDim.x) {
        const uint32_t k = 2 * k2;
        const float a = (m < M && k < K) ? x[(size_t)m * x_rs + k] : 0.0f;
        const float b = (m < M && k + 1 < K) ? x[(size_t)m * x_rs + k + 1] : 0.0f;
        const uint32_t hi = pack_bf16x2(a, b);
        xr[(size_t)m * kp2 + k2] = hi;
        if (NT == 2) xr[plane_words + (size_t)m * kp2 + k2] = pack_bf16x2(a - bf16_lo_f32(hi), b - bf16_hi_f32(hi));
    }
}

template <int FMT, int NT, int MH>
__global__ void __launch_bounds__(kGemmThreads, 1)
qgemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a_lo, const QGemmParams p) {
    constexpr uint32_t kStages = stages_of(NT);
    constexpr uint32_t BM = tbm(MH), BN = tbn(MH), kTileA = tile_a(MH), kTileB = tile_b(MH);
    constexpr uint32_t kStageA = NT * kTileA, kStageB = NT * kTileB;   // [hi | lo] tiles
    constexpr uint32_t kDqPerStage = MH == 1 ? kDqWarps : kDqWarps / 2;   // MH = 2: the two halves of the dequant warps alternate stages
    static_assert(NT != 1 || MH == 1, "the single-term experiment keeps the 128 x 256 tile");
    constexpr bool kI4 = (FMT == ZG_QFMT_I4_F16);
    constexpr bool kF32 = (FMT == ZG_QFMT_I8_F32);
    constexpr uint32_t QB = kI4 ? 512u : 1024u;
    constexpr uint32_t SB = kF32 ? 32u : 16u;
    constexpr uint32_t RB = QB + 4 * SB;

    extern __shared__ uint8_t dsm_raw[];
    const uint32_t base = (smem_u32(dsm_raw) + 1023u) & ~1023u;        // 128B-swizzle atoms need 1024-byte alignment
    const uint32_t sA = base, sB = base + kStages * kStageA;
    const uint32_t bars = sB + kStages * kStageB;
    const uint32_t full_a = bars, full_b = bars + 8 * kStages, empty = bars + 16 * kStages, tmem_full = bars + 24 * kStages;
    const uint32_t tmem_slot = tmem_full + 8;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tile_n = blockIdx.x, tile_m = blockIdx.y;
    const uint32_t n_rec = p.n_kc;                 // records (32 k each) per column group
    const uint32_t n_k = (n_rec + 1) / 2;          // pipeline stages of 64 k

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < kStages; s++) {
            mbar_init(full_a + 8 * s, 1);
            mbar_init(full_b + 8 * s, kDqPerStage);   // one arrive per dequant warp working on the stage
            mbar_init(empty + 8 * s, 1);       // tcgen05.commit
        }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM accumulator: 128 columns x 128 lanes of fp32
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ── TMA producer: activations ──
        if (lane == 0) {
            for (uint32_t kt = 0; kt < n_k; kt++) {
                const uint32_t s = kt % kStages, ph = (kt / kStages) & 1;
                mbar_wait(empty + 8 * s, ph ^ 1);
                mbar_expect_tx(full_a + 8 * s, kStageA);
                tma_load_2d(sA + s * kStageA, &tmap_a, kt * BK, tile_m * BM, full_a + 8 * s);
                if constexpr (NT == 2) tma_load_2d(sA + s * kStageA + kTileA, &tmap_a_lo, kt * BK, tile_m * BM, full_a + 8 * s);
            }
        }
    } else if (warp == 1) {
        // ── MMA issuer ──
        // instruction descriptor (kind::f16): D = F32, A = B = BF16, both K-major, N = BN, M = 128
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((BN >> 3) << 17) | ((128u >> 4) << 24);
        for (uint32_t kt = 0; kt < n_k; kt++) {
            const uint32_t s = kt % kStages, ph = (kt / kStages) & 1;
            mbar_wait(full_a + 8 * s, ph);
            mbar_wait(full_b + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
#pragma unroll
                for (uint32_t k = 0; k < BK / 16; k++) {   // one MMA = 16 bf16 of k = 32 bytes along the swizzled row
                    // term 0: hi*hi; NT == 2 adds hi*lo and lo*hi (lo*lo is below fp32 resolution)
#pragma unroll
                    for (int term = 0; term < (NT == 2 ? 3 : 1); term++) {
                        const uint32_t a_off = (term == 2) ? kTileA : 0u, b_off = (term == 1) ? kTileB : 0u;
                        const uint64_t db = make_desc(sB + s * kStageB + b_off + k * 32);
                        const uint32_t accumulate = (kt | k | (uint32_t)term) ? 1u : 0u;
#pragma unroll
                        for (int mh = 0; mh < MH; mh++) {   // 128-row halves of the A tile -> their own TMEM accumulator (BN columns apart)
                            const uint64_t da = make_desc(sA + s * kStageA + a_off + (uint32_t)mh * (128u * 128u) + k * 32);
                            asm volatile(
                                "{\n\t.reg .pred p;\n\t"
                                "setp.ne.b32 p, %4, 0;\n\t"
                                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                ::"r"(tmem_base + (uint32_t)mh * BN), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
                        }
                    }
                }
                // frees the stage when the MMAs that read it have completed
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty + 8 * s) : "memory");
                if (kt + 1 == n_k)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tmem_full) : "memory");
            }
            __syncwarp();
        }
    } else {
        // ── dequantize warps: warp (dw, ct) owns 16 columns (column tile ct of column group dw) of BOTH records of a
        //    stage (one record = 32 columns x 32 k; a stage = 64 k = records 2 kt and 2 kt + 1).
        //    Lanes with odd g take the two records in swapped order (slot s holds record 2 kt + (s ^ (g & 1))): one
        //    64-bit store instruction then spreads a half-warp's four rows g over four different 8-bank groups of the
        //    128B-swizzled tile.  With one record per warp (the first version) rows 2j and 2j + 1 always shared a bank
        //    group — ncu: 57 % excess store wavefronts, and shared-memory bandwidth (operand reads of the MMAs + these
        //    stores + the TMA writes) is what bounds this kernel. ──
        const uint32_t dwi = warp - 2;                   // 0..15
        // MH = 1: 8 column groups x 2 column tiles, every stage.  MH = 2: 4 column groups x 2 column tiles x 2 stage parities.
        const uint32_t dw = MH == 1 ? (dwi & 7) : (dwi & 3), ct = MH == 1 ? (dwi >> 3) : ((dwi >> 2) & 1);
        const uint32_t par = MH == 1 ? 0u : (dwi >> 3), kstep = MH == 1 ? 1u : 2u;   // this warp's stages: par, par + kstep, ...
        const uint32_t g = lane >> 2, t = lane & 3, gx = g & 1;
        const uint32_t nb = tile_n * (BN / 32) + dw;
        const bool nb_ok = nb < p.n_nb;
        const uint8_t* rec = p.recs + (size_t)(nb_ok ? nb : 0) * p.n_kc * RB;
        // register ring: the records of the next kPf stages are in flight while the current one is converted
        // (f32 scales take twice the registers: one stage ahead — a full MMA stage time — still covers an L2 hit)
        constexpr int kPf = kF32 ? 1 : 2;
        uint4 rq[kPf][2], rs0[kPf][2], rs1[kPf][2];
        auto load_rec = [&](int slot, uint32_t j) {   // this warp's j-th stage; records past the end: zeros
            const uint32_t kt = par + j * kstep;
#pragma unroll
            for (int sl = 0; sl < 2; sl++) {
                const uint32_t ri = 2 * kt + ((uint32_t)sl ^ gx);
                rq[slot][sl] = make_uint4(0, 0, 0, 0); rs0[slot][sl] = rq[slot][sl]; rs1[slot][sl] = rq[slot][sl];
                if (nb_ok && kt < n_k && ri < n_rec) {
                    const uint8_t* r = rec + (size_t)ri * RB;
                    if constexpr (kI4) {
                        const uint2 v = __ldg(reinterpret_cast<const uint2*>(r + lane * 16 + ct * 8));
                        rq[slot][sl].x = v.x; rq[slot][sl].y = v.y;
                    } else {
                        rq[slot][sl] = __ldg(reinterpret_cast<const uint4*>(r + ct * 512 + lane * 16));
                    }
                    rs0[slot][sl] = __ldg(reinterpret_cast<const uint4*>(r + QB + t * SB));
                    if constexpr (kF32) rs1[slot][sl] = __ldg(reinterpret_cast<const uint4*>(r + QB + t * SB + 16));
                }
            }
        };
#pragma unroll
        for (int i = 0; i < kPf; i++) load_rec(i, i);
        for (uint32_t j0 = 0; par + j0 * kstep < n_k; j0 += kPf) {
#pragma unroll
          for (int slot = 0; slot < kPf; slot++) {
            const uint32_t kt = par + (j0 + slot) * kstep;
            if (kt >= n_k) break;
            const uint32_t s = kt % kStages, ph = (kt / kStages) & 1;
            // Convert into registers FIRST: nothing here needs the stage to be free, so the only work left between the
            // MMA releasing the stage and this warp's arrive is the stores.
            uint32_t pk[2][4][2 * NT];   // [record slot][unit] -> {hi k0k1, hi k2k3, lo k0k1, lo k2k3}
#pragma unroll
            for (int sl = 0; sl < 2; sl++) {
            const uint4 q = rq[slot][sl], s0 = rs0[slot][sl], s1 = rs1[slot][sl];
            float sc[8];   // scales of rows k = 4t + i (i < 4) and 16 + 4t + (i - 4)
            if constexpr (kF32) {
                sc[0] = __uint_as_float(s0.x); sc[1] = __uint_as_float(s0.y); sc[2] = __uint_as_float(s0.z); sc[3] = __uint_as_float(s0.w);
                sc[4] = __uint_as_float(s1.x); sc[5] = __uint_as_float(s1.y); sc[6] = __uint_as_float(s1.z); sc[7] = __uint_as_float(s1.w);
            } else {
                const uint32_t hw[4] = {s0.x, s0.y, s0.z, s0.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[i]));
                    sc[2 * i] = f.x; sc[2 * i + 1] = f.y;
                }
            }
#pragma unroll
            for (int r = 0; r < 4; r++) {
                // unit r: column n = 16 ct + g + 8 (r & 1), rows k = 4 t + b + 16 (r >> 1), b = 0..3.
                // int -> float through the mantissa: bits 0x4B000000 | u are the float 2^23 + u exactly.
                float qf[4];
                if constexpr (!kI4) {
                    const uint32_t w = (r == 0 ? q.x : (r == 1 ? q.y : (r == 2 ? q.z : q.w))) ^ 0x80808080u;   // u = q + 128
#pragma unroll
                    for (int b = 0; b < 4; b++) qf[b] = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650 + b)) - 8388736.0f;
                } else {
                    // {w0, w1}: byte b low nibble = column g, high nibble = column g + 8 (both biased by 8)
                    const uint32_t w = (r >> 1) ? q.y : q.x;
                    const uint32_t nib = (r & 1) ? ((w >> 4) & 0x0F0F0F0Fu) : (w & 0x0F0F0F0Fu);
#pragma unroll
                    for (int b = 0; b < 4; b++) qf[b] = __uint_as_float(__byte_perm(nib, 0x4B000000u, 0x7650 + b)) - 8388616.0f;
                }
                float wv[4];
#pragma unroll
                for (int b = 0; b < 4; b++) wv[b] = qf[b] * sc[4 * (r >> 1) + b];   // f32(q) * scale, src/quant.zig:612-615
                const uint32_t h0 = pack_bf16x2(wv[0], wv[1]), h1 = pack_bf16x2(wv[2], wv[3]);
                pk[sl][r][0] = h0; pk[sl][r][1] = h1;
                if constexpr (NT == 2) {
                    pk[sl][r][2] = pack_bf16x2(wv[0] - bf16_lo_f32(h0), wv[1] - bf16_hi_f32(h0));
                    pk[sl][r][3] = pack_bf16x2(wv[2] - bf16_lo_f32(h1), wv[3] - bf16_hi_f32(h1));
                }
            }
            }
            load_rec(slot, j0 + slot + kPf);
            // Keep the packed values live ACROSS the wait: ptxas otherwise sinks the packs below the wait loop (seen in the
            // SASS).  A fold of all of them feeds a store that never executes (M is never 2^32 - 1), which pins them here.
            uint32_t fold = 0;
#pragma unroll
            for (int sl = 0; sl < 2; sl++)
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int i = 0; i < 2 * NT; i++) fold ^= pk[sl][r][i];
            if (p.M == 0xFFFFFFFFu) asm volatile("st.shared.u32 [%0], %1;" ::"r"(tmem_slot), "r"(fold) : "memory");
            mbar_wait(empty + 8 * s, ph ^ 1);
            const uint32_t stage = sB + s * kStageB;
#pragma unroll
            for (int sl = 0; sl < 2; sl++) {
            const uint32_t hh = (uint32_t)sl ^ gx;               // which record (k half of the stage) this slot holds
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const uint32_t n = dw * 32 + ct * 16 + g + 8 * (r & 1);   // B-tile row (output column)
                // four consecutive k starting at 32 hh + 16 (r >> 1) + 4 t  ->  8 bytes at byte 2 k of the 128-byte row
                const uint32_t chunk = 4 * hh + 2 * (r >> 1) + (t >> 1);         // 16-byte chunk = eight consecutive k
                const uint32_t addr = stage + n * 128 + ((chunk ^ (n & 7)) << 4) + ((t & 1) << 3);
                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(pk[sl][r][0]), "r"(pk[sl][r][1]) : "memory");
                if constexpr (NT == 2)
                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr + kTileB), "r"(pk[sl][r][2]), "r"(pk[sl][r][3]) : "memory");
            }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(full_b + 8 * s);
          }
        }
        // ── epilogue: TMEM -> registers -> global.  Warp w may touch TMEM lanes 32 (w % 4) .. + 31 ──
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t quarter = warp & 3;
        // the four warps of a TMEM lane quarter split the work: MH = 1: 4 x 64 columns; MH = 2: accumulator half x 64 columns
        const uint32_t e_mh = MH == 1 ? 0u : ((dwi >> 2) & 1u);
        const uint32_t m_base = tile_m * BM + e_mh * 128 + quarter * 32;   // first of this warp's 32 output rows
        uint4* tb = reinterpret_cast<uint4*>(dsm_raw + (base - smem_u32(dsm_raw)) + dwi * (32 * 36 * 4));
        const uint32_t c_begin = MH == 1 ? (dwi >> 2) * 64u : (dwi >> 3) * 64u;
#pragma unroll 1
        for (uint32_t c0 = c_begin; c0 < c_begin + 64; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((quarter * 32) << 16) + e_mh * BN + c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            // lane = output row here; a direct store would touch 32 rows (32 sectors) per instruction — ncu: 19 % of all
            // warp samples stalled on those.  Transpose the 32 x 32 block through this warp's patch of the (now idle)
            // stage memory (row pitch 36 words: conflict-free both ways) and write whole 128-byte row segments.
#pragma unroll
            for (int i = 0; i < 8; i++)
                tb[lane * 9 + i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            __syncwarp();
            const uint32_t n0 = tile_n * BN + c0;
            if (n0 < p.N) {                                  // N % 32 == 0: whole column groups only
                if (p.out_vec4) {
#pragma unroll
                    for (int it = 0; it < 8; it++) {         // four rows x eight 16-byte pieces per instruction
                        const uint32_t rr = it * 4 + (lane >> 3), mm = m_base + rr;
                        const uint4 o = tb[rr * 9 + (lane & 7)];
                        if (mm < p.M) *reinterpret_cast<uint4*>(p.out + (size_t)mm * p.out_rs + n0 + 4 * (lane & 7)) = o;
                    }
                } else {
                    const uint32_t* tw = reinterpret_cast<const uint32_t*>(tb);
#pragma unroll 4
                    for (int rr = 0; rr < 32; rr++)          // unaligned destination: one row segment per instruction
                        if (m_base + rr < p.M) p.out[(size_t)(m_base + rr) * p.out_rs + n0 + lane] = __uint_as_float(tw[rr * 36 + lane]);
                }
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

bool get_encode() {
    if (g_encode) return true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
        zg_set_error("cuTensorMapEncodeTiled entry point not available");
        return false;
    }
    g_encode = (EncodeTiledFn)fn;
    return true;
}

template <int FMT, int NT, int MH>
bool launch_gemm(const CUtensorMap& map, const CUtensorMap& map_lo, const QGemmParams& p, cudaStream_t st) {
    dim3 grid((p.N + tbn(MH) - 1) / tbn(MH), (p.M + tbm(MH) - 1) / tbm(MH));
    qgemm_bf16_kernel<FMT, NT, MH><<<grid, kGemmThreads, smem_of(NT), st>>>(map, map_lo, p);
    ZG_COUNT_LAUNCH();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { zg_set_error("qgemm launch failed: %s", cudaGetErrorString(e)); return false; }
    return true;
}

template <int FMT, int NT, int MH>
bool set_attr() {
    cudaError_t e = cudaFuncSetAttribute(qgemm_bf16_kernel<FMT, NT, MH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of(NT));
    if (e != cudaSuccess) { zg_set_error("cudaFuncSetAttribute(qgemm) failed: %s", cudaGetErrorString(e)); return false; }
    return true;
}

template <int FMT>
bool launch_fmt(int NT, int MH, const CUtensorMap& map, const CUtensorMap& map_lo, const QGemmParams& p, cudaStream_t st) {
    if (NT == 1) return launch_gemm<FMT, 1, 1>(map, map_lo, p, st);
    return MH == 2 ? launch_gemm<FMT, 2, 2>(map, map_lo, p, st) : launch_gemm<FMT, 2, 1>(map, map_lo, p, st);
}

int g_terms = 2;   // ZG_GEMM_X1=1 selects the single-term mode (throughput experiments only: 4e-3 relative)
int g_mh = 1;      // ZG_GEMM_MH=2 selects the 256 x 128 two-accumulator tile (experiment, see above); default 128 x 256

int pick_mh(uint32_t) { return (g_terms == 2 && g_mh == 2) ? 2 : 1; }
int g_cta2 = 1;    // the CTA-pair kernel of qgemm_cta2.cu for M > 128 (validated round 2: parity green, +11 % over the 1-CTA tile); ZG_GEMM_CTA2=0 selects the 1-CTA tile

} // namespace

bool zg_qgemm_cta2_launch(const CUtensorMap& map, const CUtensorMap& map_lo, const ZgCudaQWeight* w, uint32_t M, float* d_out,
                          uint32_t out_rs, cudaStream_t st);   // qgemm_cta2.cu

bool zg_qgemm_init(ZgCudaCtx*) {
    if (const char* e = getenv("ZG_GEMM_X1")) g_terms = (e[0] == '1') ? 1 : 2;
    if (const char* e = getenv("ZG_GEMM_MH")) g_mh = (e[0] == '2') ? 2 : 1;
    if (const char* e = getenv("ZG_GEMM_CTA2")) g_cta2 = (e[0] == '1') ? 1 : 0;
    return set_attr<ZG_QFMT_I8_F32, 1, 1>() && set_attr<ZG_QFMT_I8_F16, 1, 1>() && set_attr<ZG_QFMT_I4_F16, 1, 1>() &&
           set_attr<ZG_QFMT_I8_F32, 2, 1>() && set_attr<ZG_QFMT_I8_F16, 2, 1>() && set_attr<ZG_QFMT_I4_F16, 2, 1>() &&
           set_attr<ZG_QFMT_I8_F32, 2, 2>() && set_attr<ZG_QFMT_I8_F16, 2, 2>() && set_attr<ZG_QFMT_I4_F16, 2, 2>() && get_encode();
}

// BF16 hi and lo planes of the activations the GEMM's TMA reads: 2 x [round_up(M, 256)][round_up(K, 64)] bf16,
// counted in f32 elements (the workspace unit).  256 rows: whole TMA boxes for either tile shape.
size_t zg_qgemm_scratch_elems(const ZgCudaQWeight* w, uint32_t M) {
    if (w->fmt == ZG_QFMT_GENERIC || M <= 8) return 0;
    return (size_t)((M + 255) / 256 * 256) * ((size_t)(w->n_kc + 1) / 2 * BK);
}

bool zg_qgemm_launch(ZgCudaCtx* ctx, const ZgCudaQWeight* w, const float* d_in, float* d_out, uint32_t M, uint32_t in_rs,
                     uint32_t out_rs, float* scratch, cudaStream_t st) {
    (void)ctx;
    if (!get_encode()) return false;
    const uint32_t Kp = (w->n_kc + 1) / 2 * BK, Mp = (M + 255) / 256 * 256;
    const size_t plane_words = (size_t)Mp * (Kp / 2);   // one bf16 plane in 32-bit words
    const bool cta2 = g_cta2 && g_terms == 2 && M > 128;
    const int NT = g_terms, MH = cta2 ? 1 : pick_mh(M);   // the pair kernel loads 128-row boxes like the 128 x 256 tile
    uint32_t* planes = reinterpret_cast<uint32_t*>(scratch);
    k_split_bf16<<<dim3((Kp / 2 + 255) / 256, Mp), 256, 0, st>>>(d_in, in_rs, planes, Kp / 2, M, (uint32_t)w->K, plane_words, NT);
    ZG_COUNT_LAUNCH();
    CUtensorMap map[2];
    const cuuint64_t gdim[2] = {Kp, Mp};
    const cuuint64_t gstride[1] = {(cuuint64_t)Kp * 2};
    const cuuint32_t box[2] = {BK, tbm(MH)};
    const cuuint32_t estr[2] = {1, 1};
    for (int i = 0; i < 2; i++) {
        CUresult r = g_encode(&map[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, planes + (size_t)i * plane_words, gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { zg_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return false; }
    }
    if (cta2) return zg_qgemm_cta2_launch(map[0], map[1], w, M, d_out, out_rs, st);
    QGemmParams p;
    p.recs = w->recs; p.n_kc = w->n_kc; p.n_nb = w->n_nb; p.M = M; p.N = (uint32_t)w->N; p.out = d_out; p.out_rs = out_rs;
    p.out_vec4 = ((reinterpret_cast<uintptr_t>(d_out) & 15) == 0 && (out_rs & 3) == 0) ? 1u : 0u;
    switch (w->fmt) {
        case ZG_QFMT_I8_F32: return launch_fmt<ZG_QFMT_I8_F32>(NT, MH, map[0], map[1], p, st);
        case ZG_QFMT_I8_F16: return launch_fmt<ZG_QFMT_I8_F16>(NT, MH, map[0], map[1], p, st);
        case ZG_QFMT_I4_F16: return launch_fmt<ZG_QFMT_I4_F16>(NT, MH, map[0], map[1], p, st);
        default: break;
    }
    zg_set_error("qgemm: unknown weight format %d", w->fmt);
    return false;
}
