// The prefill quantized matmul of qgemm.cu as a CTA-PAIR kernel: tcgen05.mma.cta_group::2, one 256 x 256 output tile per
// cluster of two CTAs.  Default for M > 128 since round 2 (parity tests green with it, 427 / 458 / 457 TFLOP/s against
// 384 / 404 / 411 for the 1-CTA tile at M = 2048 on the Llama-3-8B linears); ZG_GEMM_CTA2=0 falls back to the 1-CTA tile.
//
// Why (profiles/r01_qgemm_ncu_full_stall_summary.txt): the 1-CTA 128 x 256 tile is bound by shared-memory
// bandwidth — every 128 x 256 x 16 MMA reads 12 KB of operands in its 128 cycles, and the dequantized weight tile has to be
// stored through the same pipe: 1152 + 512 + 256 = 1920 wavefronts per 64-k stage against 1536 cycles of MMA.  In a CTA pair
// each CTA holds its own 128 activation rows and only HALF of the weight tile (128 columns); the pair's MMA (M = 256,
// N = 256) reads 8 KB per CTA per k16 and each CTA dequantizes / stores half as much: 768 + 256 + 256 = 1280 < 1536.
// Stages shrink to 64 KB, so the ring is 3 deep.
//
// Protocol (same roles as qgemm.cu; rank = %cluster_ctarank, leader = rank 0):
//   * warp 0 of BOTH CTAs: TMA of its own 128 activation rows (hi and lo planes) with the .cta_group::2 form, completing
//     on the LEADER's full_a barrier (the leader's expect_tx covers both CTAs' bytes).
//   * warps 2-17 of both CTAs dequantize their CTA's 128 weight columns (4 column groups x 2 column tiles per stage; the two
//     halves of the warps alternate stages), fence.proxy.async, then arrive REMOTELY (mapa + mbarrier.arrive.release.cluster)
//     on the leader's full_b barrier: 8 warps x 2 CTAs = 16 arrivals.
//   * warp 1 of the leader issues the MMAs for the pair and commits with .multicast::cluster to both CTAs' empty barriers
//     (and, after the last stage, to both tmem_full barriers).  Each CTA's TMEM holds its own 128 rows x 256 columns.
//   * epilogue per CTA as in qgemm.cu (transpose through idle stage memory, 128-byte row stores).
//   * cluster barrier after the TMEM allocation and before the deallocation.
// Numerics identical to qgemm.cu (3xBF16, same conversions).
#include "zg_internal.cuh"

#include <cuda.h>

namespace {

constexpr uint32_t BK = 64, kStages = 3;
constexpr uint32_t kTileA = 128 * BK * 2, kTileB = 128 * BK * 2;      // one BF16 operand tile per CTA: 16 KB each
constexpr uint32_t kStageA = 2 * kTileA, kStageB = 2 * kTileB;        // [hi | lo]
constexpr uint32_t kDqWarps = 16, kThreads = 64 + 32 * kDqWarps, kTmemCols = 256;
constexpr uint32_t kSmem = kStages * (kStageA + kStageB) + 1024 + 256;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;                           // clears the pair-rank bit of a shared::cluster address

struct Params {
    const uint8_t* recs;
    uint32_t n_kc, n_nb, M, N;
    float* out;
    uint32_t out_rs, out_vec4;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// arrive on the barrier at the same offset in CTA `cta` of the cluster, releasing this thread's prior writes cluster-wide
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(cta) : "memory");
}
// Default (.acquire.cta) wait, as CUTLASS uses for barriers its peer CTA arrives on: the data guarded by full_a / full_b is
// consumed by the tensor cores (async proxy), not by this thread's loads — the producers' fence.proxy.async + release
// arrive order it.  A cluster-scope acquire here makes ptxas emit an L1 invalidate (CCTL.IVALL) after every wait, which
// would throw away the dequantize warps' record prefetches each stage.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// this CTA's box into its own shared memory, bytes completing on the leader CTA's barrier
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t c0, uint32_t c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar & kPeerMask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ float bf16_lo_f32(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16_hi_f32(uint32_t packed) { return __uint_as_float(packed & 0xFFFF0000u); }
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {   // K-major, 128-byte swizzle, 8-row groups 1024 B apart
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <int FMT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
qgemm_bf16_cta2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a_lo, const Params p) {
    constexpr bool kI4 = (FMT == ZG_QFMT_I4_F16);
    constexpr bool kF32 = (FMT == ZG_QFMT_I8_F32);
    constexpr uint32_t QB = kI4 ? 512u : 1024u;
    constexpr uint32_t SB = kF32 ? 32u : 16u;
    constexpr uint32_t RB = QB + 4 * SB;

    extern __shared__ uint8_t dsm_raw[];
    const uint32_t base = (smem_u32(dsm_raw) + 1023u) & ~1023u;
    const uint32_t sA = base, sB = base + kStages * kStageA;
    const uint32_t bars = sB + kStages * kStageB;
    const uint32_t full_a = bars, full_b = bars + 8 * kStages, empty = bars + 16 * kStages, tmem_full = bars + 24 * kStages;
    const uint32_t tmem_slot = tmem_full + 8;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const uint32_t tile_n = blockIdx.x >> 1, tile_m = blockIdx.y;   // the pair's 256 x 256 tile
    const uint32_t n_rec = p.n_kc, n_k = (n_rec + 1) / 2;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < kStages; s++) {
            mbar_init(full_a + 8 * s, 1);              // the leader's arrive.expect_tx (peer CTAs only add bytes)
            mbar_init(full_b + 8 * s, kDqWarps);       // 8 dequant warps per stage in each of the two CTAs
            mbar_init(empty + 8 * s, 1);               // multicast tcgen05.commit
        }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // one warp of EACH CTA allocates the pair's accumulator columns
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();   // both CTAs' barriers initialised and TMEM allocated before anything crosses the pair
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ── TMA producer (both CTAs): this CTA's 128 activation rows ──
        if (lane == 0) {
            for (uint32_t kt = 0; kt < n_k; kt++) {
                const uint32_t s = kt % kStages, ph = (kt / kStages) & 1;
                mbar_wait(empty + 8 * s, ph ^ 1);
                if (rank == 0) mbar_expect_tx(full_a + 8 * s, 2 * kStageA);   // both CTAs' hi + lo tiles
                tma_load_2d_pair(sA + s * kStageA, &tmap_a, kt * BK, tile_m * 256 + rank * 128, full_a + 8 * s);
                tma_load_2d_pair(sA + s * kStageA + kTileA, &tmap_a_lo, kt * BK, tile_m * 256 + rank * 128, full_a + 8 * s);
            }
        }
    } else if (warp == 1) {
        // ── MMA issuer: the leader CTA issues for the pair ──
        if (rank == 0) {
            // instruction descriptor (kind::f16): D = F32, A = B = BF16, both K-major, N = 256, M = 256 (two CTAs x 128 rows)
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
            for (uint32_t kt = 0; kt < n_k; kt++) {
                const uint32_t s = kt % kStages, ph = (kt / kStages) & 1;
                mbar_wait(full_a + 8 * s, ph);
                mbar_wait(full_b + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
#pragma unroll
                    for (uint32_t k = 0; k < BK / 16; k++) {
#pragma unroll
                        for (int term = 0; term < 3; term++) {   // hi*hi, hi*lo, lo*hi
                            const uint32_t a_off = (term == 2) ? kTileA : 0u, b_off = (term == 1) ? kTileB : 0u;
                            const uint64_t da = make_desc(sA + s * kStageA + a_off + k * 32);
                            const uint64_t db = make_desc(sB + s * kStageB + b_off + k * 32);
                            const uint32_t accumulate = (kt | k | (uint32_t)term) ? 1u : 0u;
                            asm volatile(
                                "{\n\t.reg .pred p;\n\t"
                                "setp.ne.b32 p, %4, 0;\n\t"
                                "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                ::"r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
                        }
                    }
                    // frees the stage in BOTH CTAs when the MMAs that read it have completed
                    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                 ::"r"(empty + 8 * s), "h"((uint16_t)3) : "memory");
                    if (kt + 1 == n_k)
                        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                     ::"r"(tmem_full), "h"((uint16_t)3) : "memory");
                }
                __syncwarp();
            }
        }
    } else {
        // ── dequantize warps: this CTA's 128 weight columns = 4 column groups; warp (dw, ct) owns 16 columns of both records
        //    of every second stage (mapping and conversions exactly as qgemm.cu's MH = 2 shape) ──
        const uint32_t dwi = warp - 2;
        const uint32_t dw = dwi & 3, ct = (dwi >> 2) & 1, par = dwi >> 3;
        const uint32_t g = lane >> 2, t = lane & 3, gx = g & 1;
        const uint32_t nb = tile_n * 8 + rank * 4 + dw;
        const bool nb_ok = nb < p.n_nb;
        const uint8_t* rec = p.recs + (size_t)(nb_ok ? nb : 0) * p.n_kc * RB;
        constexpr int kPf = kF32 ? 1 : 2;
        uint4 rq[kPf][2], rs0[kPf][2], rs1[kPf][2];
        auto load_rec = [&](int slot, uint32_t j) {
            const uint32_t kt = par + 2 * j;
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
        for (uint32_t j0 = 0; par + 2 * j0 < n_k; j0 += kPf) {
#pragma unroll
          for (int slot = 0; slot < kPf; slot++) {
            const uint32_t kt = par + 2 * (j0 + slot);
            if (kt >= n_k) break;
            const uint32_t s = kt % kStages, ph = (kt / kStages) & 1;
            uint32_t pk[2][4][4];   // [record slot][unit] -> {hi k0k1, hi k2k3, lo k0k1, lo k2k3}
#pragma unroll
            for (int sl = 0; sl < 2; sl++) {
            const uint4 q = rq[slot][sl], s0 = rs0[slot][sl], s1 = rs1[slot][sl];
            float sc[8];
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
                float qf[4];
                if constexpr (!kI4) {
                    const uint32_t w = (r == 0 ? q.x : (r == 1 ? q.y : (r == 2 ? q.z : q.w))) ^ 0x80808080u;
#pragma unroll
                    for (int b = 0; b < 4; b++) qf[b] = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650 + b)) - 8388736.0f;
                } else {
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
                pk[sl][r][2] = pack_bf16x2(wv[0] - bf16_lo_f32(h0), wv[1] - bf16_hi_f32(h0));
                pk[sl][r][3] = pack_bf16x2(wv[2] - bf16_lo_f32(h1), wv[3] - bf16_hi_f32(h1));
            }
            }
            load_rec(slot, j0 + slot + kPf);
            uint32_t fold = 0;   // keeps the packed values live across the wait (see qgemm.cu)
#pragma unroll
            for (int sl = 0; sl < 2; sl++)
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int i = 0; i < 4; i++) fold ^= pk[sl][r][i];
            if (p.M == 0xFFFFFFFFu) asm volatile("st.shared.u32 [%0], %1;" ::"r"(tmem_slot), "r"(fold) : "memory");
            mbar_wait(empty + 8 * s, ph ^ 1);
            const uint32_t stage = sB + s * kStageB;
#pragma unroll
            for (int sl = 0; sl < 2; sl++) {
            const uint32_t hh = (uint32_t)sl ^ gx;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const uint32_t n = dw * 32 + ct * 16 + g + 8 * (r & 1);   // row of this CTA's B tile (its local column)
                const uint32_t chunk = 4 * hh + 2 * (r >> 1) + (t >> 1);
                const uint32_t addr = stage + n * 128 + ((chunk ^ (n & 7)) << 4) + ((t & 1) << 3);
                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(pk[sl][r][0]), "r"(pk[sl][r][1]) : "memory");
                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr + kTileB), "r"(pk[sl][r][2]), "r"(pk[sl][r][3]) : "memory");
            }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to this SM's tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(full_b + 8 * s, 0);          // on the leader's barrier
          }
        }
        // ── epilogue: this CTA's 128 rows x 256 columns ──
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t quarter = warp & 3;
        const uint32_t m_base = tile_m * 256 + rank * 128 + quarter * 32;
        uint4* tb = reinterpret_cast<uint4*>(dsm_raw + (base - smem_u32(dsm_raw)) + dwi * (32 * 36 * 4));
        const uint32_t c_begin = (dwi >> 2) * 64u;
#pragma unroll 1
        for (uint32_t c0 = c_begin; c0 < c_begin + 64; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((quarter * 32) << 16) + c0;
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
#pragma unroll
            for (int i = 0; i < 8; i++)
                tb[lane * 9 + i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            __syncwarp();
            const uint32_t n0 = tile_n * 256 + c0;
            if (n0 < p.N) {
                if (p.out_vec4) {
#pragma unroll
                    for (int it = 0; it < 8; it++) {
                        const uint32_t rr = it * 4 + (lane >> 3), mm = m_base + rr;
                        const uint4 o = tb[rr * 9 + (lane & 7)];
                        if (mm < p.M) *reinterpret_cast<uint4*>(p.out + (size_t)mm * p.out_rs + n0 + 4 * (lane & 7)) = o;
                    }
                } else {
                    const uint32_t* tw = reinterpret_cast<const uint32_t*>(tb);
#pragma unroll 4
                    for (int rr = 0; rr < 32; rr++)
                        if (m_base + rr < p.M) p.out[(size_t)(m_base + rr) * p.out_rs + n0 + lane] = __uint_as_float(tw[rr * 36 + lane]);
                }
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();   // neither CTA releases TMEM or exits while the other may still read its memory
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
}

template <int FMT>
bool launch(const CUtensorMap& map, const CUtensorMap& map_lo, const Params& p, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(qgemm_bf16_cta2_kernel<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
        if (e != cudaSuccess) { zg_set_error("cudaFuncSetAttribute(qgemm cta2) failed: %s", cudaGetErrorString(e)); return false; }
        attr_done = true;
    }
    dim3 grid(2 * ((p.N + 255) / 256), (p.M + 255) / 256);   // x: cluster of two CTAs per 256-column tile
    qgemm_bf16_cta2_kernel<FMT><<<grid, kThreads, kSmem, st>>>(map, map_lo, p);
    ZG_COUNT_LAUNCH();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { zg_set_error("qgemm cta2 launch failed: %s", cudaGetErrorString(e)); return false; }
    return true;
}

} // namespace

// Called by zg_qgemm_launch (qgemm.cu) for M > 128 (unless ZG_GEMM_CTA2=0); `map` / `map_lo` are the 64 x 128-row boxes of the hi / lo planes.
bool zg_qgemm_cta2_launch(const CUtensorMap& map, const CUtensorMap& map_lo, const ZgCudaQWeight* w, uint32_t M, float* d_out,
                          uint32_t out_rs, cudaStream_t st) {
    Params p;
    p.recs = w->recs; p.n_kc = w->n_kc; p.n_nb = w->n_nb; p.M = M; p.N = (uint32_t)w->N; p.out = d_out; p.out_rs = out_rs;
    p.out_vec4 = ((reinterpret_cast<uintptr_t>(d_out) & 15) == 0 && (out_rs & 3) == 0) ? 1u : 0u;
    switch (w->fmt) {
        case ZG_QFMT_I8_F32: return launch<ZG_QFMT_I8_F32>(map, map_lo, p, st);
        case ZG_QFMT_I8_F16: return launch<ZG_QFMT_I8_F16>(map, map_lo, p, st);
        case ZG_QFMT_I4_F16: return launch<ZG_QFMT_I4_F16>(map, map_lo, p, st);
        default: break;
    }
    zg_set_error("qgemm cta2: unknown weight format %d", w->fmt);
    return false;
}
