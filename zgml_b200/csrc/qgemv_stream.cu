// Streamed decode matvec (one activation row) for sm_100a: the large-launch form of qgemv.cu.
//
//   dst[n] = sum_k ( x[k] * s[(k*N + n) / 32] ) * q[k, n]        (QuantizedWeight.matmul, src/quant.zig:475-578, M == 1)
//
// Same arithmetic as qgemv_kernel (digit planes of c = s * x on the integer tensor cores, see qgemv.cu), different
// work decomposition.  Measured on the Llama-3-70B matvecs (ncu, round 2): the k-split kernel spends as many issue
// slots on the per-column-group flush / CTA reduction / split arrival as on streaming, and a launch of 360 equal CTAs
// for 444 resident slots leaves SMs with 3 CTAs running 1.4x longer than SMs with 2.  Here
//
//  * the launch is ONE linear space: (matvec of the batch, block of 8 column groups, k) measured in chunks of G
//    records; CTA c takes chunks [c per, (c + 1) per) — every CTA the same amount whatever the shape, one round of
//    CTAs, `per` snapped to a divisor / multiple of a column group's chunk count so that cuts fall on few, aligned places;
//  * inside a CTA each warp owns ONE column group of the block and walks the CTA's whole k-range for it (a contiguous
//    run of records -> one TMA bulk copy per chunk into the warp's ring): no cross-warp reduction, no CTA barrier in
//    the streaming loop, one flush per (column group, segment) instead of one per 16 records;
//  * the activations of a segment (<= kSegRecs records of k) are staged ONCE per CTA, scaled by 0.499 / max|x| of the
//    segment; the column group's power-of-two scale ceiling is folded into the block scale (exact);
//  * a column group whose k-range is cut (by a CTA boundary or the staging capacity) gets one partial per piece in
//    global scratch; the warp that arrives last (atomic counter, result consumed one segment later so nobody waits
//    for the round trip) adds the pieces in k order -> deterministic.
#include "zg_internal.cuh"

#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <type_traits>

ZG_TRACE_DECL
void zg_trace_set_gemv_stream(unsigned long long* d_buf) { cudaMemcpyToSymbol(c_zg_trace, &d_buf, sizeof(d_buf)); }

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr uint32_t kSegRecs = 128;            // records of k staged per segment (16 KB of activations)
constexpr uint32_t kSegPerThread = kSegRecs * ZG_KR / kThreads;
constexpr uint32_t kPlaneRow = 144;           // bytes of digit planes per record: 4 x 32 B + 16 B bank skew

struct QGemvSOp {
    const uint8_t* recs; const float* smax; const float* x; float* out; float* partials; uint32_t* counters;
};
struct QGemvSParams {
    QGemvSOp op[kZgGemvBatch];
    uint32_t n_kc, n_nb, K, N;
    uint32_t GB;       // blocks of 8 column groups per matvec
    uint32_t nq;       // chunks per column group = ceil(n_kc / G)
    uint32_t TQ;       // chunks of the whole launch = count * GB * nq
    uint32_t per;      // chunks per CTA: CTA c takes [c * per, (c + 1) * per)
    uint32_t Lq;       // chunks per segment at most
    uint32_t NS;       // ring slots per warp
    uint32_t slots;    // partial slots per column group
    uint32_t rev;      // debug: CTA c takes the range of CTA grid - 1 - c
    uint32_t early;    // 1: release the dependent launch at kernel entry (its CTAs then compete for this kernel's SM slots)
};

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
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
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

template <int FMT>
__global__ void __launch_bounds__(kThreads, 3)
qgemv_stream_kernel(const __grid_constant__ QGemvSParams P) {
    constexpr bool kI4 = (FMT == ZG_QFMT_I4_F16);
    constexpr bool kF32 = (FMT == ZG_QFMT_I8_F32);
    constexpr uint32_t QB = kI4 ? 512u : 1024u;
    constexpr uint32_t SB = kF32 ? 32u : 16u;
    constexpr uint32_t RB = QB + 4 * SB;            // record bytes
    constexpr uint32_t G = kI4 ? 4u : 2u;           // records per chunk (one TMA bulk copy, one ring slot)
    constexpr uint32_t slot_bytes = G * RB;

    __shared__ float s_red[kWarps];
    // dynamic: [kSegRecs * 32] scaled activations | [warp][slot] ring | [warp][G] digit planes | [warp][slot] mbarriers
    extern __shared__ __align__(128) uint8_t dsm[];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t g = lane >> 2, t = lane & 3, j = g & 3;
    const uint32_t NS = P.NS;
    const uint32_t n_cta = gridDim.x;
    const uint32_t cta = P.rev ? n_cta - 1 - blockIdx.x : blockIdx.x;   // rev: debug (is a slow CTA slow by index or by address?)
    const uint32_t q_lo = cta * P.per;
    const uint32_t q_hi = min(q_lo + P.per, P.TQ);

    const uint32_t dsm_u32 = smem_u32(dsm);
    float* xs = reinterpret_cast<float*>(dsm);
    constexpr uint32_t xs_bytes = kSegRecs * ZG_KR * 4;
    const uint32_t ring = dsm_u32 + xs_bytes + warp * NS * slot_bytes;
    const uint32_t planes = dsm_u32 + xs_bytes + kWarps * NS * slot_bytes + warp * G * kPlaneRow;
    const uint32_t bars = dsm_u32 + xs_bytes + kWarps * NS * slot_bytes + kWarps * G * kPlaneRow + warp * NS * 8;

#ifdef ZG_STREAM_TRACE_ALL
    // tracing (zg_cuda_trace): EVERY CTA records entry / first activation staged / exit, with its SM id in the top byte of word 1
    unsigned long long* zt_slot = nullptr;
    if (c_zg_trace && threadIdx.x == 0) {
        unsigned long long tt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
        const unsigned long long sl = atomicAdd(c_zg_trace, 1ull);
        if (sl < 16000) { zt_slot = c_zg_trace + 1 + 3 * sl; zt_slot[0] = (12ull << 56) | (tt & 0xFFFFFFFFFFFFFFull); }
    }
#else
    ZG_TRACE_BEGIN(10)
#endif
    // Programmatic dependent launch is released LATE (after this CTA's last chunk): the grid fills every SM slot exactly once, and
    // a dependent kernel's CTAs becoming resident early would push some of this grid's CTAs into a second round.
    if (P.early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // the segment that starts at chunk q of this CTA's range: (linear block of column groups, first chunk inside it) -> chunks
    auto seg_at = [&](uint32_t q, uint32_t& gbl, uint32_t& qi) -> uint32_t {
        gbl = q / P.nq; qi = q - gbl * P.nq;
        return min(min(q_hi - q, P.nq - qi), P.Lq);
    };

    // ── producer (lane 0 of every warp): walks the same segments ahead of the consumer; weights are immutable, so the
    //    first NS chunks are requested before griddepcontrol.wait ──
    uint32_t pq = q_lo, p_recs = 0, pf_slot = 0;
    const uint8_t* p_src = nullptr;
    auto p_advance = [&]() {
        p_recs = 0;
        while (pq < q_hi) {
            uint32_t gbl, qi;
            const uint32_t n = seg_at(pq, gbl, qi);
            pq += n;
            const uint32_t i = gbl / P.GB, nb = (gbl - i * P.GB) * kWarps + warp;
            if (nb >= P.n_nb) continue;
            const uint32_t k0 = qi * G, k1 = min(k0 + n * G, P.n_kc);
            p_recs = k1 - k0;
            p_src = P.op[i].recs + ((size_t)nb * P.n_kc + k0) * RB;
            break;
        }
    };
    auto issue_chunk = [&]() {   // lane 0, p_recs > 0
        const uint32_t cnt = min(G, p_recs);
        const uint32_t bar = bars + pf_slot * 8;
        mbar_expect_tx(bar, cnt * RB);
        bulk_g2s(ring + pf_slot * slot_bytes, p_src, cnt * RB, bar);
        p_src += cnt * RB; p_recs -= cnt;
        if (++pf_slot == NS) pf_slot = 0;
        if (!p_recs) p_advance();
    };
    if (lane == 0) {
        for (uint32_t s = 0; s < NS; s++) mbar_init(bars + s * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        p_advance();
        for (uint32_t i = 0; i < NS && p_recs; i++) issue_chunk();
    }
    // the constant ones plane (digit index 3) of every record: B column of ones -> sum_k q
    for (uint32_t i = lane; i < G * 8; i += 32)
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(planes + (i >> 3) * kPlaneRow + 96 + (i & 7) * 4), "r"(0x01010101u) : "memory");
    __syncwarp();

    const uint32_t brow = planes + j * 32 + 4 * t;    // B column g = digit j of the activation row
    const uint32_t q_off = lane * 16;
    const uint32_t sc_off = QB + (kF32 ? 4u : 2u) * (8 * ((lane & 15) >> 2) + 4 * (lane >> 4) + (lane & 3));  // zg_scale_row_slot(lane)

    // split arrival whose outcome is consumed one segment later (lane 0 holds the counter's old value)
    uint32_t pend_old = 0, pend_nseg = 0, pend_nb = 0, pend_i = 0;
    auto resolve_pending = [&]() {
        if (!pend_nseg) return;
        __syncwarp();
        const uint32_t last = __shfl_sync(0xffffffffu, (pend_old == pend_nseg - 1) ? 1u : 0u, 0);
        if (last) {
            const QGemvSOp& o = P.op[pend_i];
            float v = 0.0f;
            for (uint32_t s2 = 0; s2 < pend_nseg; s2++) v += __ldcg(o.partials + ((size_t)pend_nb * P.slots + s2) * ZG_TN + lane);
            if (pend_nb * ZG_TN + lane < P.N) o.out[pend_nb * ZG_TN + lane] = v;
            if (lane == 0) o.counters[pend_nb] = 0u;   // re-arm for the next launch
        }
        pend_nseg = 0;
    };

    uint32_t slot = 0, parity = 0, slot_u32 = ring, bar_u32 = bars;
    bool first = true;
    for (uint32_t q = q_lo; q < q_hi;) {
        uint32_t gbl, qi;
        const uint32_t nq_seg = seg_at(q, gbl, qi);
        const uint32_t op_i = gbl / P.GB, nb = (gbl - op_i * P.GB) * kWarps + warp;
        const uint32_t k0 = qi * G, k1 = min(k0 + nq_seg * G, P.n_kc), L = k1 - k0;
        const QGemvSOp& o = P.op[op_i];

        // ── stage x'[k] = x[k] * 0.499 / max|x| of the segment's k-range (|s / smax * x'| <= 0.499).  Non-finite activations
        //    poison the sums (NaN out, like the reference); k >= K reads as zero ──
        if (first) {
            asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef ZG_STREAM_TRACE_ALL
            if (zt_slot) {
                unsigned long long tt; uint32_t smid;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                zt_slot[1] = ((unsigned long long)smid << 56) | (tt & 0xFFFFFFFFFFFFFFull);
            }
#else
            ZG_TRACE_MARK(1)
#endif
            first = false;
        }
        else __syncthreads();    // every warp is done with the previous segment's activations
        float mx = 0.0f;
        {
            float v[kSegPerThread];
            const uint32_t kb = k0 * ZG_KR, n_el = L * ZG_KR;
#pragma unroll
            for (uint32_t e = 0; e < kSegPerThread; e++) {
                const uint32_t idx = tid + e * kThreads;
                v[e] = (idx < n_el && kb + idx < P.K) ? o.x[kb + idx] : 0.0f;
                const float aa = fabsf(v[e]);
                mx = (aa <= 3.0e38f) ? fmaxf(mx, aa) : INFINITY;
            }
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
            if (lane == 0) s_red[warp] = mx;
            __syncthreads();
            mx = s_red[0];
#pragma unroll
            for (int w2 = 1; w2 < kWarps; w2++) mx = fmaxf(mx, s_red[w2]);
            const float f = (mx <= 3.0e38f && mx >= 1.0e-30f) ? (0.499f / mx) : 0.0f;   // tiny segments flush to zero
#pragma unroll
            for (uint32_t e = 0; e < kSegPerThread; e++) {
                const uint32_t idx = tid + e * kThreads;
                if (idx < n_el) xs[idx] = v[e] * f;
            }
        }
        __syncthreads();
        q += nq_seg;
        if (nb >= P.n_nb) continue;   // ragged last block of column groups: this warp has no group (uniform per warp)

        const float sm = __ldg(o.smax + nb);          // power of two >= every scale of the column group
        const float rsm = 1.0f / sm;
        int acc[2][4];
        uint32_t dsum = 0;
#pragma unroll
        for (int ct = 0; ct < 2; ct++)
#pragma unroll
            for (int i = 0; i < 4; i++) acc[ct][i] = 0;

        uint32_t xa = dsm_u32 + lane * 4;   // this lane's (k = lane) staged activation of the chunk's first record
        for (uint32_t rec = 0; rec < L; rec += G, xa += G * ZG_KR * 4) {
            const uint32_t cnt = min(G, L - rec);
            mbar_wait(bar_u32, parity);
            auto chunk = [&](auto full_tag) {
                constexpr bool FULL = decltype(full_tag)::value;
                // digit generation, lane = k: F = (s / smax) * x' + 1.5 in (1, 2); the three low bytes of F are base-256 digits
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
                        const uint32_t F = __float_as_uint(fmaf(__fmul_rn(sc, rsm), __uint_as_float(lds32(xa + r * ZG_KR * 4)), 1.5f));
                        const uint32_t pa = planes + lane + r * kPlaneRow;
                        sts8(pa, F);
                        sts8(pa + 32, F >> 8);
                        sts8(pa + 64, F >> 16);
                    }
                }
                __syncwarp();
                // MMA: the ring's shared-memory bytes ARE the A fragments, the planes ARE the B fragments
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
                            const uint4 q0 = lds128(qa);   // row g: 16 u[n + 8] + u[n], row g + 8: u[n]
                            a[0][0] = q0.x; a[0][1] = q0.x & 0x0F0F0F0Fu; a[0][2] = q0.y; a[0][3] = q0.y & 0x0F0F0F0Fu;
                            a[1][0] = q0.z; a[1][1] = q0.z & 0x0F0F0F0Fu; a[1][2] = q0.w; a[1][3] = q0.w & 0x0F0F0F0Fu;
                        }
                        const uint32_t b0 = lds32(brow + r * kPlaneRow), b1 = lds32(brow + r * kPlaneRow + 16);
                        if constexpr (kI4) {
                            dsum = __dp4a(b0, 0x01010101u, __dp4a(b1, 0x01010101u, dsum));
                            imma_u8u8(acc[0], a[0][0], a[0][1], a[0][2], a[0][3], b0, b1);
                            imma_u8u8(acc[1], a[1][0], a[1][1], a[1][2], a[1][3], b0, b1);
                        } else {
                            imma_s8u8(acc[0], a[0][0], a[0][1], a[0][2], a[0][3], b0, b1);
                            imma_s8u8(acc[1], a[1][0], a[1][1], a[1][2], a[1][3], b0, b1);
                        }
                    }
                }
            };
            if (cnt == G) chunk(std::true_type{}); else chunk(std::false_type{});
            __syncwarp();   // every lane is done with the slot and the planes: lane 0 requests the chunk NS ahead
            if (lane == 0 && p_recs) issue_chunk();
            slot_u32 += slot_bytes; bar_u32 += 8;
            if (++slot == NS) { slot = 0; slot_u32 = ring; bar_u32 = bars; parity ^= 1; }
        }

        resolve_pending();   // the previous cut group of this warp: its arrival has long returned

        // ── flush: integer sums -> the 32 column sums of this (column group, segment), 4 per lane with t == 0 ──
        const bool whole = (L == P.n_kc);
        uint32_t ord = 0, nseg = 1;
        if (!whole) {   // pieces of this block of column groups in k order (CTA boundaries, then the staging capacity): mine, and how many
            uint32_t qq = gbl * P.nq, cntp = 0;
            const uint32_t end = qq + P.nq, q_mine = qq + qi;
            while (qq < end) {
                const uint32_t hi = min((qq / P.per + 1) * P.per, end);   // the CTA that holds chunk qq ends here
                if (q_mine >= qq && q_mine < hi) ord = cntp + (q_mine - qq) / P.Lq;
                cntp += (hi - qq + P.Lq - 1) / P.Lq;
                qq = hi;
            }
            nseg = cntp;
        }
        {
            const uint32_t kcnt = L * ZG_KR;   // rows fed to the MMA
            // 2^-23 * E2, E2 = max|x| * smax / 0.499
            const float esc = (mx <= 3.0e38f) ? ((mx >= 1.0e-30f) ? (mx * (2.004008016f * 1.1920928955078125e-07f)) * sm : 0.0f) : __int_as_float(0x7fc00000);
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
            float* dstp = whole ? o.out + (size_t)nb * ZG_TN : o.partials + ((size_t)nb * P.slots + ord) * ZG_TN;
#pragma unroll
            for (int ct = 0; ct < 2; ct++) {
                int pz[4];
#pragma unroll
                for (int i = 0; i < 4; i++) pz[i] = __shfl_xor_sync(0xffffffffu, acc[ct][i], 1);
                if (t == 0) {
                    const int* oo = acc[ct];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        long long T;
                        if constexpr (!kI4) {
                            // own columns 2t, 2t+1 = digits 0, 1; partner's = digit 2 and sum_k q; row g + 8h
                            T = (long long)oo[2 * h] + ((long long)oo[2 * h + 1] << 8) + ((long long)pz[2 * h] << 16) - 12582912LL * (long long)pz[2 * h + 1];
                        } else {
                            // Y (row g + 8) = sum u[n] d ; X (row g) = sum (16 u[n+8] + u[n]) d
                            long long u0, u1, u2, us;
                            if (h == 0) { u0 = oo[2]; u1 = oo[3]; u2 = pz[2]; us = pz[3]; }
                            else { u0 = (oo[0] - oo[2]) >> 4; u1 = (oo[1] - oo[3]) >> 4; u2 = (pz[0] - pz[2]) >> 4; us = (pz[1] - pz[3]) >> 4; }
                            // sum (u - 8)(Mk - 3*2^22) = sum u Mk - 8 sum Mk - 3*2^22 (sum u - 8 kcnt)
                            T = u0 + (u1 << 8) + (u2 << 16) - 8 * dS - 12582912LL * (us - 8LL * (long long)kcnt);
                        }
                        const uint32_t col = ct * 16 + g + 8 * h;
                        if (!whole || nb * ZG_TN + col < P.N) dstp[col] = __ll2float_rn(T) * esc;
                    }
                }
            }
        }
        if (!whole) {
            __syncwarp();   // the warp's partial stores are ordered before lane 0's release
            if (lane == 0)
                asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(pend_old) : "l"(o.counters + nb) : "memory");
            pend_nseg = nseg; pend_nb = nb; pend_i = op_i;
        }
    }
    if (!P.early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    resolve_pending();
#ifdef ZG_STREAM_TRACE_ALL
    if (zt_slot) {   // exit time (48 bits of ns) with the CTA index above it
        unsigned long long tt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
        zt_slot[2] = ((unsigned long long)blockIdx.x << 48) | (tt & 0xFFFFFFFFFFFFull);
    }
#else
    ZG_TRACE_MARK(2)
#endif
}

template <int FMT>
bool launch_stream(const QGemvSParams& p, uint32_t grid, uint32_t smem, cudaStream_t st, bool pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, qgemv_stream_kernel<FMT>, p);
    ZG_COUNT_LAUNCH();
    g_zg_stream_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) { zg_set_error("qgemv stream launch failed: %s (grid %u, %u B shared)", cudaGetErrorString(e), grid, smem); return false; }
    return true;
}

template <int FMT>
bool set_stream_attr() {
    if (cudaFuncSetAttribute(qgemv_stream_kernel<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
        zg_set_error("cudaFuncSetAttribute(qgemv stream) failed"); return false;
    }
    return true;
}

}  // namespace

// once per context, outside any stream capture
bool zg_qgemv_stream_init(ZgCudaCtx* ctx) {
    if (const char* e = getenv("ZG_GEMV_STREAM")) ctx->gemv_stream = atoi(e);            // 0: never, 1: when a launch has the work (default), 2: every M == 1 launch
    if (const char* e = getenv("ZG_GEMV_STREAM_EARLY")) ctx->stream_early = atoi(e);     // 1: griddepcontrol.launch_dependents at kernel entry (A/B)
    if (const char* e = getenv("ZG_GEMV_STREAM_NS")) ctx->stream_ns = atoi(e);           // ring slots per warp (2: three CTAs per SM; 4: two)
    if (const char* e = getenv("ZG_GEMV_STREAM_MIN")) ctx->stream_min_chunks = atoi(e);  // chunks in a launch below which the k-split kernel keeps it
    if (const char* e = getenv("ZG_GEMV_STREAM_ALIGN")) ctx->stream_align = atoi(e);     // 1: always snap, 2: never, 0: snap below two column groups per CTA
    if (const char* e = getenv("ZG_GEMV_STREAM_CHUNKS")) ctx->stream_chunks = atoi(e);   // chunks per warp and CTA at least
    return set_stream_attr<ZG_QFMT_I8_F32>() && set_stream_attr<ZG_QFMT_I8_F16>() && set_stream_attr<ZG_QFMT_I4_F16>();
}

// Plan of the streamed form for `count` matvecs of w's shape in one launch; use == false: the k-split kernel of qgemv.cu
// takes the launch (little work per CTA: its finer split keeps more warps busy).
ZgGemvStreamPlan zg_qgemv_stream_plan(const ZgCudaCtx* ctx, const ZgCudaQWeight* w, uint32_t count) {
    ZgGemvStreamPlan pl;
    if (w->fmt == ZG_QFMT_GENERIC || ctx->gemv_stream == 0 || count == 0) return pl;
    const uint32_t G = w->fmt == ZG_QFMT_I4_F16 ? 4u : 2u;
    pl.GB = (w->n_nb + kWarps - 1) / kWarps;
    pl.nq = (w->n_kc + G - 1) / G;
    pl.TQ = count * pl.GB * pl.nq;
    pl.Lq = kSegRecs / G;
    pl.NS = ctx->stream_ns > 0 ? (uint32_t)ctx->stream_ns : 4;   // 4 slots: two CTAs per SM, 147 KB in flight per SM
    pl.smem_bytes = kSegRecs * ZG_KR * 4 + kWarps * pl.NS * G * w->rec_bytes + kWarps * G * kPlaneRow + kWarps * pl.NS * 8;
    const uint32_t occ = std::max(1u, std::min(3u, (227u * 1024u) / (pl.smem_bytes + 1024u + 256u)));   // CTAs per SM (80 registers allow 3)
    const uint32_t slots_max = (uint32_t)ctx->sm_count * occ;
    // CTAs of `per` chunks per warp, ONE round of CTAs (grid <= resident slots).  `per` is snapped up to a divisor of a column
    // group's chunk count nq (or a multiple of it): CTA boundaries then fall on the same few places of every column group
    // (per == nq: none at all — whole groups, direct stores) and a CTA never straddles two blocks of column groups.  Measured
    // (132 MB Llama-3-70B matvec in a dependent chain): 224 aligned CTAs of 32 chunks 31.3 us, 444 CTAs of 16.1 chunks 36 us
    // (three CTAs per SM stream at unequal speed under saturation and the slow quarter finishes latency-bound).
    const uint32_t per_min = ctx->stream_chunks > 0 ? (uint32_t)ctx->stream_chunks : 16u;
    const uint32_t min_total = ctx->stream_min_chunks > 0 ? (uint32_t)ctx->stream_min_chunks : 3072u;
    if (ctx->gemv_stream == 1 && pl.TQ < min_total) return pl;
    // batches of three or more int8 matvecs already stream at 92-97 % of HBM through the k-split kernel's many short CTAs
    // (measured, 8 x 4096x4096 / 4096x14336: 6.0-6.4 TB/s against 5.7-6.2 here); the int4 batches do not (5.0 -> 5.6 TB/s here)
    if (ctx->gemv_stream == 1 && count >= 3 && w->fmt != ZG_QFMT_I4_F16) return pl;
    uint32_t per = std::max(per_min, (pl.TQ + slots_max - 1) / slots_max);
    if (ctx->stream_align == 1 || (ctx->stream_align == 0 && per < 2 * pl.nq)) {
        if (per < pl.nq && pl.TQ / pl.nq >= 128) per = pl.nq;       // whole column groups still give >= 128 CTAs: no cuts, no partials
        else if (per < pl.nq) { while (pl.nq % per) per++; }          // smallest divisor of nq that is >= per (nq itself at the latest)
        else per = ((per + pl.nq - 1) / pl.nq) * pl.nq;
    }
    pl.per = per;
    pl.grid = (pl.TQ + per - 1) / per;
    // pieces a column group can be cut into: staging capacity + CTA boundaries inside its nq chunks
    pl.slots = (pl.nq + pl.Lq - 1) / pl.Lq + (pl.nq + per - 1) / per + 2;
    pl.use = true;
    return pl;
}

bool zg_qgemv_stream_launch(ZgCudaCtx* ctx, const ZgGemvStreamPlan& pl, uint32_t count, const ZgCudaQWeight* const* ws_w, const float* const* d_in,
                            float* const* d_out, const ZgGemvWs* ws, cudaStream_t st) {
    const ZgCudaQWeight* w0 = ws_w[0];
    QGemvSParams p;
    memset(&p, 0, sizeof(p));
    const size_t pe = (size_t)w0->n_nb * pl.slots * ZG_TN;
    for (uint32_t i = 0; i < count; i++) {
        const ZgCudaQWeight* w = ws_w[i];
        if (w->fmt != w0->fmt || w->K != w0->K || w->N != w0->N) { zg_set_error("internal: mixed matvec batch"); return false; }
        if (ws[i].partials_elems < pe || ws[i].counters_n < w0->n_nb || !ws[i].partials || !ws[i].counters) {
            zg_set_error("internal: stream split workspace too small (%zu/%u needed)", pe, w0->n_nb);
            return false;
        }
        p.op[i] = QGemvSOp{w->recs, w->smax, d_in[i], d_out[i], ws[i].partials, ws[i].counters};
    }
    p.n_kc = w0->n_kc; p.n_nb = w0->n_nb; p.K = (uint32_t)w0->K; p.N = (uint32_t)w0->N;
    p.GB = pl.GB; p.nq = pl.nq; p.TQ = pl.TQ; p.per = pl.per; p.Lq = pl.Lq; p.NS = pl.NS; p.slots = pl.slots; p.early = ctx->stream_early ? 1u : 0u; p.rev = getenv("ZG_GEMV_STREAM_REV") ? 1u : 0u;
    switch (w0->fmt) {
        case ZG_QFMT_I8_F32: return launch_stream<ZG_QFMT_I8_F32>(p, pl.grid, pl.smem_bytes, st, ctx->pdl);
        case ZG_QFMT_I8_F16: return launch_stream<ZG_QFMT_I8_F16>(p, pl.grid, pl.smem_bytes, st, ctx->pdl);
        case ZG_QFMT_I4_F16: return launch_stream<ZG_QFMT_I4_F16>(p, pl.grid, pl.smem_bytes, st, ctx->pdl);
        default: zg_set_error("qmatmul: unknown weight format %d", w0->fmt); return false;
    }
}
