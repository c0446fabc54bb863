// Decode-path quantized matvec / small-batch matmul for sm_100a.
//
//   dst[m, n] = sum_k ( x[m, k] * s[(k*N + n) / 32] ) * q[k, n]
//
// i.e. zgml's W8·f32 algorithm (QuantizedWeight.matmul, src/quant.zig:475-578;
// DeviceOp.qmatmul, src/backend/reference.zig:499-566) on the packed GPU-resident
// records of zg_internal.cuh.  The kernel is HBM-bound by construction:
//
//  * Stream-K work split: the matrix is one contiguous run of 2-4 KB records
//    ([n_tile][k_chunk]); CTA c owns records [c*T/G, (c+1)*T/G).  A producer warp
//    streams the run through a ring of shared-memory slots with cp.async.bulk
//    (TMA bulk copy, mbarrier complete_tx), 8 records per copy; 8 consumer warps
//    take one record each per slot.
//  * The weights never pass through a dequantize step.  c[k] = x[k] * s[k, nb]
//    (the reference's first rounding, `scale * input_v`) is computed once per
//    32 weights, converted to 23-bit fixed point in a per-(row, quant-block)
//    scale E2 ~ 2 max|x| max(s) with one FFMA (F = c / E2 + 1.5 -> the three low
//    bytes of F are base-256 digits), and the digits multiply the raw int8 / int4
//    weights on the integer tensor-core path (mma.sync.m16n8k32.s8.u8, IMMA): the
//    record's shared-memory bytes are loaded straight into A fragments.  A column
//    of ones in B yields sum_k q[k, n], which removes the digit bias.  Integer
//    accumulation is exact; the only rounding beyond the reference's is the 2^-23
//    fixed-point grid of c relative to max|x| * max(s).
//  * Deterministic: per-warp partials are combined in fixed order, split tiles go
//    through a partials buffer + arrival counter (last CTA sums in CTA order).
#include "zg_internal.cuh"

namespace {

constexpr int kWarps = 8;                    // consumer warps
constexpr int kThreads = (kWarps + 1) * 32;  // + one producer warp
constexpr int kSlotRecs = kWarps;            // records per ring slot (one per consumer warp)
constexpr int kMaxTl = 3;                    // column tiles one CTA run may touch
constexpr uint32_t kRingBytes = 76 * 1024;

struct QGemvParams {
    const uint8_t* recs;
    const float* smax;
    uint32_t rec_bytes, q_bytes, n_kc, n_tiles;
    uint32_t K, N, M;
    const float* x;
    uint32_t x_rs;
    float* out;
    uint32_t out_rs;
    uint32_t total_recs, n_ring, max_contrib;
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// D(16x8, s32) += A(16x32, s8: weights) * B(32x8, u8: digits of x*s)
__device__ __forceinline__ void imma_16832(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__host__ __device__ __forceinline__ uint32_t run_begin(uint32_t cta, uint32_t grid, uint32_t total) {
    return (uint32_t)(((uint64_t)cta * total) / grid);
}

// FMT: ZG_QFMT_I8_F32 / ZG_QFMT_I8_F16 / ZG_QFMT_I4_F16.  MP: pairs of activation rows (M <= 2*MP).
template <int FMT, int MP>
__global__ void __launch_bounds__(kThreads, MP <= 2 ? 2 : 1)
qgemv_kernel(const QGemvParams p) {
    constexpr bool kI4 = (FMT == ZG_QFMT_I4_F16);
    constexpr int MR = 2 * MP;  // activation rows handled per launch
    constexpr uint32_t kPlaneBytes = MR * 2 * 3 * 32;  // per warp: [m][nb][digit][k]

    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t r0 = run_begin(blockIdx.x, gridDim.x, p.total_recs);
    const uint32_t r1 = run_begin(blockIdx.x + 1, gridDim.x, p.total_recs);
    const uint32_t n_slots = (r1 - r0 + kSlotRecs - 1) / kSlotRecs;
    const uint32_t slot_bytes = kSlotRecs * p.rec_bytes;
    const uint32_t tile_first = r0 / p.n_kc, tile_last = (r1 - 1) / p.n_kc;

    uint8_t* ring = smem;
    uint8_t* planes = ring + (size_t)p.n_ring * slot_bytes;
    float* part = reinterpret_cast<float*>(planes + kWarps * kPlaneBytes);  // [tl][warp][m][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(part + kMaxTl * kWarps * MR * ZG_TN);
    uint64_t* full = bars;
    uint64_t* empty = bars + p.n_ring;
    __shared__ float s_xmax[kWarps][MR];
    __shared__ float s_smax[2 * kMaxTl];
    __shared__ uint32_t s_last[kMaxTl], s_clo[kMaxTl], s_nc[kMaxTl];

    if (tid == 0) {
        for (uint32_t s = 0; s < p.n_ring; s++) {
            mbar_init(smem_u32(&full[s]), 1);
            mbar_init(smem_u32(&empty[s]), kWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = tid; i < kMaxTl * kWarps * MR * ZG_TN; i += kThreads) part[i] = 0.0f;
    __syncthreads();  // barriers initialised: the producer starts streaming before anyone touches x
    // Programmatic dependent launch: the next kernel in the stream may begin (and prefetch ITS weights, which
    // nobody writes) as soon as SM resources free up.  Everything mutable (x, out, partials, counters) is
    // only touched after griddepcontrol.wait, i.e. after the previous kernel has fully completed.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == kWarps) {
        if (lane >= 1 && lane <= kMaxTl) {
            // CTAs whose runs intersect the records [lo, hi] of tile tile_first + lane - 1
            const uint32_t tile = tile_first + lane - 1;
            if (tile <= tile_last) {
                const uint32_t lo = tile * p.n_kc, hi = (tile + 1) * p.n_kc - 1;
                uint32_t c_lo = (uint32_t)(((uint64_t)lo * gridDim.x) / p.total_recs);
                while (run_begin(c_lo + 1, gridDim.x, p.total_recs) <= lo) c_lo++;
                while (run_begin(c_lo, gridDim.x, p.total_recs) > lo) c_lo--;
                uint32_t c_hi = (uint32_t)(((uint64_t)hi * gridDim.x) / p.total_recs);
                while (run_begin(c_hi + 1, gridDim.x, p.total_recs) <= hi) c_hi++;
                while (run_begin(c_hi, gridDim.x, p.total_recs) > hi) c_hi--;
                s_clo[lane - 1] = c_lo;
                s_nc[lane - 1] = c_hi - c_lo + 1;
            }
        }
        // ── producer: stream this CTA's run of records through the ring ──
        if (lane == 0) {
            const uint8_t* src = p.recs + (size_t)r0 * p.rec_bytes;
            uint32_t left = r1 - r0, slot = 0, phase = 1;  // phase of the `empty` wait: first pass never waits
            const uint32_t ring_u32 = smem_u32(ring), full_u32 = smem_u32(full), empty_u32 = smem_u32(empty);
            for (uint32_t s = 0; s < n_slots; s++) {
                if (s >= p.n_ring) mbar_wait(empty_u32 + slot * 8, phase);
                const uint32_t nrec = min((uint32_t)kSlotRecs, left);
                const uint32_t bytes = nrec * p.rec_bytes;
                mbar_expect_tx(full_u32 + slot * 8, bytes);
                bulk_g2s(ring_u32 + slot * slot_bytes, src, bytes, full_u32 + slot * 8);
                src += slot_bytes; left -= nrec;
                if (++slot == p.n_ring) { slot = 0; phase ^= 1; }
                if (s + 1 == p.n_ring || s + 1 == n_slots) asm volatile("griddepcontrol.wait;" ::: "memory");
            }
        }
        if (lane != 0) asm volatile("griddepcontrol.wait;" ::: "memory");
    } else {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        // ── consumers ──
        // record cursors: `cur` is the record this warp takes in the current slot, `pre` runs two slots
        // ahead for the activation prefetch (32 k of x per record, L2 hits)
        uint32_t cur_rec = r0 + warp, cur_tile_of = cur_rec / p.n_kc, cur_kc = cur_rec - cur_tile_of * p.n_kc;
        uint32_t pre_rec = cur_rec, pre_kc = cur_kc;
        float xv[MR], xn1[MR], xn2[MR];
        auto fetch_x = [&](float (&dst)[MR]) {
            const uint32_t k = pre_kc * ZG_KR + lane;
            const bool ok = pre_rec < r1 && k < p.K;
#pragma unroll
            for (int m = 0; m < MR; m++) dst[m] = (ok && m < (int)p.M) ? __ldg(p.x + (size_t)m * p.x_rs + k) : 0.0f;
            pre_rec += kSlotRecs; pre_kc += kSlotRecs;
            while (pre_kc >= p.n_kc) pre_kc -= p.n_kc;
        };
        fetch_x(xn1);
        fetch_x(xn2);
        // max |x[m, :]| (non-finite activations poison the row: outputs become NaN like the reference's)
        float xmax[MR];
#pragma unroll
        for (int m = 0; m < MR; m++) {
            float mx = 0.0f;
            if (m < (int)p.M) {
                const float* xr = p.x + (size_t)m * p.x_rs;
                for (uint32_t k0 = tid; k0 < p.K; k0 += kWarps * 32 * 16) {  // 16 independent loads in flight per thread
                    float a[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const uint32_t k = k0 + i * kWarps * 32;
                        a[i] = (k < p.K) ? fabsf(__ldg(xr + k)) : 0.0f;
                    }
#pragma unroll
                    for (int i = 0; i < 16; i++) mx = (a[i] <= 3.0e38f) ? fmaxf(mx, a[i]) : INFINITY;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (lane == 0) s_xmax[warp][m] = mx;
        }
        if (tid < 2 * kMaxTl) {
            const uint32_t nbi = tile_first * 2 + tid;
            s_smax[tid] = (nbi < p.n_tiles * 2) ? __ldg(p.smax + nbi) : 0.0f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kWarps * 32) : "memory");
#pragma unroll
        for (int m = 0; m < MR; m++) {
            float v = 0.0f;
#pragma unroll
            for (int w = 0; w < kWarps; w++) v = fmaxf(v, s_xmax[w][m]);
            xmax[m] = v;
        }
        const uint32_t g = lane >> 2, t = lane & 3;
        uint8_t* my_planes = planes + warp * kPlaneBytes;
        uint8_t* st_plane = my_planes + lane;                          // c-gen: lane = k
        const bool is_digit = (g & 3) < 3;                             // B column g: digit j of row g / 4, or the ones column
        const uint32_t* ld_plane = reinterpret_cast<const uint32_t*>(my_planes + (((g >> 2) * 2) * 3 + (g & 3)) * 32) + t;
        const uint32_t b_const = (g == 3) ? 0x01010101u : 0u;
        int acc[4][MP][4];
#pragma unroll
        for (int ct = 0; ct < 4; ct++)
#pragma unroll
            for (int mp = 0; mp < MP; mp++)
#pragma unroll
                for (int i = 0; i < 4; i++) acc[ct][mp][i] = 0;
        float inv_e2[MR][2], e2[MR][2];
        uint32_t cur_tile = 0xffffffffu;

        auto flush = [&](uint32_t tile) {
            float* dstp = part + ((size_t)(tile - tile_first) * kWarps + warp) * MR * ZG_TN;
#pragma unroll
            for (int ct = 0; ct < 4; ct++) {
                const int nb = ct >> 1;
#pragma unroll
                for (int mp = 0; mp < MP; mp++) {
                    int x1[4], s3a, s3b;
#pragma unroll
                    for (int i = 0; i < 4; i++) x1[i] = __shfl_xor_sync(0xffffffffu, acc[ct][mp][i], 1);
                    s3a = __shfl_xor_sync(0xffffffffu, acc[ct][mp][1], 3);
                    s3b = __shfl_xor_sync(0xffffffffu, acc[ct][mp][3], 3);
                    if ((t & 1) == 0) {
                        const int m = 2 * mp + (int)(t >> 1);
                        const float esc = (t == 0) ? e2[2 * mp][nb] : e2[2 * mp + 1][nb];
#pragma unroll
                        for (int h = 0; h < 2; h++) {  // rows g and g + 8 of the 16-column tile
                            const int d0 = acc[ct][mp][2 * h], d1 = acc[ct][mp][2 * h + 1], d2 = x1[2 * h];
                            const int sv = (t == 0) ? x1[2 * h + 1] : (h ? s3b : s3a);
                            // sum_k (mant_k - 2^22) q_k with every integer small enough to convert exactly
                            float T = fmaf((float)sv, 32896.0f, (float)(d0 - 128 * sv));
                            T = fmaf((float)(d1 - 128 * sv), 256.0f, T);
                            T = fmaf((float)(d2 - 192 * sv), 65536.0f, T);
                            const float y = (T * (kI4 ? 7.450580596923828e-09f : 1.1920928955078125e-07f)) * esc;
                            if (m < (int)p.M) dstp[m * ZG_TN + ct * 16 + g + 8 * h] = y;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; i++) acc[ct][mp][i] = 0;
                }
            }
        };

        uint32_t slot = 0, phase = 0;
        const uint32_t full_u32 = smem_u32(full), empty_u32 = smem_u32(empty);
        const uint8_t* recp = ring + (size_t)warp * p.rec_bytes;
        for (uint32_t s = 0; s < n_slots; s++) {
#pragma unroll
            for (int m = 0; m < MR; m++) { xv[m] = xn1[m]; xn1[m] = xn2[m]; }
            fetch_x(xn2);
            const bool have = cur_rec < r1;
            if (have && cur_tile_of != cur_tile) {
                if (cur_tile != 0xffffffffu) flush(cur_tile);
                cur_tile = cur_tile_of;
#pragma unroll
                for (int m = 0; m < MR; m++)
#pragma unroll
                    for (int nb = 0; nb < 2; nb++) {
                        // E2 = max|c| / 0.499 >= 2 max|x*s|: F = c / E2 + 1.5 stays inside (1, 2)
                        const float pm = xmax[m] * s_smax[(cur_tile - tile_first) * 2 + nb];
                        const bool finite = pm <= 3.0e38f;             // false for inf and NaN
                        const bool usable = finite && pm >= 1.0e-30f;  // products below 1e-30 flush to zero
                        inv_e2[m][nb] = usable ? __fdividef(0.499f, pm) : 0.0f;
                        e2[m][nb] = finite ? (usable ? pm * 2.004008016f : 0.0f) : __int_as_float(0x7fc00000);
                    }
            }
            mbar_wait(full_u32 + slot * 8, phase);
            if (have) {
                // ── c = x*s -> fixed point digits -> byte planes [m][nb][digit][k] ──
                float sc[2];
                if constexpr (FMT == ZG_QFMT_I8_F32) {
                    const float* sp = reinterpret_cast<const float*>(recp + p.q_bytes);
                    sc[0] = sp[lane]; sc[1] = sp[32 + lane];
                } else {
                    const __half* sp = reinterpret_cast<const __half*>(recp + p.q_bytes);
                    sc[0] = __half2float(sp[lane]); sc[1] = __half2float(sp[32 + lane]);
                }
#pragma unroll
                for (int m = 0; m < MR; m++)
#pragma unroll
                    for (int nb = 0; nb < 2; nb++) {
                        const float c = __fmul_rn(sc[nb], xv[m]);                 // reference.zig:548 `scale * input_v`
                        const uint32_t F = __float_as_uint(fmaf(c, inv_e2[m][nb], 1.5f));
                        uint8_t* pl = st_plane + ((m * 2 + nb) * 3) * 32;
                        pl[0] = (uint8_t)F; pl[32] = (uint8_t)(F >> 8); pl[64] = (uint8_t)(F >> 16);
                    }
                __syncwarp();
                uint32_t bfr[MP][2][2];
#pragma unroll
                for (int mp = 0; mp < MP; mp++)
#pragma unroll
                    for (int nb = 0; nb < 2; nb++) {
                        const uint32_t* pl = ld_plane + ((mp * 4 + nb) * 3 * 32) / 4;   // rows 2*mp (+ g / 4), block nb
                        bfr[mp][nb][0] = is_digit ? pl[0] : b_const;   // ones column: sum_k q
                        bfr[mp][nb][1] = is_digit ? pl[4] : b_const;
                    }
                // ── weights: shared memory bytes are the A fragments ──
                if constexpr (!kI4) {
#pragma unroll
                    for (int ct = 0; ct < 4; ct++) {
                        const uint4 v = *reinterpret_cast<const uint4*>(recp + ct * 512 + lane * 16);
                        const uint32_t a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int mp = 0; mp < MP; mp++) imma_16832(acc[ct][mp], a, bfr[mp][ct >> 1][0], bfr[mp][ct >> 1][1]);
                    }
                } else {
#pragma unroll
                    for (int pr = 0; pr < 2; pr++) {
                        const uint4 v = *reinterpret_cast<const uint4*>(recp + pr * 512 + lane * 16);
                        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int c2 = 0; c2 < 2; c2++) {
                            const int ct = 2 * pr + c2;
                            const uint32_t w0 = w4[2 * c2], w1 = w4[2 * c2 + 1];
                            // two's-complement nibble in the high half of each byte = 16*q as s8
                            const uint32_t a[4] = {(w0 << 4) & 0xF0F0F0F0u, w0 & 0xF0F0F0F0u, (w1 << 4) & 0xF0F0F0F0u, w1 & 0xF0F0F0F0u};
#pragma unroll
                            for (int mp = 0; mp < MP; mp++) imma_16832(acc[ct][mp], a, bfr[mp][ct >> 1][0], bfr[mp][ct >> 1][1]);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_u32 + slot * 8);
            cur_rec += kSlotRecs; cur_kc += kSlotRecs;
            while (cur_kc >= p.n_kc) { cur_kc -= p.n_kc; cur_tile_of++; }
            recp += slot_bytes;
            if (++slot == p.n_ring) { slot = 0; phase ^= 1; recp = ring + (size_t)warp * p.rec_bytes; }
        }
        if (cur_tile != 0xffffffffu) flush(cur_tile);
    }
    __syncthreads();

    // ── combine the 8 warps (fixed order); whole tiles go straight out, split tiles via partials ──
    // partials layout: [tile][j = cta - c_lo(tile)][m][col]
    const uint32_t n_tl = tile_last - tile_first + 1;
    bool any_partial = false;
    for (uint32_t tl = 0; tl < n_tl; tl++) {
        const uint32_t tile = tile_first + tl;
        const bool whole = (r0 <= tile * p.n_kc) && (r1 >= (tile + 1) * p.n_kc);
        any_partial |= !whole;
        float* pdst = p.partials + ((size_t)tile * p.max_contrib + (blockIdx.x - s_clo[tl])) * MR * ZG_TN;
        for (uint32_t idx = tid; idx < p.M * ZG_TN; idx += kThreads) {
            const uint32_t m = idx / ZG_TN, col = idx % ZG_TN;
            float v = 0.0f;
#pragma unroll
            for (int w = 0; w < kWarps; w++) v += part[((size_t)tl * kWarps + w) * MR * ZG_TN + m * ZG_TN + col];
            const uint32_t n = tile * ZG_TN + col;
            if (whole) {
                if (n < p.N) p.out[(size_t)m * p.out_rs + n] = v;
            } else {
                pdst[m * ZG_TN + col] = v;
            }
        }
    }
    if (!any_partial) return;
    __syncthreads();
    if (tid < n_tl) {
        const uint32_t tile = tile_first + tid;
        const bool whole = (r0 <= tile * p.n_kc) && (r1 >= (tile + 1) * p.n_kc);
        uint32_t last = 0;
        if (!whole) {
            // release: publishes the whole CTA's partials (ordered before this thread by the barrier above);
            // acquire: the last arriver sees every earlier contributor's partials
            uint32_t old;
            asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(p.counters + tile) : "memory");
            last = (old == s_nc[tid] - 1) ? 1u : 0u;
            if (last) p.counters[tile] = 0u;  // re-arm for the next launch
        }
        s_last[tid] = last;
    }
    __syncthreads();
    for (uint32_t tl = 0; tl < n_tl; tl++) {
        if (!s_last[tl]) continue;
        const uint32_t tile = tile_first + tl, n_c = s_nc[tl];
        const float* psrc = p.partials + (size_t)tile * p.max_contrib * MR * ZG_TN;
        // contributors summed in CTA order; loads issued in independent batches of 8
        for (uint32_t idx = tid; idx < p.M * ZG_TN; idx += kThreads) {
            const uint32_t m = idx / ZG_TN, col = idx % ZG_TN;
            const uint32_t n = tile * ZG_TN + col;
            float v = 0.0f;
            for (uint32_t j0 = 0; j0 < n_c; j0 += 8) {
                float t8[8];
#pragma unroll
                for (int j = 0; j < 8; j++)
                    t8[j] = (j0 + j < n_c) ? __ldcg(psrc + (size_t)(j0 + j) * MR * ZG_TN + m * ZG_TN + col) : 0.0f;
#pragma unroll
                for (int j = 0; j < 8; j++) v += t8[j];
            }
            if (n < p.N) p.out[(size_t)m * p.out_rs + n] = v;
        }
    }
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

template <int MP>
constexpr uint32_t extra_smem() {  // planes + per-warp partials + barriers
    return kWarps * (2 * MP) * 2 * 3 * 32 + kMaxTl * kWarps * (2 * MP) * ZG_TN * 4 + 2 * 16 * 8;
}

template <int FMT, int MP>
bool set_smem_attr() {
    cudaError_t e = cudaFuncSetAttribute(qgemv_kernel<FMT, MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { zg_set_error("cudaFuncSetAttribute(qgemv) failed: %s", cudaGetErrorString(e)); return false; }
    return true;
}

template <int FMT, int MP>
void launch_fast(const ZgGemvPlan& plan, const QGemvParams& p, cudaStream_t st, bool pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(plan.grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = plan.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, qgemv_kernel<FMT, MP>, p);
    if (e != cudaSuccess) zg_set_error("qgemv launch failed: %s", cudaGetErrorString(e));
    ZG_COUNT_LAUNCH();
}

template <int FMT>
bool launch_fmt(const ZgGemvPlan& plan, const QGemvParams& p, cudaStream_t st, bool pdl) {
    switch (plan.mp) {
        case 1: launch_fast<FMT, 1>(plan, p, st, pdl); return true;
        case 2: launch_fast<FMT, 2>(plan, p, st, pdl); return true;
        case 4: launch_fast<FMT, 4>(plan, p, st, pdl); return true;
        default: zg_set_error("qmatmul: bad row-pair count %u", plan.mp); return false;
    }
}

} // namespace

// Work split for one launch of up to 8 activation rows.
ZgGemvPlan zg_qgemv_plan(const ZgCudaCtx* ctx, const ZgCudaQWeight* w, uint32_t M) {
    ZgGemvPlan pl;
    const uint32_t rows = M > 8 ? 8 : M;
    pl.mp = rows <= 2 ? 1 : (rows <= 4 ? 2 : 4);
    const uint32_t total = w->n_tiles * w->n_kc;
    // One CTA per SM: the kernel is built for 2 resident CTAs (MP <= 2), the second slot is left to the NEXT
    // kernel in the stream, which (programmatic dependent launch) prefetches its weights under this one.
    uint32_t grid = (uint32_t)ctx->sm_count;
    const uint32_t by_work = (total + kSlotRecs - 1) / kSlotRecs;  // at least one record per warp
    if (grid > by_work) grid = by_work;
    const uint32_t min_grid = (total + 2 * w->n_kc) / (2 * w->n_kc + 1);  // a run may touch <= kMaxTl tiles
    if (grid < min_grid) grid = min_grid;
    if (grid < 1) grid = 1;
    if (grid > total) grid = total;
    pl.grid = grid;
    pl.max_contrib = w->n_kc / (total / grid) + 2;  // runs of >= floor(total/grid) records intersecting one tile
    const uint32_t run_max = (total + grid - 1) / grid;
    const uint32_t slot_bytes = kSlotRecs * w->rec_bytes;
    uint32_t n_ring = kRingBytes / slot_bytes;
    const uint32_t need = (run_max + kSlotRecs - 1) / kSlotRecs;
    if (n_ring > need) n_ring = need;
    if (n_ring < 1) n_ring = 1;
    if (n_ring > 16) n_ring = 16;
    pl.n_ring = n_ring;
    const uint32_t extra = pl.mp == 1 ? extra_smem<1>() : (pl.mp == 2 ? extra_smem<2>() : extra_smem<4>());
    pl.smem_bytes = n_ring * slot_bytes + extra;
    return pl;
}

// Opt every instantiation into >48 KB dynamic shared memory once per context,
// outside any stream capture.
bool zg_qgemv_init(ZgCudaCtx*) {
    return set_smem_attr<ZG_QFMT_I8_F32, 1>() && set_smem_attr<ZG_QFMT_I8_F32, 2>() && set_smem_attr<ZG_QFMT_I8_F32, 4>() &&
           set_smem_attr<ZG_QFMT_I8_F16, 1>() && set_smem_attr<ZG_QFMT_I8_F16, 2>() && set_smem_attr<ZG_QFMT_I8_F16, 4>() &&
           set_smem_attr<ZG_QFMT_I4_F16, 1>() && set_smem_attr<ZG_QFMT_I4_F16, 2>() && set_smem_attr<ZG_QFMT_I4_F16, 4>();
}

void zg_qgemv_ws_need(const ZgCudaCtx* ctx, const ZgCudaQWeight* w, uint32_t M, size_t* partial_elems,
                      size_t* counters) {
    *partial_elems = 0; *counters = 0;
    if (w->fmt == ZG_QFMT_GENERIC || M == 0) return;
    ZgGemvPlan plan = zg_qgemv_plan(ctx, w, M);
    *partial_elems = (size_t)w->n_tiles * plan.max_contrib * (2 * plan.mp) * ZG_TN;
    *counters = w->n_tiles;
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
    size_t pe = 0, nc = 0;
    zg_qgemv_ws_need(ctx, w, M, &pe, &nc);
    if (!ws || ws->partials_elems < pe || ws->counters_n < nc) {
        zg_set_error("internal: split workspace too small (%zu/%zu needed)", pe, nc);
        return false;
    }
    for (uint32_t m0 = 0; m0 < M; m0 += 8) {  // larger batches: 8 rows per pass (prefill uses the GEMM path)
        const uint32_t rows = (M - m0) > 8 ? 8 : (M - m0);
        ZgGemvPlan plan = zg_qgemv_plan(ctx, w, rows);
        QGemvParams p;
        p.recs = w->recs; p.smax = w->smax; p.rec_bytes = w->rec_bytes; p.q_bytes = w->q_bytes;
        p.n_kc = w->n_kc; p.n_tiles = w->n_tiles;
        p.K = (uint32_t)w->K; p.N = (uint32_t)w->N; p.M = rows;
        p.x = d_in + (size_t)m0 * in_rs; p.x_rs = in_rs;
        p.out = d_out + (size_t)m0 * out_rs; p.out_rs = out_rs;
        p.total_recs = w->n_tiles * w->n_kc; p.n_ring = plan.n_ring; p.max_contrib = plan.max_contrib;
        p.partials = ws->partials; p.counters = ws->counters;
        bool ok;
        switch (w->fmt) {
            case ZG_QFMT_I8_F32: ok = launch_fmt<ZG_QFMT_I8_F32>(plan, p, st, ctx->pdl); break;
            case ZG_QFMT_I8_F16: ok = launch_fmt<ZG_QFMT_I8_F16>(plan, p, st, ctx->pdl); break;
            case ZG_QFMT_I4_F16: ok = launch_fmt<ZG_QFMT_I4_F16>(plan, p, st, ctx->pdl); break;
            default: zg_set_error("qmatmul: unknown weight format %d", w->fmt); ok = false;
        }
        if (!ok) return false;
    }
    return true;
}
