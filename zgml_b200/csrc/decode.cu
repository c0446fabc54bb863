// Fused decode step for sm_100a: the layers of a single-token LLaMA DeviceProgram in ONE persistent kernel.
//
// A decode step is a chain of ~12 dependent small kernels per layer; on B200 each dependent launch costs 3-5 us
// even with programmatic dependent launch, while the HBM time of a whole SmolLM-1.7B layer is 6 us
// (profiles/r01_trace_1p7b_final.txt).  This kernel keeps one CTA of 16 warps resident on every SM for the whole
// step and replaces the launches by grid barriers:
//
//   phase 1  x = a (+ b)            -> rmsnorm -> * gamma         -> q | k | v matvecs        (quant.zig:475-578)
//   phase 2  rope(q), rope(k), KV store, attention over the cache, split over CTAs -> partial (m, l, acc) states
//   phase 3  merge of the partial states -> attn_out / concat buffer -> o matvec
//   phase 4  x = a + o              -> rmsnorm -> * gamma         -> gate | up matvecs
//   phase 5  act(gate) * up                                        -> down matvec
//   (+ one NVLink peer all-reduce phase after phases 3 and 5 when the program is sharded)
//
// Everything that is not a matvec runs in the PROLOGUE of the matvec phase that consumes it (every CTA evaluates the
// small vector ops for the k-range it needs; one designated CTA also stores the absorbed ops' output buffers, so every
// DeviceOp result stays observable exactly as the op-by-op execution leaves it).  Matvec outputs that are split over k
// leave the phase as S partial sums; the consumer adds them in split order (deterministic).
//
// Weights: same packed records, TMA bulk ring and integer-MMA digit-plane arithmetic as qgemv.cu (see there for the
// numerics).  The per-warp rings are fed by a chunk STREAM that runs across phase and layer boundaries: weights are
// immutable, so while a CTA waits at a barrier or evaluates a prologue the next phase's first chunks are already in
// flight (148 SMs x 144 KB of ring = 21 MB, about 3 us of HBM time).
//
// CODE SIZE is a first-class constraint here.  The CTAs run in lockstep and every phase's code executes once per layer,
// so whatever does not fit the instruction caches (L1.5: 32 KB) is re-fetched from L2 by all 148 SMs for every layer:
// the first version (everything unrolled and inlined per call site, 300 KB of SASS) spent 65 us per layer almost entirely
// on instruction fetch.  Hence: one call site per routine (a loop over the four matvec phases), loops instead of
// unrolling wherever latency allows, the chunk producer and the rare paths out of line, descriptor copies by cp.async.
//
// No spin in this kernel is unbounded: grid barrier, mbarrier and peer waits give up after ~2 s, set the sticky error
// word (reported by zg_cuda_execute through zg_cuda_last_error) and let the kernel drain.
#include "zg_internal.cuh"

#include <stdlib.h>
#include <string.h>
#include <type_traits>

ZG_TRACE_DECL
void zg_trace_set_decode(unsigned long long* d_buf) { cudaMemcpyToSymbol(c_zg_trace, &d_buf, sizeof(d_buf)); }

namespace {

constexpr int kW = 16;                         // warps per CTA
constexpr int kT = kW * 32;
constexpr uint32_t kNS = 4;                    // ring slots per warp
constexpr uint32_t kSlotBytes = 2304;          // 4 int4 records (576 B) or 2 int8 records (1088 / 1152 B)
constexpr uint32_t kPlaneRow = 144;            // digit planes of one record: 4 x 32 B + 16 B bank skew
constexpr uint32_t kPlaneBytes = 4 * kPlaneRow;
constexpr uint32_t kAttnMaxDh = 256;
constexpr long long kSpinLimit = 4000000000LL; // clock64 ticks (~2 s) before a wait gives up
enum { ERR_BARRIER = 1, ERR_MBAR = 2, ERR_PEER = 3 };

// per-layer descriptor block: fixed offsets, identical in device memory and in shared memory (three shared buffers:
// the layer being executed, the next one (the weight stream runs ahead into it), and the one being fetched)
constexpr uint32_t kDescLy = 0, kDescPh = 1024, kDescHeads = kDescPh + 768, kDescKvs = kDescHeads + kZgDecMaxHeads * sizeof(ZgDecHead);
constexpr uint32_t kDescGlobalBytes = kDescKvs + kZgDecMaxHeads * sizeof(ZgDecKv);          // what device memory holds per layer
constexpr uint32_t kDescSeq = kDescGlobalBytes, kDescKdst = kDescSeq + kZgDecMaxHeads * 4, kDescVdst = kDescKdst + kZgDecMaxHeads * 4;
constexpr uint32_t kDescBytes = kDescVdst + kZgDecMaxHeads * 4;                             // + the layer's run-time values
static_assert(sizeof(ZgDecLayer) <= 1024 && 4 * sizeof(ZgDecPhase) <= 768 && sizeof(ZgDecHead) == 40 && sizeof(ZgDecKv) == 32, "descriptor block layout");
static_assert(sizeof(ZgDecLayer) % 8 == 0 && sizeof(ZgDecPhase) % 8 == 0 && kDescGlobalBytes % 16 == 0 && kDescGlobalBytes <= 16 * kT, "descriptor copy");

struct StreamState {   // one warp's chunk producer (owned by its lane 0)
    uint32_t pi, item, c, slot, outstanding, valid;      // phase index, item, chunk, ring slot to fill next, in flight, position valid
    uint32_t L, n_chunk, G, RB, n_items, n_slots;        // this warp's share of phase pi
    const uint8_t* src;                                  // records of (item, chunk 0)
    uint32_t k0, n_kc;
};
static_assert(sizeof(StreamState) == 64, "stream state");

// dynamic shared memory layout (bytes)
constexpr uint32_t kOffXin = 0;                                  // kZgDecMaxD floats: the phase's staged input vector
constexpr uint32_t kOffAttn = 0;                                 // (phase 2 only, aliases xin) sq | sk | sv [256], sh_m | sh_l [16], sh_acc [16][256]
constexpr uint32_t kOffRing = kOffXin + kZgDecMaxD * 4;          // [warp][slot][kSlotBytes]
constexpr uint32_t kOffPlanes = kOffRing + kW * kNS * kSlotBytes;
constexpr uint32_t kOffPart = kOffPlanes + kW * kPlaneBytes;     // [2][warp][32] floats
constexpr uint32_t kOffBars = kOffPart + 2 * kW * 32 * 4;        // [warp][slot] mbarriers
constexpr uint32_t kOffRed = kOffBars + kW * kNS * 8;            // 64 floats
constexpr uint32_t kOffSmax = kOffRed + 64 * 4;                  // scale ceilings of this CTA's column groups in the current phase
constexpr uint32_t kOffTrace = kOffSmax + kZgDecMaxItems * 4;    // Trace state (16 B)
constexpr uint32_t kOffStream = kOffTrace + 16;                  // [warp] StreamState
constexpr uint32_t kOffDesc = kOffStream + kW * 64;
constexpr uint32_t kSmemBytes = kOffDesc + 3 * kDescBytes;
static_assert((3 * kAttnMaxDh + 2 * kW + kW * kAttnMaxDh) * 4 <= kZgDecMaxD * 4, "attention scratch aliases the input vector");
static_assert(kSmemBytes <= 227 * 1024, "decode kernel shared memory");
static_assert(kOffDesc % 16 == 0 && kDescBytes % 16 == 0 && kOffStream % 8 == 0, "descriptor buffers are copied 16 bytes at a time and hold pointers");

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
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
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
__device__ __forceinline__ void sts8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_vol(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// value i of a matvec-phase output: complete, or the sum of its S partial sums in split order
__device__ __noinline__ float vec_get(const ZgDecVec& v, uint32_t i) {
    if (v.S == 0) return __ldcg(v.full + i);
    float a = __ldcg(v.part + i);
    for (uint32_t s = 1; s < v.S; s++) a += __ldcg(v.part + (size_t)s * v.n + i);
    return a;
}

// ── timeline (zg_cuda_trace): thread 0 of CTA 0 stamps %globaltimer at named points.  The record slots are reserved with
//    ONE atomic at kernel start (an atomic per point would stall the traced warp for a round trip and distort the picture);
//    record = {kind << 56 | t, t, t}, kind = 64 + 16 * phase + point (scripts/trace_decode.py). ──
struct Trace { unsigned long long* next; unsigned long long* end; };   // lives in shared memory; only thread 0 of CTA 0 uses it
__device__ __noinline__ void trace_open(Trace* t, uint32_t reserve) {
    t->next = nullptr; t->end = nullptr;
    if (!c_zg_trace) return;
    const unsigned long long s = atomicAdd(c_zg_trace, (unsigned long long)reserve);
    if (s >= 16000) return;
    t->next = c_zg_trace + 1 + 3 * s;
    t->end = c_zg_trace + 1 + 3 * min((unsigned long long)16000, s + reserve);
}
__device__ __noinline__ void trace_pt_impl(Trace* t, uint32_t kind) {
    if (!t->next || t->next >= t->end) return;
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    t->next[0] = ((unsigned long long)kind << 56) | (now & 0xFFFFFFFFFFFFFFull);
    t->next[1] = now; t->next[2] = now;
    t->next += 3;
}
#define trace_pt(cx, phase, point) do { if ((cx).trace_on) trace_pt_impl((cx).tr, 64 + 16 * (phase) + (point)); } while (0)

// ── grid barrier: monotonic counter, `target` arrivals expected in total ──
__device__ __noinline__ void grid_barrier(uint32_t* sync, uint32_t target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(sync, 1u);
        uint32_t spins = 0;
        long long t0 = 0;
        while (ld_vol(sync) < target) {
            if ((++spins & 255u) == 0) {
                if (ld_vol(sync + 64)) break;
                if (t0 == 0) t0 = clock64();
                else if (clock64() - t0 > kSpinLimit) { atomicExch(sync + 64, (uint32_t)ERR_BARRIER); break; }
            }
        }
        __threadfence();
    }
    __syncthreads();
}

// ── descriptors: everything a phase needs is in shared memory before the phase starts ──
struct Desc {
    const ZgDecLayer* ly; const ZgDecPhase* ph; const ZgDecHead* heads; const ZgDecKv* kvs;
    uint32_t* seq_kv; uint32_t* k_dst; uint32_t* v_dst;
};
__device__ __forceinline__ Desc desc_of(uint8_t* smem, uint32_t layer) {
    uint8_t* b = smem + kOffDesc + (layer % 3) * kDescBytes;
    Desc d;
    d.ly = reinterpret_cast<const ZgDecLayer*>(b + kDescLy); d.ph = reinterpret_cast<const ZgDecPhase*>(b + kDescPh);
    d.heads = reinterpret_cast<const ZgDecHead*>(b + kDescHeads); d.kvs = reinterpret_cast<const ZgDecKv*>(b + kDescKvs);
    d.seq_kv = reinterpret_cast<uint32_t*>(b + kDescSeq); d.k_dst = reinterpret_cast<uint32_t*>(b + kDescKdst); d.v_dst = reinterpret_cast<uint32_t*>(b + kDescVdst);
    return d;
}
__device__ __forceinline__ const ZgDecPhase* phase_of(uint8_t* smem, uint32_t pi) {
    return reinterpret_cast<const ZgDecPhase*>(smem + kOffDesc + ((pi >> 2) % 3) * kDescBytes + kDescPh) + (pi & 3);
}
// a layer's block: device memory -> shared memory, 16 bytes per thread, asynchronously (cp.async); complete after
// cp.async.wait_all + a block barrier
__device__ __forceinline__ void block_fetch(const ZgDecodePlan& P, uint8_t* smem, uint32_t layer, uint32_t tid) {
    if (tid * 16 < kDescGlobalBytes) {
        const uint32_t dst = smem_u32(smem + kOffDesc + (layer % 3) * kDescBytes) + tid * 16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(P.blocks + (size_t)layer * P.blk_bytes + tid * 16) : "memory");
    }
}
// the layer's run-time values (patched per step by refresh): seq_kv per head, cache store offsets per KV head
__device__ __forceinline__ uint32_t dyn_issue(const ZgDecodePlan& P, const Desc& d, uint32_t tid) {
    const uint32_t n_heads = d.ly->n_heads, n_kv = d.ly->n_kv;
    if (tid < n_heads) return __ldg(P.dyn + d.heads[tid].dyn);
    if (tid >= 64 && tid < 64 + n_kv) return __ldg(P.dyn + d.kvs[tid - 64].k_dyn);
    if (tid >= 128 && tid < 128 + n_kv) return __ldg(P.dyn + d.kvs[tid - 128].v_dyn);
    return 0u;
}
__device__ __forceinline__ void dyn_commit(const Desc& d, uint32_t tid, uint32_t v) {
    const uint32_t n_heads = d.ly->n_heads, n_kv = d.ly->n_kv;
    if (tid < n_heads) d.seq_kv[tid] = v;
    else if (tid >= 64 && tid < 64 + n_kv) d.k_dst[tid - 64] = v;
    else if (tid >= 128 && tid < 128 + n_kv) d.v_dst[tid - 128] = v;
}

// ── work split of a matvec phase: S = 2^lS k-splits x (grid >> lS) column-group slots; a CTA owns split (cta & (S - 1)) of the
//    column groups slot, slot + n_slots, ...; its 16 warps split the k-range ──
struct Geo { uint32_t split, slot, ks, ke, k0, L; bool cta_busy; };
__device__ __forceinline__ Geo phase_geo(const ZgDecPhase* ph, uint32_t cta, uint32_t warp) {
    Geo g;
    const uint32_t lS = ph->lS, n_kc = ph->n_kc;
    g.split = cta & ((1u << lS) - 1); g.slot = cta >> lS;
    g.ks = (g.split * n_kc) >> lS; g.ke = ((g.split + 1) * n_kc) >> lS;
    g.k0 = g.ks + ((warp * (g.ke - g.ks)) >> 4);
    g.L = g.ks + (((warp + 1) * (g.ke - g.ks)) >> 4) - g.k0;
    g.cta_busy = g.slot < ph->n_slots && g.slot < ph->n_items;
    return g;
}
__device__ __forceinline__ uint32_t item_op(const ZgDecPhase* ph, uint32_t item) {
    return (item >= ph->mv[1].first_item ? 1u : 0u) + (item >= ph->mv[2].first_item ? 1u : 0u);
}

// ── the chunk stream of one warp: (phase, item, chunk) in execution order, across phases and layers.  Run by lane 0 only,
//    out of line, state in shared memory.  Requests chunks until the ring is full, the stream ends, or the next chunk belongs
//    to a layer whose descriptors are not in shared memory yet (ready_layer). ──
__device__ __noinline__ void stream_fill(uint8_t* smem, uint32_t cta, uint32_t warp, uint32_t n_ph, uint32_t ready_layer) {
    StreamState* s = reinterpret_cast<StreamState*>(smem + kOffStream) + warp;
    const uint32_t ring = smem_u32(smem + kOffRing) + warp * kNS * kSlotBytes, bars = smem_u32(smem + kOffBars) + warp * kNS * 8;
    while (s->outstanding < kNS) {
        while (!s->valid && s->pi < n_ph && (s->pi >> 2) <= ready_layer) {
            const ZgDecPhase* ph = phase_of(smem, s->pi);
            const Geo g = phase_geo(ph, cta, warp);
            if (g.cta_busy && g.L > 0) {
                s->G = ph->fmt == ZG_QFMT_I4_F16 ? 4u : 2u; s->RB = ph->rec_bytes; s->L = g.L; s->n_chunk = (g.L + s->G - 1) / s->G;
                s->n_items = ph->n_items; s->n_slots = ph->n_slots; s->k0 = g.k0; s->n_kc = ph->n_kc;
                s->item = g.slot; s->c = 0; s->valid = 1;
                const uint32_t o = item_op(ph, s->item);
                s->src = ph->mv[o].recs + ((size_t)(s->item - ph->mv[o].first_item) * s->n_kc + s->k0) * s->RB;
            } else s->pi++;
        }
        if (!s->valid) return;
        const uint32_t cnt = min(s->G, s->L - s->c * s->G), bar = bars + s->slot * 8;
        mbar_expect_tx(bar, cnt * s->RB);
        bulk_g2s(ring + s->slot * kSlotBytes, s->src + (size_t)s->c * s->G * s->RB, cnt * s->RB, bar);
        s->outstanding++;
        if (++s->slot == kNS) s->slot = 0;
        if (++s->c == s->n_chunk) {
            s->c = 0;
            s->item += s->n_slots;
            if (s->item >= s->n_items) { s->valid = 0; s->pi++; }
            else {
                const ZgDecPhase* ph = phase_of(smem, s->pi);
                const uint32_t o = item_op(ph, s->item);
                s->src = ph->mv[o].recs + ((size_t)(s->item - ph->mv[o].first_item) * s->n_kc + s->k0) * s->RB;
            }
        }
    }
}

struct Ctx {          // per-thread constants + ring state
    uint32_t tid, lane, warp, cta, grid;
    uint32_t ring, bars, planes;         // shared-memory addresses of this warp's ring / barriers / digit planes
    uint32_t slot, parity;               // consumer position
    uint32_t pbuf;                       // partial-sum double buffer
    uint32_t n_ph, ready_layer;
    uint32_t* sync;
    uint8_t* smem;
    float* xin; float* part; float* red; float* smax;
    Trace* tr; bool trace_on;
};

// scale ceiling of the i-th column group this CTA owns in the phase (thread i < kZgDecMaxItems): issued before the prologue's
// loads, stored to shared memory after them
__device__ __forceinline__ float smax_issue(const ZgDecPhase* ph, const Geo& g, uint32_t tid) {
    if (!g.cta_busy || tid >= kZgDecMaxItems) return 1.0f;
    const uint32_t item = g.slot + tid * ph->n_slots;
    if (item >= ph->n_items) return 1.0f;
    const uint32_t o = item_op(ph, item);
    return __ldg(ph->mv[o].smax + (item - ph->mv[o].first_item));
}

// ── one matvec phase: the CTA's column groups x its k-split, activations already staged in xin[] (element x_base + i) ──
template <int FMT>
__device__ __forceinline__ void mv_phase(Ctx& cx, const ZgDecPhase* ph, const Geo& g, uint32_t x_base) {
    constexpr bool kI4 = (FMT == ZG_QFMT_I4_F16);
    constexpr bool kF32 = (FMT == ZG_QFMT_I8_F32);
    constexpr uint32_t QB = kI4 ? 512u : 1024u;
    constexpr uint32_t SB = kF32 ? 32u : 16u;
    constexpr uint32_t RB = QB + 4 * SB;
    constexpr uint32_t G = kI4 ? 4u : 2u;
    const uint32_t lane = cx.lane, gq = lane >> 2, t = lane & 3, j = gq & 3;
    const uint32_t n_chunk = (g.L + G - 1) / G, n_items = ph->n_items, n_slots = ph->n_slots;
    const uint32_t S = ph->S;

    // this warp's slice of the staged activations, scaled in place: x' = x * 0.499 / (max|x| * smax) (qgemv.cu)
    float* xw = cx.xin + ((size_t)g.k0 * ZG_KR - x_base);
    float xm = 0.0f;
    float sm_prev = cx.smax[0];
    if (g.L > 0) {
        float mx = 0.0f;
#pragma unroll 1
        for (uint32_t i = 0; i < g.L; i++) {
            const float aa = fabsf(xw[32 * i + lane]);
            mx = (aa <= 3.0e38f) ? fmaxf(mx, aa) : INFINITY;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        xm = mx;
        const float f = (mx <= 3.0e38f && mx >= 1.0e-30f) ? (0.499f / mx) * (1.0f / sm_prev) : 0.0f;
#pragma unroll 1
        for (uint32_t i = 0; i < g.L; i++) xw[32 * i + lane] *= f;
        __syncwarp();
    }
    const uint32_t xs_u32 = smem_u32(xw);
    const uint32_t brow = cx.planes + j * 32 + 4 * t;
    const uint32_t q_off = lane * 16;
    const uint32_t sc_off = QB + (kF32 ? 4u : 2u) * (8 * ((lane & 15) >> 2) + 4 * (lane >> 4) + (lane & 3));

    uint32_t ord = 0;   // ordinal of the item among this CTA's items
#pragma unroll 1
    for (uint32_t item = g.slot; item < n_items; item += n_slots, ord++) {
        const uint32_t o = item_op(ph, item);
        const uint32_t nb = item - ph->mv[o].first_item;
        float* part_w = cx.part + ((size_t)cx.pbuf * kW + cx.warp) * 32;
        if (g.L > 0) {
            const float sm = cx.smax[ord];
            if (sm != sm_prev) {   // re-normalise for this column group's scale ceiling: exact (powers of two)
                const float ratio = sm_prev / sm;
#pragma unroll 1
                for (uint32_t i = 0; i < g.L; i++) xw[32 * i + lane] *= ratio;
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
#pragma unroll 1
            for (uint32_t c = 0; c < n_chunk; c++, xa += G * ZG_KR * 4) {
                const uint32_t cnt = min(G, g.L - c * G);
                const uint32_t slot_u32 = cx.ring + cx.slot * kSlotBytes, bar_u32 = cx.bars + cx.slot * 8;
                {   // bounded wait for the chunk's bytes
                    uint32_t spins = 0;
                    long long t0 = 0;
                    while (!mbar_try(bar_u32, cx.parity)) {
                        if ((++spins & 63u) == 0) {
                            if (t0 == 0) t0 = clock64();
                            else if (clock64() - t0 > kSpinLimit) { atomicExch(cx.sync + 64, (uint32_t)ERR_MBAR); break; }
                        }
                    }
                }
                // digit planes of the chunk's records: lane = k row, F = s * x' + 1.5 in (1, 2), low three bytes = base-256 digits
#pragma unroll 1
                for (uint32_t r = 0; r < cnt; r++) {
                    float sc;
                    if constexpr (kF32) {
                        sc = __uint_as_float(lds32(slot_u32 + sc_off + r * RB));
                    } else {
                        unsigned short h;
                        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(slot_u32 + sc_off + r * RB));
                        sc = __half2float(__ushort_as_half(h));
                    }
                    const uint32_t F = __float_as_uint(fmaf(sc, __uint_as_float(lds32(xa + r * ZG_KR * 4)), 1.5f));
                    const uint32_t pa = cx.planes + lane + r * kPlaneRow;
                    sts8(pa, F);
                    sts8(pa + 32, F >> 8);
                    sts8(pa + 64, F >> 16);
                }
                __syncwarp();
#pragma unroll 1
                for (uint32_t r = 0; r < cnt; r++) {
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
                __syncwarp();
                // the slot is free: request the chunk kNS ahead in the stream (possibly of a later phase / layer)
                if (++cx.slot == kNS) { cx.slot = 0; cx.parity ^= 1; }
                if (lane == 0) {
                    reinterpret_cast<StreamState*>(cx.smem + kOffStream)[cx.warp].outstanding--;
                    stream_fill(cx.smem, cx.cta, cx.warp, cx.n_ph, cx.ready_layer);
                }
            }
            // flush: integer sums -> this warp's float partial of the column group
            {
                const uint32_t kcnt = g.L * ZG_KR;
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
                    if (t == 0) {   // B columns 0..3 = the three digits + the ones plane of the (single) activation row
                        const int* oo = acc[ct];
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            long long T;
                            if constexpr (!kI4) {
                                T = (long long)oo[2 * h] + ((long long)oo[2 * h + 1] << 8) + ((long long)pz[2 * h] << 16) -
                                    12582912LL * (long long)pz[2 * h + 1];
                            } else {
                                long long u0, u1, u2, us;
                                if (h == 0) { u0 = oo[2]; u1 = oo[3]; u2 = pz[2]; us = pz[3]; }
                                else { u0 = (oo[0] - oo[2]) >> 4; u1 = (oo[1] - oo[3]) >> 4; u2 = (pz[0] - pz[2]) >> 4; us = (pz[1] - pz[3]) >> 4; }
                                T = u0 + (u1 << 8) + (u2 << 16) - 8 * dS - 12582912LL * (us - 8LL * (long long)kcnt);
                            }
                            part_w[ct * 16 + gq + 8 * h] = __ll2float_rn(T) * esc;
                        }
                    }
                }
            }
        } else {
            part_w[lane] = 0.0f;
        }
        __syncthreads();
        if (cx.tid < 32) {
            const float* pr = cx.part + (size_t)cx.pbuf * kW * 32 + cx.tid;
            float v = 0.0f;
#pragma unroll
            for (int w = 0; w < kW; w++) v += pr[w * 32];
            float* out = S == 1 ? ph->mv[o].out : ph->mv[o].part + (size_t)g.split * ph->mv[o].N;
            out[nb * ZG_TN + cx.tid] = v;
        }
        cx.pbuf ^= 1;
    }
}

// fixed-order block sum (every CTA computes the identical value)
__device__ __forceinline__ float block_sum(Ctx& cx, float v) {
    v = warp_sum(v);
    __syncthreads();
    if (cx.lane == 0) cx.red[cx.warp] = v;
    __syncthreads();
    float r = 0.0f;
#pragma unroll
    for (int i = 0; i < kW; i++) r += cx.red[i];
    return r;
}

// ── prologue of phases 1 and 4: x = a (+ b -> sum) ; bare = x * inv_rms ; grep = gamma ; norm = bare * gamma -> xin[0 .. n_pad).
//    Four elements per thread and pass, their loads in flight together; the sums wait in xin[] for the scale. ──
__device__ __forceinline__ void prologue_norm(Ctx& cx, const float* a, const ZgDecVec& b, float* sum, const float* gamma, float* bare,
                                              float* grep, float* norm, float eps, uint32_t D, uint32_t n_pad) {
    constexpr int NU = 4;
    const bool has_b = b.full != nullptr, writer = cx.cta == 0;
    const float* bsrc = b.S ? b.part : b.full;
    float ss = 0.0f;
#pragma unroll 1
    for (uint32_t e0 = 0; e0 < D; e0 += NU * kT) {
        float v[NU];
#pragma unroll
        for (int u = 0; u < NU; u++) { const uint32_t i = e0 + cx.tid + u * kT; v[u] = i < D ? __ldcg(a + i) : 0.0f; }
        if (has_b) {
            float t2[NU];
#pragma unroll
            for (int u = 0; u < NU; u++) { const uint32_t i = e0 + cx.tid + u * kT; t2[u] = i < D ? __ldcg(bsrc + i) : 0.0f; }
#pragma unroll 1
            for (uint32_t s = 1; s < b.S; s++) {
#pragma unroll
                for (int u = 0; u < NU; u++) { const uint32_t i = e0 + cx.tid + u * kT; if (i < D) t2[u] += __ldcg(b.part + (size_t)s * b.n + i); }
            }
#pragma unroll
            for (int u = 0; u < NU; u++) {
                const uint32_t i = e0 + cx.tid + u * kT;
                v[u] = __fadd_rn(v[u], t2[u]);
                if (writer && i < D) { sum[i] = v[u]; if (b.S) b.full[i] = t2[u]; }
            }
        }
#pragma unroll
        for (int u = 0; u < NU; u++) {
            const uint32_t i = e0 + cx.tid + u * kT;
            if (i < D) cx.xin[i] = v[u];
            ss = fmaf(v[u], v[u], ss);
        }
    }
    const float tot = block_sum(cx, ss);
    const float inv_rms = 1.0f / sqrtf(tot / (float)D + eps);
#pragma unroll 4
    for (uint32_t i = cx.tid; i < n_pad; i += kT) {
        float xv = 0.0f;
        if (i < D) {
            const float bz = __fmul_rn(cx.xin[i], inv_rms), gm = __ldg(gamma + i);
            xv = __fmul_rn(bz, gm);
            if (writer) { bare[i] = bz; grep[i] = gm; norm[i] = xv; }
        }
        cx.xin[i] = xv;
    }
}

__device__ __forceinline__ uint32_t attn_splits(uint32_t seq_kv, uint32_t max_splits) {
    const uint32_t s = (seq_kv + 31u) / 32u;
    return s < 1u ? 1u : (s > max_splits ? max_splits : s);
}

// ── prologue of phase 3: merge the split-KV partial states of the heads this CTA's k-range covers.  First the weight of
//    every (head, split) state — exp(m_s - max) / sum — by one lane each (16 lanes per head, shuffles inside the half warp),
//    then every element is a weighted sum of its splits' accumulators. ──
__device__ __forceinline__ void prologue_attn_merge(Ctx& cx, const ZgDecodePlan& P, const Desc& d, const Geo& g, uint32_t K) {
    const ZgDecLayer& ly = *d.ly;
    const uint32_t n = (g.ke - g.ks) * ZG_KR, dh = ly.d_head, stride = 2 + P.part_dh;
    const uint32_t k_lo = g.ks * ZG_KR, k_hi = min(k_lo + n, K);
    const uint32_t h_lo = k_lo / dh, n_h = k_hi > k_lo ? (k_hi - 1) / dh - h_lo + 1 : 0;
    float* wsm = cx.part;   // [head - h_lo][16] weights (the partial-sum buffers are idle during the prologue)
#pragma unroll 1
    for (uint32_t p0 = 0; p0 < n_h * 16; p0 += kT) {
        const uint32_t pidx = p0 + cx.tid, h = h_lo + pidx / 16, sp = pidx & 15;
        const bool in = pidx < n_h * 16 && sp < attn_splits(d.seq_kv[min(h, ly.n_heads - 1)], P.max_splits);
        const float* pp = P.attn_part + ((size_t)h * P.max_splits + sp) * stride;
        const float m = in ? __ldcg(pp) : -INFINITY, lsum = in ? __ldcg(pp + 1) : 0.0f;
        float gmax = m;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
        const float e = (m == -INFINITY) ? 0.0f : expf(m - gmax);
        float lt = lsum * e;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) lt += __shfl_xor_sync(0xffffffffu, lt, o);
        if (pidx < n_h * 16) wsm[pidx] = lt > 0.0f ? e / lt : 0.0f;
    }
    __syncthreads();
    const bool writer = g.slot == 0;
#pragma unroll 1
    for (uint32_t i = cx.tid; i < n; i += kT) {
        const uint32_t k = k_lo + i;
        float val = 0.0f;
        if (k < K) {
            const uint32_t h = k / dh, dd = k - h * dh;
            const uint32_t splits = attn_splits(d.seq_kv[h], P.max_splits);
            const float* base = P.attn_part + (size_t)h * P.max_splits * stride + 2 + dd;
            const float* w = wsm + (h - h_lo) * 16;
#pragma unroll 1
            for (uint32_t sp0 = 0; sp0 < splits; sp0 += 4) {
                float a4[4];
#pragma unroll
                for (int q = 0; q < 4; q++) a4[q] = sp0 + q < splits ? __ldcg(base + (size_t)(sp0 + q) * stride) : 0.0f;
#pragma unroll
                for (int q = 0; q < 4; q++) val = fmaf(a4[q], w[min(sp0 + q, 15u)], val);
            }
            if (writer) { d.heads[h].attn_out[dd] = val; ly.attn_buf[d.heads[h].buf_off + dd] = val; }
        }
        cx.xin[i] = val;
    }
    __syncthreads();   // wsm (the partial-sum buffers) is free again
}

__device__ __noinline__ float apply_unary(uint32_t op, float v) {
    switch (op) {
        case ZG_EW_NEG: return -v;
        case ZG_EW_ABS: return fabsf(v);
        case ZG_EW_RELU: return fmaxf(v, 0.0f);
        case ZG_EW_SQRT: return sqrtf(v);
        case ZG_EW_RECIP: return 1.0f / v;
        case ZG_EW_EXP: return expf(v);
        case ZG_EW_LOG: return logf(v);
        case ZG_EW_GELU: { const float kk = 0.7978845608f * (v + 0.044715f * v * v * v); return 0.5f * v * (1.0f + tanhf(kk)); }
        default: return v;
    }
}
// the general activation chain of one element (fused_elementwise steps, src/backend/reference.zig:262-300); out of line
__device__ __noinline__ float act_chain(const ZgDecLayer* ly, float gt, float up, float sec0, int first_ext, uint32_t k) {
    float v = gt;
    for (uint32_t s = 0; s < ly->n_steps; s++) {
        const ZgDecStep st = ly->steps[s];
        if (st.op == ZG_EW_ADD || st.op == ZG_EW_MUL) {
            const float o2 = st.sec_kind == 1 ? gt : (st.sec_kind == 2 ? up : ((int)s == first_ext ? sec0 : __ldg(st.sec + k)));
            if (st.op == ZG_EW_ADD) v = st.is_swapped ? __fadd_rn(o2, v) : __fadd_rn(v, o2);
            else v = st.is_swapped ? __fmul_rn(o2, v) : __fmul_rn(v, o2);
        } else v = apply_unary(st.op, v);
    }
    return v;
}

// ── prologue of phase 5: mid = steps(gate) ; hidden = mid * up for the CTA's k-range.  The SiLU chain of the lowering (neg,
//    exp, + ones, recip, * gate: src/nn.zig:38-44) is evaluated inline, op by op with the same roundings; any other chain by
//    the general interpreter. ──
__device__ __forceinline__ void prologue_act(Ctx& cx, const Desc& d, const Geo& g) {
    constexpr int NU = 4;
    const ZgDecLayer& ly = *d.ly;
    const uint32_t n = (g.ke - g.ks) * ZG_KR, F = ly.F, k_lo = g.ks * ZG_KR;
    const bool writer = g.slot == 0;
    int first_ext = -1;
#pragma unroll 1
    for (uint32_t s = 0; s < ly.n_steps; s++)
        if ((ly.steps[s].op == ZG_EW_ADD || ly.steps[s].op == ZG_EW_MUL) && ly.steps[s].sec_kind == 0) { first_ext = (int)s; break; }
    const float* gsrc = ly.gate.S ? ly.gate.part : ly.gate.full;
    const float* usrc = ly.up.S ? ly.up.part : ly.up.full;
    const float* esrc = first_ext >= 0 ? ly.steps[first_ext].sec : nullptr;
    const uint32_t Smax = max(ly.gate.S, ly.up.S);
#pragma unroll 1
    for (uint32_t e0 = 0; e0 < n; e0 += NU * kT) {
        float gt[NU], up[NU], sec0[NU];
#pragma unroll
        for (int u = 0; u < NU; u++) {
            const uint32_t i = e0 + cx.tid + u * kT, k = k_lo + i;
            const bool in = i < n && k < F;
            gt[u] = in ? __ldcg(gsrc + k) : 0.0f;
            up[u] = in ? __ldcg(usrc + k) : 0.0f;
            sec0[u] = (in && esrc) ? __ldg(esrc + k) : 0.0f;
        }
#pragma unroll 1
        for (uint32_t sx = 1; sx < Smax; sx++) {
#pragma unroll
            for (int u = 0; u < NU; u++) {
                const uint32_t i = e0 + cx.tid + u * kT, k = k_lo + i;
                if (i < n && k < F) {
                    if (sx < ly.gate.S) gt[u] += __ldcg(ly.gate.part + (size_t)sx * ly.gate.n + k);
                    if (sx < ly.up.S) up[u] += __ldcg(ly.up.part + (size_t)sx * ly.up.n + k);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < NU; u++) {
            const uint32_t i = e0 + cx.tid + u * kT, k = k_lo + i;
            if (i < n) {
                float hv = 0.0f;
                if (k < F) {
                    float v;
                    if (ly.act_silu) v = __fmul_rn(gt[u], 1.0f / __fadd_rn(expf(-gt[u]), sec0[u]));
                    else v = act_chain(d.ly, gt[u], up[u], sec0[u], first_ext, k);
                    hv = __fmul_rn(v, up[u]);
                    if (writer) {
                        ly.silu[k] = v; ly.hidden[k] = hv;
                        if (ly.gate.S) ly.gate.full[k] = gt[u];
                        if (ly.up.S) ly.up.full[k] = up[u];
                    }
                }
                cx.xin[i] = hv;
            }
        }
    }
}

// ── phase 2: one (head, kv-split) item per CTA round.  Scores: lane = kv position (LPP lanes share a position when the
//    range is short), V: lane = head dimension — the structure of ops.cu k_attention_fast, written for small code; the
//    position written by this step is served from shared memory (its rope'd key / value are computed here), the cache
//    rows are stored by the split-0 item of each KV head's first query head.  The cache rows were pulled into L2 during
//    phase 1 (attention_prefetch). ──
template <int NI>
__device__ __forceinline__ void attention_item(Ctx& cx, const ZgDecodePlan& P, const Desc& d, uint32_t h, uint32_t sp, uint32_t splits) {
    const ZgDecLayer& ly = *d.ly;
    float* sq = reinterpret_cast<float*>(cx.smem + kOffAttn);
    float* sk = sq + kAttnMaxDh;
    float* sv = sk + kAttnMaxDh;
    float* sh_m = sv + kAttnMaxDh;
    float* sh_l = sh_m + kW;
    float* sh_acc = sh_l + kW;   // [kW][kAttnMaxDh]
    const ZgDecHead hd = d.heads[h];
    const ZgDecKv kv = d.kvs[hd.kv];
    const uint32_t dh = ly.d_head, hd2 = dh >> 1, tid = cx.tid, lane = cx.lane, warp = cx.warp;
    const uint32_t seq_kv = d.seq_kv[h];
    const uint32_t k_dst = d.k_dst[hd.kv], v_dst = d.v_dst[hd.kv];
    // the kv position this step writes, as this head's attention op numbers its rows (UINT32_MAX: not one of its rows)
    uint32_t s_new = UINT32_MAX;
    if (k_dst >= hd.k_off && (k_dst - hd.k_off) % ly.k_cs == 0 && v_dst >= hd.v_off && (v_dst - hd.v_off) % ly.v_cs == 0 &&
        (k_dst - hd.k_off) / ly.k_cs == (v_dst - hd.v_off) / ly.v_cs)
        s_new = (k_dst - hd.k_off) / ly.k_cs;
    const bool store = sp == 0 && ((h == 0) || (d.heads[h - 1].kv != hd.kv));
    if (tid < NI * 32) {
        const uint32_t r = tid;
        float qv = 0.0f, kvv = 0.0f, vv = 0.0f;
        if (r < dh) {
            const bool lo = r < hd2;
            const uint32_t pair = lo ? r : r - hd2;
            const float c = __ldcg(ly.cs + pair), sn = __ldcg(ly.cs + pair + hd2);
            // this element and its rotation partner; separate roundings like the reference (reference.zig:474-475)
            const float q_me = vec_get(ly.q, hd.q_src + r), q_pt = vec_get(ly.q, hd.q_src + (lo ? r + hd2 : pair));
            const float k_me = vec_get(ly.k, kv.k_src + r), k_pt = vec_get(ly.k, kv.k_src + (lo ? r + hd2 : pair));
            vv = vec_get(ly.v, kv.v_src + r);
            qv = lo ? __fsub_rn(__fmul_rn(q_me, c), __fmul_rn(q_pt, sn)) : __fadd_rn(__fmul_rn(q_me, c), __fmul_rn(q_pt, sn));
            kvv = lo ? __fsub_rn(__fmul_rn(k_me, c), __fmul_rn(k_pt, sn)) : __fadd_rn(__fmul_rn(k_me, c), __fmul_rn(k_pt, sn));
            if (sp == 0) {
                hd.q_rot[r] = qv;
                if (ly.q.S) ly.q.full[hd.q_src + r] = q_me;
            }
            if (store) {
                kv.k_rot[r] = kvv;
                ly.k_cache[(size_t)k_dst + r] = kvv;
                ly.v_cache[(size_t)v_dst + r] = vv;
                if (ly.k.S) ly.k.full[kv.k_src + r] = k_me;
                if (ly.v.S) ly.v.full[kv.v_src + r] = vv;
            }
        }
        sq[r] = qv; sk[r] = kvv; sv[r] = vv;
    }
    __syncthreads();
    const uint32_t dh4 = dh >> 2;
    float acc[NI];
#pragma unroll
    for (int i = 0; i < NI; i++) acc[i] = 0.0f;
    float m_val = -INFINITY, l = 0.0f;
    const uint32_t chunk = ((seq_kv + splits - 1) / splits + 31) & ~31u;
    const uint32_t kv_lo = sp * chunk, kv_hi = min(kv_lo + chunk, seq_kv);
    uint32_t lpp = 1;
    {
        const uint32_t len = kv_hi > kv_lo ? kv_hi - kv_lo : 0;
        while (lpp < 8 && len <= (kW * 32u) / (2 * lpp) && (dh4 % (2 * lpp)) == 0) lpp *= 2;
    }
    const uint32_t pw = 32 / lpp, seg = lane & (lpp - 1), pos_in_warp = lane / lpp;
    const uint32_t f4 = dh4 / lpp;
    const float* kbase = ly.k_cache + hd.k_off;
    const float* vbase = ly.v_cache + hd.v_off;
#pragma unroll 1
    for (uint32_t s0 = kv_lo + warp * pw; s0 < kv_hi; s0 += kW * pw) {
        const uint32_t s = s0 + pos_in_warp;
        const bool in_range = s < kv_hi;
        float mask_add = (in_range && ly.has_mask) ? __ldcg(ly.mask + ly.mask_off + (size_t)s * ly.mask_rs) : 0.0f;
        if (!in_range) mask_add = -INFINITY;
        float dot = 0.0f;
        if (in_range) {
            const float4* kr = reinterpret_cast<const float4*>(s == s_new ? sk : kbase + (size_t)s * ly.k_cs) + seg * f4;
            const float4* q4 = reinterpret_cast<const float4*>(sq) + seg * f4;
            float d0 = 0.0f, d1 = 0.0f;
#pragma unroll 4
            for (uint32_t dq = 0; dq < f4; dq++) {
                const float4 a = kr[dq], qa = q4[dq];
                d0 = fmaf(qa.x, a.x, d0); d1 = fmaf(qa.y, a.y, d1); d0 = fmaf(qa.z, a.z, d0); d1 = fmaf(qa.w, a.w, d1);
            }
            dot = d0 + d1;
        }
        for (uint32_t o = 1; o < lpp; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        float score = -INFINITY;
        bool ok = false;
        if (isfinite(mask_add)) {
            score = dot * ly.scale + mask_add;
            ok = isfinite(score);
            if (!ok) score = -INFINITY;
        }
        float bm = score;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
        if (bm == -INFINITY) continue;   // warp-uniform
        const float new_m = fmaxf(m_val, bm);
        const float alpha = (m_val == -INFINITY) ? 0.0f : expf(m_val - new_m);
        const float wgt = ok ? expf(score - new_m) : 0.0f;
        l = l * alpha + warp_sum(seg == 0 ? wgt : 0.0f);
        m_val = new_m;
#pragma unroll
        for (int i = 0; i < NI; i++) acc[i] *= alpha;
        const uint32_t nj = min(pw, kv_hi - s0);
#pragma unroll 1
        for (uint32_t j0 = 0; j0 < nj; j0 += 8) {
            float vvv[8][NI];
#pragma unroll
            for (int jj = 0; jj < 8; jj++) {
                const bool in = j0 + jj < nj;
                const uint32_t sj = s0 + j0 + jj;
                const float* vr = (sj == s_new) ? sv : vbase + (size_t)sj * ly.v_cs;
#pragma unroll
                for (int i = 0; i < NI; i++) vvv[jj][i] = (in && lane + 32 * i < dh) ? vr[lane + 32 * i] : 0.0f;
            }
#pragma unroll
            for (int jj = 0; jj < 8; jj++) {
                const float wj = __shfl_sync(0xffffffffu, wgt, ((j0 + jj) * lpp) & 31);
#pragma unroll
                for (int i = 0; i < NI; i++) acc[i] = fmaf(wj, vvv[jj][i], acc[i]);
            }
        }
    }
    if (lane == 0) { sh_m[warp] = m_val; sh_l[warp] = l; }
#pragma unroll
    for (int i = 0; i < NI; i++) sh_acc[warp * kAttnMaxDh + lane + 32 * i] = acc[i];
    __syncthreads();
    // merge the warps' states in warp order: thread r < dh owns output element r (thread dh: the row sum)
    if (tid <= dh) {
        float gm = -INFINITY;
#pragma unroll 1
        for (int w = 0; w < kW; w++) gm = fmaxf(gm, sh_m[w]);
        float a = 0.0f;
#pragma unroll 1
        for (int w = 0; w < kW; w++) {
            const float ws = (sh_m[w] == -INFINITY) ? 0.0f : expf(sh_m[w] - gm);
            a += (tid < dh ? sh_acc[w * kAttnMaxDh + tid] : sh_l[w]) * ws;
        }
        float* mine = P.attn_part + ((size_t)h * P.max_splits + sp) * (2 + P.part_dh);
        if (tid < dh) mine[2 + tid] = a;
        else { mine[0] = gm; mine[1] = a; }
    }
    __syncthreads();   // the shared buffers are reused by this CTA's next item
}

__device__ __forceinline__ void attention_phase(Ctx& cx, const ZgDecodePlan& P, const Desc& d) {
    const ZgDecLayer& ly = *d.ly;
#pragma unroll 1
    for (uint32_t it = cx.cta; it < ly.n_heads * P.max_splits; it += cx.grid) {
        const uint32_t h = it / P.max_splits, sp = it % P.max_splits;
        const uint32_t splits = attn_splits(d.seq_kv[h], P.max_splits);
        if (sp >= splits) continue;   // CTA-uniform
        if (ly.d_head <= 64) attention_item<2>(cx, P, d, h, sp, splits);
        else if (ly.d_head <= 128) attention_item<4>(cx, P, d, h, sp, splits);
        else attention_item<8>(cx, P, d, h, sp, splits);
    }
}
// pull the cache rows of this CTA's phase-2 items into L2 while phase 1 runs (rows below the position written by this step
// are final since the previous step)
__device__ __forceinline__ void attention_prefetch(Ctx& cx, const ZgDecodePlan& P, const Desc& d) {
    const ZgDecLayer& ly = *d.ly;
#pragma unroll 1
    for (uint32_t it = cx.cta; it < ly.n_heads * P.max_splits; it += cx.grid) {
        const uint32_t h = it / P.max_splits, sp = it % P.max_splits;
        const uint32_t seq_kv = d.seq_kv[h], splits = attn_splits(seq_kv, P.max_splits);
        if (sp >= splits || (h > 0 && d.heads[h - 1].kv == d.heads[h].kv)) continue;   // one prefetch per KV head and split
        const uint32_t chunk = ((seq_kv + splits - 1) / splits + 31) & ~31u;
        const uint32_t kv_lo = sp * chunk, kv_hi = min(kv_lo + chunk, seq_kv);
        const char* kb = reinterpret_cast<const char*>(ly.k_cache + d.heads[h].k_off);
        const char* vb = reinterpret_cast<const char*>(ly.v_cache + d.heads[h].v_off);
#pragma unroll 1
        for (uint32_t s = kv_lo + cx.tid; s < kv_hi; s += kT)
#pragma unroll 1
            for (uint32_t b = 0; b < ly.d_head * 4; b += 128) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(kb + (size_t)s * ly.k_cs * 4 + b));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(vb + (size_t)s * ly.v_cs * 4 + b));
            }
    }
}

// ── all-reduce phase over NVLink peer memory: the protocol of ops.cu k_allreduce_peer (16-byte cells {x, epoch, y, epoch}),
//    run by the first kZgPeerCtas CTAs; local values come from the matvec's partial sums ──
__device__ __noinline__ void allreduce_phase(Ctx& cx, const ZgDecodePlan& P, const ZgDecVec& local, float* out, uint32_t& ar_seq) {
    if (cx.cta >= kZgPeerCtas) return;
    const ZgPeerComm& pc = P.pc;
    const uint32_t n2 = local.n >> 1;
    const uint32_t set = ar_seq % kZgPeerSets, epoch = ar_seq + 1;
    const uint32_t chunk = (n2 + kZgPeerCtas - 1) / kZgPeerCtas, lo = cx.cta * chunk, hi = min(lo + chunk, n2);
    const size_t slot_pairs = pc.max_n >> 1;
    const size_t my_cell = ((size_t)set * pc.world + pc.rank) * slot_pairs;
    const uint4* mine = reinterpret_cast<const uint4*>(pc.slots[pc.rank]) + (size_t)set * pc.world * slot_pairs;
    for (uint32_t jx = lo + cx.tid; jx < hi; jx += kT) {
        const float2 own = make_float2(vec_get(local, 2 * jx), vec_get(local, 2 * jx + 1));
        for (int pr = 0; pr < pc.world; pr++) {
            if (pr == pc.rank) continue;
            uint4* cell = reinterpret_cast<uint4*>(pc.slots[pr]) + my_cell + jx;
            asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(cell), "r"(__float_as_uint(own.x)), "r"(epoch),
                         "r"(__float_as_uint(own.y)), "r"(epoch) : "memory");
        }
        uint4 got[kZgMaxRanks];
        uint32_t pending = 0;
#pragma unroll
        for (int r = 0; r < kZgMaxRanks; r++)
            if (r < pc.world && r != pc.rank) pending |= 1u << r;
        long long t0 = 0;
        uint32_t spins = 0;
        bool dead = false;
        while (pending) {
#pragma unroll
            for (int r = 0; r < kZgMaxRanks; r++)
                if (pending & (1u << r)) {
                    const uint4* cell = mine + (size_t)r * slot_pairs + jx;
                    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(got[r].x), "=r"(got[r].y), "=r"(got[r].z), "=r"(got[r].w) : "l"(cell) : "memory");
                }
#pragma unroll
            for (int r = 0; r < kZgMaxRanks; r++)
                if ((pending & (1u << r)) && got[r].y == epoch && got[r].w == epoch) pending &= ~(1u << r);
            if (pending && (++spins & 255u) == 0) {
                if (t0 == 0) t0 = clock64();
                else if (clock64() - t0 > 4 * kSpinLimit || ld_vol(cx.sync + 64)) { atomicExch(cx.sync + 64, (uint32_t)ERR_PEER); dead = true; break; }
            }
        }
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kZgMaxRanks; r++)
            if (r < pc.world) {
                const float2 v = (r == pc.rank) ? own : make_float2(__uint_as_float(got[r].x), __uint_as_float(got[r].z));
                acc = (r == 0) ? v : make_float2(acc.x + v.x, acc.y + v.y);
            }
        if (dead) acc = make_float2(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000));   // a peer never arrived: poison, never garbage
        reinterpret_cast<float2*>(out)[jx] = acc;
    }
    ar_seq++;
}

__global__ void __launch_bounds__(kT, 1)
k_decode_layers(const __grid_constant__ ZgDecodePlan P) {
    extern __shared__ __align__(128) uint8_t smem[];
    Ctx cx;
    cx.tid = threadIdx.x; cx.lane = cx.tid & 31; cx.warp = cx.tid >> 5; cx.cta = blockIdx.x; cx.grid = gridDim.x;
    const uint32_t smem_base = smem_u32(smem);
    cx.ring = smem_base + kOffRing + cx.warp * kNS * kSlotBytes;
    cx.bars = smem_base + kOffBars + cx.warp * kNS * 8;
    cx.planes = smem_base + kOffPlanes + cx.warp * kPlaneBytes;
    cx.slot = 0; cx.parity = 0; cx.pbuf = 0;
    cx.sync = P.sync; cx.smem = smem;
    cx.xin = reinterpret_cast<float*>(smem + kOffXin);
    cx.part = reinterpret_cast<float*>(smem + kOffPart);
    cx.red = reinterpret_cast<float*>(smem + kOffRed);
    cx.smax = reinterpret_cast<float*>(smem + kOffSmax);
    cx.n_ph = 4 * P.n_layers;
    const uint32_t L = P.n_layers;
    cx.tr = reinterpret_cast<Trace*>(smem + kOffTrace);
    cx.trace_on = false;
    if (cx.cta == 0 && cx.tid == 0) { trace_open(cx.tr, 24 * L + 8); cx.trace_on = cx.tr->next != nullptr; }

    // descriptor blocks of layers 0 and 1
    block_fetch(P, smem, 0, cx.tid);
    if (L > 1) block_fetch(P, smem, 1, cx.tid);
    // ring barriers, producer state and the constant ones plane (digit index 3) of every record slot
    if (cx.lane == 0) {
        for (uint32_t s = 0; s < kNS; s++) mbar_init(cx.bars + s * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        StreamState* st = reinterpret_cast<StreamState*>(smem + kOffStream) + cx.warp;
        st->pi = 0; st->item = 0; st->c = 0; st->slot = 0; st->outstanding = 0; st->valid = 0;
    }
    for (uint32_t i = cx.lane; i < 4 * 8; i += 32)
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(cx.planes + (i >> 3) * kPlaneRow + 96 + (i & 7) * 4), "r"(0x01010101u) : "memory");
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    cx.ready_layer = L > 1 ? 1u : 0u;
    if (cx.lane == 0) stream_fill(smem, cx.cta, cx.warp, cx.n_ph, cx.ready_layer);
    {   // layer 0's run-time values
        const Desc d0 = desc_of(smem, 0);
        dyn_commit(d0, cx.tid, dyn_issue(P, d0, cx.tid));
        __syncthreads();
    }

    uint32_t ar_seq = 0;
    if (P.pc.world > 1 && cx.cta < kZgPeerCtas) ar_seq = ld_vol(P.pc.seq + 2 + cx.cta);
    uint32_t arrivals = 0;   // grid-barrier arrivals expected so far

#pragma unroll 1
    for (uint32_t l = 0; l < L; l++) {
        const Desc d = desc_of(smem, l);
        const ZgDecLayer& ly = *d.ly;
        // One loop over the four matvec phases (0: q|k|v, 1: o, 2: gate|up, 3: down): every routine has exactly one call site.
#pragma unroll 1
        for (uint32_t k = 0; k < 4; k++) {
            if (k == 1) {   // phase 2: attention
                trace_pt(cx, 3, 8);
                attention_phase(cx, P, d);
                trace_pt(cx, 3, 9);
                arrivals += cx.grid; grid_barrier(P.sync, arrivals);
            }
            trace_pt(cx, k, 0);
            const ZgDecPhase* ph = d.ph + k;
            const Geo g = phase_geo(ph, cx.cta, cx.warp);
            const float smv = smax_issue(ph, g, cx.tid);
            uint32_t ndyn = 0;
            uint32_t x_base = 0;
            if (k == 0 || k == 2) {
                if (k == 0) {
                    // fetches that ride along with the prologue's loads: layer l + 2's descriptors, layer l + 1's run-time values
                    if (l + 2 < L) block_fetch(P, smem, l + 2, cx.tid);
                    if (l + 1 < L) ndyn = dyn_issue(P, desc_of(smem, l + 1), cx.tid);
                    attention_prefetch(cx, P, d);
                }
                const bool first = k == 0;
                prologue_norm(cx, first ? ly.x1_a : ly.x2_a, first ? ly.x1_b : ly.o, first ? ly.x1_sum : ly.x2_sum, first ? ly.gamma1 : ly.gamma2,
                              first ? ly.bare1 : ly.bare2, first ? ly.grep1 : ly.grep2, first ? ly.norm1 : ly.norm2, first ? ly.eps1 : ly.eps2,
                              ly.D, ph->n_kc * ZG_KR);
            } else {
                x_base = g.ks * ZG_KR;
                if (g.cta_busy) {
                    if (k == 1) prologue_attn_merge(cx, P, d, g, ph->K);
                    else prologue_act(cx, d, g);
                }
            }
            if (cx.tid < kZgDecMaxItems) cx.smax[cx.tid] = smv;
            if (k == 0) {
                if (l + 1 < L) dyn_commit(desc_of(smem, l + 1), cx.tid, ndyn);
                asm volatile("cp.async.wait_all;" ::: "memory");
            }
            __syncthreads();
            if (k == 0 && l + 2 < L) cx.ready_layer = l + 2;
            if (cx.lane == 0) stream_fill(smem, cx.cta, cx.warp, cx.n_ph, cx.ready_layer);
            trace_pt(cx, k, 2);
            if (g.cta_busy) {   // CTA-uniform
                if (ph->fmt == ZG_QFMT_I4_F16) mv_phase<ZG_QFMT_I4_F16>(cx, ph, g, x_base);
                else if (ph->fmt == ZG_QFMT_I8_F16) mv_phase<ZG_QFMT_I8_F16>(cx, ph, g, x_base);
                else mv_phase<ZG_QFMT_I8_F32>(cx, ph, g, x_base);
            }
            trace_pt(cx, k, 10);
            arrivals += cx.grid; grid_barrier(P.sync, arrivals);
            trace_pt(cx, k, 11);
            if ((k == 1 && ly.ar_o) || (k == 3 && ly.ar_down)) {
                allreduce_phase(cx, P, k == 1 ? ly.o_local : ly.down_local, k == 1 ? ly.o.full : ly.down.full, ar_seq);
                trace_pt(cx, k, 12);
                arrivals += cx.grid; grid_barrier(P.sync, arrivals);
            }
        }
    }
    // the last layer's down projection leaves the kernel as a complete vector
    {
        const ZgDecLayer& ly = *desc_of(smem, L - 1).ly;
        if (ly.down.S) {
            for (uint32_t i = cx.cta * kT + cx.tid; i < ly.down.n; i += cx.grid * kT) ly.down.full[i] = vec_get(ly.down, i);
        }
    }
    if (P.pc.world > 1 && cx.cta < kZgPeerCtas && cx.tid == 0) *(volatile uint32_t*)(P.pc.seq + 2 + cx.cta) = ar_seq;
    // re-arm the barrier counter for the next launch: the last CTA to leave resets it
    __syncthreads();
    if (cx.tid == 0) {
        __threadfence();
        const uint32_t old = atomicAdd(P.sync + 32, 1u);
        if (old == cx.grid - 1) { P.sync[0] = 0u; P.sync[32] = 0u; __threadfence(); }
    }
}

} // namespace

bool zg_decode_init(ZgCudaCtx*) {
    cudaError_t e = cudaFuncSetAttribute(k_decode_layers, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) { zg_set_error("cudaFuncSetAttribute(decode) failed: %s", cudaGetErrorString(e)); return false; }
    return true;
}

uint32_t zg_decode_grid(const ZgCudaCtx* ctx) { return (uint32_t)ctx->sm_count; }
uint32_t zg_decode_block_bytes() { return kDescGlobalBytes; }

// host side of the descriptor block layout
void zg_decode_block_fill(uint8_t* block, const ZgDecLayer& ly, const ZgDecPhase* ph4, const ZgDecHead* heads, uint32_t n_heads,
                          const ZgDecKv* kvs, uint32_t n_kv) {
    memset(block, 0, kDescGlobalBytes);
    memcpy(block + kDescLy, &ly, sizeof(ly));
    memcpy(block + kDescPh, ph4, 4 * sizeof(ZgDecPhase));
    memcpy(block + kDescHeads, heads, n_heads * sizeof(ZgDecHead));
    memcpy(block + kDescKvs, kvs, n_kv * sizeof(ZgDecKv));
}

bool zg_decode_launch(ZgCudaCtx* ctx, const ZgDecodeHost& d, cudaStream_t st) {
    (void)ctx;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(d.plan.grid);
    cfg.blockDim = dim3(kT);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: the grid barriers cannot deadlock on residency
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_decode_layers, d.plan);
    ZG_COUNT_LAUNCH();
    if (e != cudaSuccess) { zg_set_error("decode kernel launch failed: %s", cudaGetErrorString(e)); return false; }
    return true;
}

void zg_decode_free(ZgDecodeHost* d) {
    cudaFree(d->d_blocks); cudaFree(d->d_part); cudaFree(d->d_attn_part); cudaFree(d->d_sync);
    if (d->h_err) cudaFreeHost(d->h_err);
    *d = ZgDecodeHost();
}
