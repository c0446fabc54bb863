// The non-quantized DeviceOps a compiled zgml program contains around the
// qmatmul anchors (elementwise, fused chains, softmax, layernorm, rmsnorm, reduce,
// repeat, slice_assign, rope, attention, dense matmul).  Semantics follow the
// reference executor, src/backend/reference.zig:201-497,568-672, op by op.
//
// Two fields change between executions (src/device_inference.zig:242-256):
// slice_assign.dst_offset and attention.seq_kv.  They are read from a small
// device array `dyn[op_index]` so the whole op list can live in one CUDA graph.
#include "zg_internal.cuh"

#include <string.h>

ZG_TRACE_DECL
void zg_trace_set_ops(unsigned long long* d_buf) { cudaMemcpyToSymbol(c_zg_trace, &d_buf, sizeof(d_buf)); }

bool g_zg_pdl = true;   // programmatic dependent launch for every kernel of a program (ZG_CUDA_PDL=0 disables)

namespace {

// Programmatic dependent launch, both directions: let the NEXT kernel of the stream become resident right away (a
// matvec then streams its immutable weights while this kernel still runs), and hold THIS kernel — launched early the
// same way — until the previous one has fully completed.  Nothing mutable is touched before the wait.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_zg_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
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
template <bool IS_MAX>
__device__ float block_reduce(float v, float* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = IS_MAX ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = IS_MAX ? -INFINITY : 0.0f;
    for (int i = 0; i < nw; i++) r = IS_MAX ? fmaxf(r, sh[i]) : r + sh[i];
    return r;
}

__device__ __forceinline__ float gelu_tanh(float a) { // reference.zig:265-269
    float kk = 0.7978845608f * (a + 0.044715f * a * a * a);
    return 0.5f * a * (1.0f + tanhf(kk));
}

__device__ __forceinline__ float apply_unary(uint32_t op, float v) {
    switch (op) {
        case ZG_EW_NEG: return -v;
        case ZG_EW_ABS: return fabsf(v);
        case ZG_EW_RELU: return fmaxf(v, 0.0f);
        case ZG_EW_SQRT: return sqrtf(v);
        case ZG_EW_RECIP: return 1.0f / v;
        case ZG_EW_EXP: return expf(v);
        case ZG_EW_LOG: return logf(v);
        case ZG_EW_GELU: return gelu_tanh(v);
        default: return v; // sgn/step/others: reference copies src0 / leaves v unchanged
    }
}

__global__ void k_elementwise(uint32_t op, float* __restrict__ dst, const float* __restrict__ s0,
                              const float* __restrict__ s1, uint32_t n) {
    ZG_TRACE_BEGIN(1)
    pdl_enter();
    ZG_TRACE_MARK(1)
#pragma unroll 4
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float a = s0[i];
        float r;
        if (op == ZG_EW_ADD) r = a + s1[i];
        else if (op == ZG_EW_MUL) r = a * s1[i];
        else r = apply_unary(op, a);
        dst[i] = r;
    }
    ZG_TRACE_MARK(2)
}

__global__ void k_fused_elementwise(const ZgDevStep* __restrict__ steps, uint32_t n_steps,
                                    float* __restrict__ dst, const float* __restrict__ src, uint32_t n) {
    ZG_TRACE_BEGIN(2)
    pdl_enter();
    ZG_TRACE_MARK(1)
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float v = src[i];
        for (uint32_t s = 0; s < n_steps; s++) {
            const ZgDevStep st = steps[s];
            if (st.op == ZG_EW_ADD) { float o = st.sec[i]; v = st.is_swapped ? o + v : v + o; }
            else if (st.op == ZG_EW_MUL) { float o = st.sec[i]; v = st.is_swapped ? o * v : v * o; }
            else v = apply_unary(st.op, v);
        }
        dst[i] = v;
    }
    ZG_TRACE_MARK(2)
}

// [a + b ->] sum ; rmsnorm(sum) -> bare ; gamma broadcast -> gamma_rep ; bare * gamma_rep -> norm for LONG rows (4096 < cols <=
// 8192): one CTA of 1024 threads per row, the row held in registers (2 float4 per thread), one load round, one
// block reduction, one store round.  Shorter rows run inside the chain kernel (kZgChainFusedNorm).
__global__ void __launch_bounds__(1024)
k_norm_macro(const ZgNormMacro m) {
    ZG_TRACE_BEGIN(3)
    pdl_enter();
    ZG_TRACE_MARK(1)
    __shared__ float sh[32];
    const uint32_t tid = threadIdx.x, r = blockIdx.x, c4 = m.cols >> 2;
    const size_t ro = (size_t)r * c4;
    const float4* a4 = reinterpret_cast<const float4*>(m.a);
    const float4* b4 = reinterpret_cast<const float4*>(m.b);
    const float4* g4 = reinterpret_cast<const float4*>(m.gamma);
    float4 x[2], g[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
        const uint32_t j = tid + u * 1024;
        x[u] = j < c4 ? a4[ro + j] : make_float4(0.f, 0.f, 0.f, 0.f);
        g[u] = j < c4 ? g4[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (b4) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const uint32_t j = tid + u * 1024;
            if (j >= c4) continue;
            const float4 t = b4[ro + j];
            x[u] = make_float4(x[u].x + t.x, x[u].y + t.y, x[u].z + t.z, x[u].w + t.w);
            reinterpret_cast<float4*>(m.sum)[ro + j] = x[u];
        }
    }
    float ss = 0.0f;
#pragma unroll
    for (int u = 0; u < 2; u++) ss += (x[u].x * x[u].x + x[u].y * x[u].y) + (x[u].z * x[u].z + x[u].w * x[u].w);
    ss = block_reduce<false>(ss, sh);
    const float inv_rms = 1.0f / sqrtf(ss / (float)m.cols + m.eps);
#pragma unroll
    for (int u = 0; u < 2; u++) {
        const uint32_t j = tid + u * 1024;
        if (j >= c4) continue;
        const float4 bz = make_float4(x[u].x * inv_rms, x[u].y * inv_rms, x[u].z * inv_rms, x[u].w * inv_rms);
        reinterpret_cast<float4*>(m.bare)[ro + j] = bz;
        reinterpret_cast<float4*>(m.gamma_rep)[ro + j] = g[u];
        reinterpret_cast<float4*>(m.norm)[ro + j] = make_float4(bz.x * g[u].x, bz.y * g[u].y, bz.z * g[u].z, bz.w * g[u].w);
    }
    ZG_TRACE_MARK(2)
}

// The same block for a long row spread over a thread-block CLUSTER of 8 CTAs x 256 threads (one float4 per thread): the single
// 1024-thread CTA above moves 96 KB in and 128 KB out through one SM's L2 port (~3 us of work on a launch that sits on the
// critical path twice per layer); here every CTA moves an eighth and the row's sum of squares crosses the cluster through
// distributed shared memory (each CTA publishes its partial, cluster barrier, everyone adds the 8 partials in rank order).
constexpr int kNormClusterCtas = 8;
__global__ void __launch_bounds__(256)
k_norm_macro_cluster(const ZgNormMacro m) {
    ZG_TRACE_BEGIN(3)
    pdl_enter();
    ZG_TRACE_MARK(1)
    __shared__ float sh[8];
    __shared__ float s_part;
    uint32_t n_cta;   // CTAs of this row's cluster (1, 2, 4 or 8: 256 float4 each)
    asm volatile("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(n_cta));
    const uint32_t tid = threadIdx.x, rank = blockIdx.x % n_cta, r = blockIdx.x / n_cta, c4 = m.cols >> 2;
    const size_t ro = (size_t)r * c4;
    const uint32_t j = rank * 256 + tid;
    const bool live = j < c4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f), g = x;
    if (live) {
        x = reinterpret_cast<const float4*>(m.a)[ro + j];
        g = reinterpret_cast<const float4*>(m.gamma)[j];
        if (m.b) {
            const float4 t = reinterpret_cast<const float4*>(m.b)[ro + j];
            x = make_float4(x.x + t.x, x.y + t.y, x.z + t.z, x.w + t.w);
            reinterpret_cast<float4*>(m.sum)[ro + j] = x;
        }
    }
    float ss = (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
    ss = warp_sum(ss);
    if ((tid & 31) == 0) sh[tid >> 5] = ss;
    __syncthreads();
    if (tid == 0) {
        float t = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) t += sh[i];
        s_part = t;
    }
    // cluster barrier (release / acquire): every CTA's partial is visible in its shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    float tot = 0.0f;
    {
        const uint32_t local = (uint32_t)__cvta_generic_to_shared(&s_part);
        for (uint32_t c = 0; c < n_cta; c++) {
            uint32_t remote; float v;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(c));
            asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
            tot += v;
        }
    }
    // nobody may exit (and release its shared memory) while a peer still reads it
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    const float inv_rms = 1.0f / sqrtf(tot / (float)m.cols + m.eps);
    if (live) {
        const float4 bz = make_float4(x.x * inv_rms, x.y * inv_rms, x.z * inv_rms, x.w * inv_rms);
        reinterpret_cast<float4*>(m.bare)[ro + j] = bz;
        reinterpret_cast<float4*>(m.gamma_rep)[ro + j] = g;
        reinterpret_cast<float4*>(m.norm)[ro + j] = make_float4(bz.x * g.x, bz.y * g.y, bz.z * g.z, bz.w * g.w);
    }
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    ZG_TRACE_MARK(2)
}

// fused_elementwise chain -> mid, then mid * other -> dst (SiLU(gate) * up): one launch for the two ops
__global__ void k_fused_ew_mul(const ZgDevStep* __restrict__ steps, uint32_t n_steps, float* __restrict__ mid,
                               const float* __restrict__ src, uint32_t n, const float* __restrict__ other, float* __restrict__ dst) {
    ZG_TRACE_BEGIN(2)
    pdl_enter();
    ZG_TRACE_MARK(1)
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float v = src[i];
        const float o2 = other[i];
        for (uint32_t s = 0; s < n_steps; s++) {
            const ZgDevStep st = steps[s];
            if (st.op == ZG_EW_ADD) { float o = st.sec[i]; v = st.is_swapped ? o + v : v + o; }
            else if (st.op == ZG_EW_MUL) { float o = st.sec[i]; v = st.is_swapped ? o * v : v * o; }
            else v = apply_unary(st.op, v);
        }
        mid[i] = v;
        dst[i] = v * o2;
    }
    ZG_TRACE_MARK(2)
}

// One-shot all-reduce over NVLink peer memory (ZG_OP_ALLREDUCE; state set up in comm.cu), low-latency protocol: every
// 16-byte store carries two floats AND the all-reduce's epoch twice ({x, epoch, y, epoch}), so data and "ready" flag
// arrive together — no fence, no separate flag round trip; the receiver polls the words of its own slot until both
// epochs match (a torn 16-byte transaction simply fails the check and is re-read).  kZgPeerCtas CTAs each own a
// contiguous slice: push it into every peer's slot [set][this rank], poll the peers' slices here, sum in RANK ORDER
// (bit-identical results on every rank), write back.  Slot sets alternate per all-reduce: a rank can only get two
// all-reduces ahead of a peer after that peer has finished reading the older set (it must have sent its data for
// the one in between, from a later kernel).  Every CTA keeps its own sequence counter; all ranks run the same
// all-reduces with the same grid, so the epochs agree.
__global__ void __launch_bounds__(256)
k_allreduce_peer(float* __restrict__ buf, uint32_t n2, const ZgPeerComm pc) {
    ZG_TRACE_BEGIN(11)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t tid = threadIdx.x, c = blockIdx.x;
    uint32_t* my_seq = pc.seq + 2 + c;
    const uint32_t chunk = (n2 + gridDim.x - 1) / gridDim.x, lo = c * chunk, hi = min(lo + chunk, n2);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    ZG_TRACE_MARK(1)
    // the counter is written by the previous all-reduce kernel, which may still be running while this one is already
    // resident (programmatic dependent launch chains several kernels deep): read it only after the wait
    const uint32_t seq = *(volatile uint32_t*)my_seq, set = seq % kZgPeerSets, epoch = seq + 1;
    float2* b2 = reinterpret_cast<float2*>(buf);
    const size_t slot_pairs = pc.max_n >> 1;                                  // 16-byte cells per (set, rank) slot
    const size_t my_cell = ((size_t)set * pc.world + pc.rank) * slot_pairs;
    for (uint32_t j = lo + tid; j < hi; j += 256) {
        const float2 v = b2[j];
        for (int pr = 0; pr < pc.world; pr++) {
            if (pr == pc.rank) continue;
            uint4* cell = reinterpret_cast<uint4*>(pc.slots[pr]) + my_cell + j;
            asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(cell), "r"(__float_as_uint(v.x)), "r"(epoch),
                         "r"(__float_as_uint(v.y)), "r"(epoch) : "memory");
        }
    }
    const uint4* mine = reinterpret_cast<const uint4*>(pc.slots[pc.rank]) + (size_t)set * pc.world * slot_pairs;
    for (uint32_t j = lo + tid; j < hi; j += 256) {
        uint4 got[kZgMaxRanks];
        uint32_t pending = 0;
#pragma unroll
        for (int r = 0; r < kZgMaxRanks; r++)
            if (r < pc.world && r != pc.rank) pending |= 1u << r;
        long long t0 = 0;   // per cell: a peer that is merely late for one cell does not shorten the others' wait
        uint32_t spins = 0;
        bool dead = false;
        while (pending) {   // every peer's cell is requested before any is checked; only the late ones are re-read
#pragma unroll
            for (int r = 0; r < kZgMaxRanks; r++)
                if (pending & (1u << r)) {
                    const uint4* cell = mine + (size_t)r * slot_pairs + j;
                    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(got[r].x), "=r"(got[r].y), "=r"(got[r].z), "=r"(got[r].w) : "l"(cell) : "memory");
                }
#pragma unroll
            for (int r = 0; r < kZgMaxRanks; r++)
                if ((pending & (1u << r)) && got[r].y == epoch && got[r].w == epoch) pending &= ~(1u << r);
            if (pending && (++spins & 1023u) == 0) {
                // A peer that never arrives (died, or its host stalled for longer than the limit) must neither hang the GPU
                // nor let stale cells into the sum: give up after ~30 s, raise the sticky error word (zg_cuda_execute /
                // zg_cuda_sync report it through zg_cuda_last_error) and poison this rank's result with NaN.
                if (t0 == 0) t0 = clock64();
                else if (clock64() - t0 > 60000000000LL || *(volatile uint32_t*)(pc.seq + 1)) { atomicExch(pc.seq + 1, epoch ? epoch : 1u); dead = true; break; }
            }
        }
        const float2 own = b2[j];
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kZgMaxRanks; r++)
            if (r < pc.world) {
                const float2 v = (r == pc.rank) ? own : make_float2(__uint_as_float(got[r].x), __uint_as_float(got[r].z));
                acc = (r == 0) ? v : make_float2(acc.x + v.x, acc.y + v.y);
            }
        if (dead) acc = make_float2(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000));
        b2[j] = acc;
    }
    __syncthreads();
    if (tid == 0) *(volatile uint32_t*)my_seq = epoch;
    ZG_TRACE_MARK(2)
}

// All-reduce over peer memory (k_allreduce_peer) + the block that consumes it in every sharded layer — [sum = a + b,]
// bare = rmsnorm(sum), gamma_rep = gamma, norm = bare * gamma_rep (ZgNormMacro, one row) — in ONE launch.  Each of the 16 CTAs
// reduces its slice exactly as k_allreduce_peer does, keeps going on the slice, and the sum of squares of the whole row is
// exchanged through 16 local {partial, epoch} cells (8-byte volatile stores, polled; added in CTA order -> identical on
// every CTA and every rank).  One kernel boundary less on the critical path of every all-reduce — and MEASURED SLOWER (round 2,
// same box A/B of the 80-layer 70B decode: 2 GPUs 131.0 -> 128.5 tok/s, 8 GPUs 191.8 -> 187.3): the separate norm kernel is
// resident and has its operands' addresses ready when the all-reduce ends (programmatic dependent launch), while here every
// CTA waits for the slowest CTA's peers and then for an L2 round trip of the partial exchange.  Off by default (ZG_CUDA_AR_NORM=1).
__global__ void __launch_bounds__(256)
k_allreduce_norm(float* __restrict__ buf, uint32_t n2, const ZgPeerComm pc, const ZgNormMacro m) {
    ZG_TRACE_BEGIN(11)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __shared__ float sh[32];
    __shared__ float s_tot;
    const uint32_t tid = threadIdx.x, c = blockIdx.x;
    uint32_t* my_seq = pc.seq + 2 + c;
    const uint32_t chunk = (n2 + gridDim.x - 1) / gridDim.x, lo = c * chunk, hi = min(lo + chunk, n2);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    ZG_TRACE_MARK(1)
    const uint32_t seq = *(volatile uint32_t*)my_seq, set = seq % kZgPeerSets, epoch = seq + 1;
    float2* b2 = reinterpret_cast<float2*>(buf);
    const size_t slot_pairs = pc.max_n >> 1;
    const size_t my_cell = ((size_t)set * pc.world + pc.rank) * slot_pairs;
    for (uint32_t j = lo + tid; j < hi; j += 256) {
        const float2 v = b2[j];
        for (int pr = 0; pr < pc.world; pr++) {
            if (pr == pc.rank) continue;
            uint4* cell = reinterpret_cast<uint4*>(pc.slots[pr]) + my_cell + j;
            asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(cell), "r"(__float_as_uint(v.x)), "r"(epoch),
                         "r"(__float_as_uint(v.y)), "r"(epoch) : "memory");
        }
    }
    const uint4* mine = reinterpret_cast<const uint4*>(pc.slots[pc.rank]) + (size_t)set * pc.world * slot_pairs;
    // the other operand of the residual add (nullptr: the macro has no add and normalises the all-reduced vector itself)
    const float2* other = nullptr;
    if (m.b) other = reinterpret_cast<const float2*>(m.a == buf ? m.b : m.a);
    float ss = 0.0f;
    for (uint32_t j = lo + tid; j < hi; j += 256) {
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
                    const uint4* cell = mine + (size_t)r * slot_pairs + j;
                    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(got[r].x), "=r"(got[r].y), "=r"(got[r].z), "=r"(got[r].w) : "l"(cell) : "memory");
                }
#pragma unroll
            for (int r = 0; r < kZgMaxRanks; r++)
                if ((pending & (1u << r)) && got[r].y == epoch && got[r].w == epoch) pending &= ~(1u << r);
            if (pending && (++spins & 1023u) == 0) {
                if (t0 == 0) t0 = clock64();
                else if (clock64() - t0 > 60000000000LL || *(volatile uint32_t*)(pc.seq + 1)) { atomicExch(pc.seq + 1, epoch ? epoch : 1u); dead = true; break; }
            }
        }
        const float2 own = b2[j];
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kZgMaxRanks; r++)
            if (r < pc.world) {
                const float2 v = (r == pc.rank) ? own : make_float2(__uint_as_float(got[r].x), __uint_as_float(got[r].z));
                acc = (r == 0) ? v : make_float2(acc.x + v.x, acc.y + v.y);
            }
        if (dead) acc = make_float2(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000));
        b2[j] = acc;
        float2 sv = acc;
        if (other) {
            const float2 o = other[j];
            sv = make_float2(acc.x + o.x, acc.y + o.y);
            reinterpret_cast<float2*>(m.sum)[j] = sv;
        }
        ss += sv.x * sv.x + sv.y * sv.y;
    }
    ss = block_reduce<false>(ss, sh);
    // exchange the CTA partials: cell = {partial, epoch}, one 8-byte volatile store; every CTA adds the 16 in CTA order
    if (tid == 0) {
        unsigned long long* cell = pc.cells + (size_t)set * kZgPeerCtas + c;
        asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(cell), "r"(__float_as_uint(ss)), "r"(epoch) : "memory");
    }
    if (tid < 32) {
        float part = 0.0f;
        if (tid < gridDim.x) {
            const unsigned long long* cell = pc.cells + (size_t)set * kZgPeerCtas + tid;
            uint32_t vx = 0, ve = 0, spins = 0;
            long long t0 = 0;
            for (;;) {
                asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(vx), "=r"(ve) : "l"(cell) : "memory");
                if (ve == epoch) break;
                if ((++spins & 1023u) == 0) {
                    if (t0 == 0) t0 = clock64();
                    else if (clock64() - t0 > 60000000000LL || *(volatile uint32_t*)(pc.seq + 1)) { atomicExch(pc.seq + 1, epoch ? epoch : 1u); vx = 0x7fc00000u; break; }
                }
            }
            part = __uint_as_float(vx);
        }
        float tot = 0.0f;
        for (uint32_t i = 0; i < gridDim.x; i++) tot += __shfl_sync(0xffffffffu, part, i);   // CTA order
        if (tid == 0) s_tot = tot;
    }
    __syncthreads();
    const float inv_rms = 1.0f / sqrtf(s_tot / (float)m.cols + m.eps);
    const float2* src = reinterpret_cast<const float2*>(other ? m.sum : buf);
    const float2* g2 = reinterpret_cast<const float2*>(m.gamma);
    for (uint32_t j = lo + tid; j < hi; j += 256) {
        const float2 sv = src[j], g = g2[j];      // this thread's own stores above
        const float2 bz = make_float2(sv.x * inv_rms, sv.y * inv_rms);
        reinterpret_cast<float2*>(m.bare)[j] = bz;
        reinterpret_cast<float2*>(m.gamma_rep)[j] = g;
        reinterpret_cast<float2*>(m.norm)[j] = make_float2(bz.x * g.x, bz.y * g.y);
    }
    __syncthreads();
    if (tid == 0) *(volatile uint32_t*)my_seq = epoch;
    ZG_TRACE_MARK(2)
}

// one block per row
__global__ void k_softmax(float* __restrict__ dst, const float* __restrict__ src, uint32_t cols) {
    pdl_enter();
    __shared__ float sh[32];
    const float* s = src + (size_t)blockIdx.x * cols;
    float* d = dst + (size_t)blockIdx.x * cols;
    float m = -INFINITY;
    for (uint32_t j = threadIdx.x; j < cols; j += blockDim.x) m = fmaxf(m, s[j]);
    m = block_reduce<true>(m, sh);
    float sum = 0.0f;
    for (uint32_t j = threadIdx.x; j < cols; j += blockDim.x) sum += expf(s[j] - m);
    sum = block_reduce<false>(sum, sh);
    float inv = sum > 0.0f ? 1.0f / sum : 0.0f;
    for (uint32_t j = threadIdx.x; j < cols; j += blockDim.x) d[j] = expf(s[j] - m) * inv;
}

__global__ void k_layernorm(float* __restrict__ dst, const float* __restrict__ src, uint32_t cols, float eps) {
    pdl_enter();
    __shared__ float sh[32];
    const float* s = src + (size_t)blockIdx.x * cols;
    float* d = dst + (size_t)blockIdx.x * cols;
    float mu = 0.0f;
    for (uint32_t j = threadIdx.x; j < cols; j += blockDim.x) mu += s[j];
    mu = block_reduce<false>(mu, sh) / (float)cols;
    float v = 0.0f;
    for (uint32_t j = threadIdx.x; j < cols; j += blockDim.x) { float df = s[j] - mu; v += df * df; }
    v = block_reduce<false>(v, sh);
    float inv_std = 1.0f / sqrtf(v / (float)cols + eps);
    for (uint32_t j = threadIdx.x; j < cols; j += blockDim.x) d[j] = (s[j] - mu) * inv_std;
}

__global__ void k_rmsnorm(float* __restrict__ dst, const float* __restrict__ src, uint32_t cols, float eps) {
    ZG_TRACE_BEGIN(3)
    pdl_enter();
    ZG_TRACE_MARK(1)
    __shared__ float sh[32];
    const float* s = src + (size_t)blockIdx.x * cols;
    float* d = dst + (size_t)blockIdx.x * cols;
    float ss = 0.0f;
    if ((cols & 3u) == 0 && (((size_t)s | (size_t)d) & 15u) == 0) {   // 128-bit loads, 4 in flight per thread (latency-bound otherwise)
        const float4* s4 = reinterpret_cast<const float4*>(s);
        float4* d4 = reinterpret_cast<float4*>(d);
        const uint32_t c4 = cols >> 2;
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll 4
        for (uint32_t j = threadIdx.x; j < c4; j += blockDim.x) { const float4 x = s4[j]; p0 += x.x * x.x; p1 += x.y * x.y; p2 += x.z * x.z; p3 += x.w * x.w; }
        ss = block_reduce<false>((p0 + p1) + (p2 + p3), sh);
        const float inv_rms = 1.0f / sqrtf(ss / (float)cols + eps);
#pragma unroll 4
        for (uint32_t j = threadIdx.x; j < c4; j += blockDim.x) { const float4 x = s4[j]; d4[j] = make_float4(x.x * inv_rms, x.y * inv_rms, x.z * inv_rms, x.w * inv_rms); }
        ZG_TRACE_MARK(2)
        return;
    }
    for (uint32_t j = threadIdx.x; j < cols; j += blockDim.x) { float x = s[j]; ss += x * x; }
    ss = block_reduce<false>(ss, sh);
    float inv_rms = 1.0f / sqrtf(ss / (float)cols + eps);
    for (uint32_t j = threadIdx.x; j < cols; j += blockDim.x) d[j] = s[j] * inv_rms;
}

// one warp per output
__global__ void k_reduce(uint32_t is_max, float* __restrict__ dst, const float* __restrict__ src,
                         uint32_t n_out, uint32_t rs) {
    pdl_enter();
    uint32_t o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31;
    if (o >= n_out) return;
    const float* s = src + (size_t)o * rs;
    float v = is_max ? -INFINITY : 0.0f;
    for (uint32_t k = lane; k < rs; k += 32) v = is_max ? fmaxf(v, s[k]) : v + s[k];
    v = is_max ? warp_max(v) : warp_sum(v);
    if (lane == 0) dst[o] = v;
}

struct RepeatParams {
    uint32_t mode; // 0 fill, 1 copy, 2 tile, 3 general
    uint32_t n, src_n;
    uint32_t src_ne[4], src_strides[4], dst_strides[4];
    uint32_t src_offset;
};
// reference.zig:391-433.  `src` is the buffer base for mode 3 (it adds src_offset
// itself), src+src_offset for the others; dst already includes dst_offset.
__global__ void k_repeat(RepeatParams p, float* __restrict__ dst, const float* __restrict__ src) {
    ZG_TRACE_BEGIN(4)
    pdl_enter();
    ZG_TRACE_MARK(1)
    for (uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x; gid < p.n; gid += gridDim.x * blockDim.x) {
        float v;
        if (p.mode == 0) v = src[0];
        else if (p.mode == 1) v = src[gid];
        else if (p.mode == 2) v = src[gid % p.src_n];
        else {
            uint32_t idx = gid, sidx = p.src_offset;
#pragma unroll
            for (int dim = 3; dim >= 0; dim--) {
                uint32_t coord = idx / p.dst_strides[dim];
                idx = idx % p.dst_strides[dim];
                sidx += (coord % p.src_ne[dim]) * p.src_strides[dim];
            }
            v = src[sidx];
        }
        dst[gid] = v;
    }
    ZG_TRACE_MARK(2)
}

// The three per-head op kinds of a decode program (rope, slice_assign, attention) are launched in BATCHES:
// blockIdx.y selects one op's parameter entry from a device table built at compile time, so all heads of a
// layer (mutually independent ops of one dependency level) cost one launch instead of one each.
__global__ void k_slice_assign(const ZgBatchEntry* __restrict__ tab, const uint32_t* __restrict__ d_dyn) {
    ZG_TRACE_BEGIN(5)
    pdl_enter();
    ZG_TRACE_MARK(1)
    const ZgBatchEntry e = tab[blockIdx.y];
    const uint32_t rows = e.u[0], cols = e.u[1], drs = e.u[2], dcs = e.u[3], soff = e.u[4], srs = e.u[5], scs = e.u[6];
    const uint32_t doff = d_dyn[e.dyn];
    float* __restrict__ dst = e.dst;
    const float* __restrict__ src = e.s0;
    uint32_t total = rows * cols;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        uint32_t row = i % rows, col = i / rows;
        dst[(size_t)doff + (size_t)row * drs + (size_t)col * dcs] = src[(size_t)soff + (size_t)row * srs + (size_t)col * scs];
    }
    ZG_TRACE_MARK(2)
}

__global__ void k_rope(const ZgBatchEntry* __restrict__ tab) {
    ZG_TRACE_BEGIN(6)
    pdl_enter();
    ZG_TRACE_MARK(1)
    const ZgBatchEntry e = tab[blockIdx.y];
    const uint32_t hd = e.u[0], seq_len = e.u[1], s_off = e.u[2], c_off = e.u[3], d_off = e.u[4], s_rs = e.u[5], s_cs = e.u[6], c_cs = e.u[7];
    float* __restrict__ dst = e.dst;
    const float* __restrict__ src = e.s0;
    const float* __restrict__ cs = e.s1;
    uint32_t total = hd * seq_len;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        uint32_t pair = i % hd, col = i / hd;
        float x_lo = src[(size_t)s_off + (size_t)pair * s_rs + (size_t)col * s_cs];
        float x_hi = src[(size_t)s_off + (size_t)(pair + hd) * s_rs + (size_t)col * s_cs];
        float c = cs[(size_t)c_off + pair + (size_t)col * c_cs];
        float sn = cs[(size_t)c_off + pair + hd + (size_t)col * c_cs];
        // separate roundings like the reference (no FMA contraction): reference.zig:474-475
        dst[(size_t)d_off + pair + (size_t)col * 2 * hd] = __fsub_rn(__fmul_rn(x_lo, c), __fmul_rn(x_hi, sn));
        dst[(size_t)d_off + pair + hd + (size_t)col * 2 * hd] = __fadd_rn(__fmul_rn(x_hi, c), __fmul_rn(x_lo, sn));
    }
    ZG_TRACE_MARK(2)
}

struct AttnParams {
    uint32_t has_mask, d_head, seq_q;
    float scale;
    uint32_t q_off, k_off, v_off, mask_off, dst_off;
    uint32_t q_rs, q_cs, k_rs, k_cs, v_rs, v_cs, mask_rs, mask_cs, dst_rs, dst_cs;
};
constexpr int kAttnWarps = 8;
constexpr int kAttnMaxPerLane = 16; // d_head <= 512

// One CTA per query row; each warp walks kv positions warp, warp+8, ... with an
// online softmax (reference.zig:599-671: non-finite mask / score entries skipped),
// the 8 partial states are merged through shared memory.
__global__ void __launch_bounds__(kAttnWarps * 32)
k_attention(const ZgBatchEntry* __restrict__ tab, const uint32_t* __restrict__ d_dyn) {
    pdl_enter();
    const ZgBatchEntry e = tab[blockIdx.y];
    AttnParams p;
    p.has_mask = e.u[0]; p.d_head = e.u[1]; p.seq_q = e.u[2]; p.scale = e.f;
    p.q_off = e.u[3]; p.k_off = e.u[4]; p.v_off = e.u[5]; p.mask_off = e.u[6]; p.dst_off = e.u[7];
    p.q_rs = e.u[8]; p.q_cs = e.u[9]; p.k_rs = e.u[10]; p.k_cs = e.u[11]; p.v_rs = e.u[12]; p.v_cs = e.u[13];
    p.mask_rs = e.u[14]; p.mask_cs = e.u[15]; p.dst_rs = e.u[16]; p.dst_cs = e.u[17];
    float* __restrict__ dst = e.dst;
    const float* __restrict__ q = e.s0;
    const float* __restrict__ k = e.s1;
    const float* __restrict__ v = e.s2;
    const float* __restrict__ mask = e.s3;
    const uint32_t seq_kv = d_dyn[e.dyn];
    const uint32_t qi = blockIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t dh = p.d_head;
    const size_t q_base = (size_t)p.q_off + (size_t)qi * p.q_cs;
    const size_t m_base = (size_t)p.mask_off + (size_t)qi * p.mask_cs;

    float qreg[kAttnMaxPerLane], acc[kAttnMaxPerLane];
#pragma unroll
    for (int j = 0; j < kAttnMaxPerLane; j++) {
        uint32_t r = lane + 32 * j;
        qreg[j] = (r < dh) ? q[q_base + (size_t)r * p.q_rs] : 0.0f;
        acc[j] = 0.0f;
    }
    float m_val = -INFINITY, l = 0.0f;
    for (uint32_t s = warp; s < seq_kv; s += kAttnWarps) {
        float mask_add = p.has_mask ? mask[m_base + (size_t)s * p.mask_rs] : 0.0f;
        if (!isfinite(mask_add)) continue; // warp-uniform
        const size_t kb = (size_t)p.k_off + (size_t)s * p.k_cs;
        float dot = 0.0f;
#pragma unroll
        for (int j = 0; j < kAttnMaxPerLane; j++) {
            uint32_t r = lane + 32 * j;
            if (r < dh) dot = fmaf(qreg[j], k[kb + (size_t)r * p.k_rs], dot);
        }
        dot = warp_sum(dot);
        float score = dot * p.scale + mask_add;
        if (!isfinite(score)) continue;
        float new_m = fmaxf(m_val, score);
        float alpha = (m_val == -INFINITY) ? 0.0f : expf(m_val - new_m);
        float w = expf(score - new_m);
        l = l * alpha + w;
        m_val = new_m;
        const size_t vb = (size_t)p.v_off + (size_t)s * p.v_cs;
#pragma unroll
        for (int j = 0; j < kAttnMaxPerLane; j++) {
            uint32_t r = lane + 32 * j;
            if (r < dh) acc[j] = acc[j] * alpha + w * v[vb + (size_t)r * p.v_rs];
        }
    }
    __shared__ float sh_m[kAttnWarps], sh_l[kAttnWarps];
    __shared__ float sh_acc[kAttnWarps][512];
    if (lane == 0) { sh_m[warp] = m_val; sh_l[warp] = l; }
#pragma unroll
    for (int j = 0; j < kAttnMaxPerLane; j++) {
        uint32_t r = lane + 32 * j;
        if (r < dh) sh_acc[warp][r] = acc[j];
    }
    __syncthreads();
    float gm = -INFINITY;
    for (int w = 0; w < kAttnWarps; w++) gm = fmaxf(gm, sh_m[w]);
    float gl = 0.0f;
    float wscale[kAttnWarps];
    for (int w = 0; w < kAttnWarps; w++) {
        wscale[w] = (sh_m[w] == -INFINITY) ? 0.0f : expf(sh_m[w] - gm);
        gl += sh_l[w] * wscale[w];
    }
    float inv_l = gl > 0.0f ? 1.0f / gl : 0.0f;
    const size_t d_base = (size_t)p.dst_off + (size_t)qi * p.dst_cs;
    for (uint32_t r = threadIdx.x; r < dh; r += blockDim.x) {
        float a = 0.0f;
        for (int w = 0; w < kAttnWarps; w++) a += sh_acc[w][r] * wscale[w];
        dst[d_base + (size_t)r * p.dst_rs] = a * inv_l;
        if (e.dst2) e.dst2[(size_t)e.d2_off + (size_t)r * e.d2_rs + (size_t)qi * e.d2_cs] = a * inv_l;
    }
}

// Decode / prefill attention fast path (unit row strides, d_head % 4 == 0, d_head <= 256): within a warp lane =
// kv position for the scores (each lane dots its own contiguous K row against q in shared memory, 128-bit loads),
// then lane = head dimension for the V accumulation (coalesced rows, weights broadcast by shuffle).  32 positions
// per warp step, 16 warps per CTA, and every global load of a step is issued before its first use (K: 4 float4 per
// lane in flight, V: 8 rows in flight), so a 512-token context is ONE pass of independent loads instead of a chain
// of dependent ones.  Same online softmax and skip rules as reference.zig:599-671; only the summation order
// differs (1e-6 relative).  NI = ceil(d_head / 32) values per lane.
constexpr int kAttnFastWarps = 16;
template <int NI>
__global__ void __launch_bounds__(kAttnFastWarps * 32)
k_attention_fast(const ZgBatchEntry* __restrict__ tab, const uint32_t* __restrict__ d_dyn, float* __restrict__ part,
                 uint32_t* __restrict__ cnt, const uint32_t max_splits) {
    ZG_TRACE_BEGIN(7)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const ZgBatchEntry e = tab[blockIdx.y];
    // the launch provides max_splits CTAs per head; a short context uses fewer (>= 128 positions each), the rest leave
    const uint32_t splits = min(max_splits, max(1u, (d_dyn[e.dyn] + 127u) / 128u));
    if (blockIdx.z >= splits) return;
    {   // While the producer kernels still run: pull this head's K and V rows into L2 (one 128-byte line per prefetch).
        // The op table and seq_kv are fixed before the step starts; the cache rows are only read after the wait (the
        // row written by this step just gets prefetched a little early — L2 is the point of coherence).
        const uint32_t seq_kv0 = d_dyn[e.dyn], dhb = e.u[1] * 4;
        const char* kb = reinterpret_cast<const char*>(e.s1 + e.u[4]);
        const char* vb0 = reinterpret_cast<const char*>(e.s2 + e.u[5]);
        const uint32_t ch0 = ((seq_kv0 + splits - 1) / splits + 31) & ~31u;
        const uint32_t p_lo = blockIdx.z * ch0, p_hi = min(p_lo + ch0, seq_kv0);
        if (blockIdx.x == 0)
            for (uint32_t s = p_lo + threadIdx.x; s < p_hi; s += blockDim.x)
                for (uint32_t b = 0; b < dhb; b += 128) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(kb + (size_t)s * e.u[11] * 4 + b));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(vb0 + (size_t)s * e.u[13] * 4 + b));
                }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    ZG_TRACE_MARK(1)
    const uint32_t has_mask = e.u[0], dh = e.u[1];
    const float scale = e.f;
    const uint32_t q_off = e.u[3], k_off = e.u[4], v_off = e.u[5], mask_off = e.u[6], dst_off = e.u[7];
    const uint32_t q_cs = e.u[9], k_cs = e.u[11], v_cs = e.u[13], mask_rs = e.u[14], mask_cs = e.u[15], dst_cs = e.u[17];
    float* __restrict__ dst = e.dst;
    const float* __restrict__ q = e.s0;
    const float* __restrict__ k = e.s1;
    const float* __restrict__ v = e.s2;
    const float* __restrict__ mask = e.s3;
    const uint32_t seq_kv = d_dyn[e.dyn];
    const uint32_t qi = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    __shared__ __align__(16) float sq[NI * 32];
    __shared__ float sh_m[kAttnFastWarps], sh_l[kAttnFastWarps];
    __shared__ float sh_acc[kAttnFastWarps][NI * 32];
    for (uint32_t r = threadIdx.x; r < NI * 32; r += blockDim.x) sq[r] = r < dh ? q[(size_t)q_off + (size_t)qi * q_cs + r] : 0.0f;
    __syncthreads();
    const size_t m_base = (size_t)mask_off + (size_t)qi * mask_cs;
    const uint32_t dh4 = dh >> 2;
    float acc[NI];
#pragma unroll
    for (int i = 0; i < NI; i++) acc[i] = 0.0f;
    float m_val = -INFINITY, l = 0.0f;
    // split-KV: this CTA covers positions [kv_lo, kv_hi) (32-aligned chunks); one head's cache is then streamed by
    // `splits` SMs instead of one (a single SM's L2 bandwidth, not HBM, bounds a 512 KB head otherwise)
    const uint32_t chunk = ((seq_kv + splits - 1) / splits + 31) & ~31u;
    const uint32_t kv_lo = blockIdx.z * chunk, kv_hi = min(kv_lo + chunk, seq_kv);
    // Lanes per position: a short range is spread over all 16 warps (PW = 32 / LPP positions per warp step) and each
    // K row is read by LPP lanes (dh / LPP contiguous floats each, then a shuffle reduction), so a step is one round of
    // independent loads for K and one for V instead of four each when only a few warps had work.
    uint32_t lpp = 1;
    {
        const uint32_t len = kv_hi > kv_lo ? kv_hi - kv_lo : 0;
        while (lpp < 8 && len <= (kAttnFastWarps * 32u) / (2 * lpp) && (dh4 % (2 * lpp)) == 0) lpp *= 2;
    }
    const uint32_t pw = 32 / lpp, seg = lane & (lpp - 1), pos_in_warp = lane / lpp;
    const uint32_t f4 = dh4 / lpp;   // float4 per lane of a K row
    for (uint32_t s0 = kv_lo + warp * pw; s0 < kv_hi; s0 += kAttnFastWarps * pw) {
        const uint32_t s = s0 + pos_in_warp;
        float mask_add = -INFINITY;
        if (s < kv_hi) mask_add = has_mask ? mask[m_base + (size_t)s * mask_rs] : 0.0f;
        bool ok = isfinite(mask_add);
        float score = -INFINITY;
        {
            float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f, d3 = 0.0f;
            if (ok) {
                const float4* kr = reinterpret_cast<const float4*>(k + (size_t)k_off + (size_t)s * k_cs) + seg * f4;
                const float4* q4 = reinterpret_cast<const float4*>(sq) + seg * f4;
                uint32_t d = 0;
                for (; d + 8 <= f4; d += 8) {   // 8 independent 128-bit loads in flight per lane
                    float4 kk[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) kk[u] = kr[d + u];
#pragma unroll
                    for (int u = 0; u < 8; u += 4) {
                        const float4 qa = q4[d + u], qb = q4[d + u + 1], qc = q4[d + u + 2], qf = q4[d + u + 3];
                        d0 = fmaf(qa.x, kk[u].x, d0); d0 = fmaf(qa.y, kk[u].y, d0); d0 = fmaf(qa.z, kk[u].z, d0); d0 = fmaf(qa.w, kk[u].w, d0);
                        d1 = fmaf(qb.x, kk[u + 1].x, d1); d1 = fmaf(qb.y, kk[u + 1].y, d1); d1 = fmaf(qb.z, kk[u + 1].z, d1); d1 = fmaf(qb.w, kk[u + 1].w, d1);
                        d2 = fmaf(qc.x, kk[u + 2].x, d2); d2 = fmaf(qc.y, kk[u + 2].y, d2); d2 = fmaf(qc.z, kk[u + 2].z, d2); d2 = fmaf(qc.w, kk[u + 2].w, d2);
                        d3 = fmaf(qf.x, kk[u + 3].x, d3); d3 = fmaf(qf.y, kk[u + 3].y, d3); d3 = fmaf(qf.z, kk[u + 3].z, d3); d3 = fmaf(qf.w, kk[u + 3].w, d3);
                    }
                }
                if (d + 4 <= f4) {
                    float4 kk[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) kk[u] = kr[d + u];
                    const float4 qa = q4[d], qb = q4[d + 1], qc = q4[d + 2], qf = q4[d + 3];
                    d0 = fmaf(qa.x, kk[0].x, d0); d0 = fmaf(qa.y, kk[0].y, d0); d0 = fmaf(qa.z, kk[0].z, d0); d0 = fmaf(qa.w, kk[0].w, d0);
                    d1 = fmaf(qb.x, kk[1].x, d1); d1 = fmaf(qb.y, kk[1].y, d1); d1 = fmaf(qb.z, kk[1].z, d1); d1 = fmaf(qb.w, kk[1].w, d1);
                    d2 = fmaf(qc.x, kk[2].x, d2); d2 = fmaf(qc.y, kk[2].y, d2); d2 = fmaf(qc.z, kk[2].z, d2); d2 = fmaf(qc.w, kk[2].w, d2);
                    d3 = fmaf(qf.x, kk[3].x, d3); d3 = fmaf(qf.y, kk[3].y, d3); d3 = fmaf(qf.z, kk[3].z, d3); d3 = fmaf(qf.w, kk[3].w, d3);
                    d += 4;
                }
                for (; d < f4; d++) {
                    const float4 a = kr[d], qa = q4[d];
                    d0 = fmaf(qa.x, a.x, d0); d0 = fmaf(qa.y, a.y, d0); d0 = fmaf(qa.z, a.z, d0); d0 = fmaf(qa.w, a.w, d0);
                }
            }
            float dot = (d0 + d1) + (d2 + d3);
            for (uint32_t o = 1; o < lpp; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);   // the lanes of a position share ok
            if (ok) {
                score = dot * scale + mask_add;
                ok = isfinite(score);
                if (!ok) score = -INFINITY;
            }
        }
        float bm = score;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
        if (bm == -INFINITY) continue;   // warp-uniform: nothing attendable in this block
        const float new_m = fmaxf(m_val, bm);
        const float alpha = (m_val == -INFINITY) ? 0.0f : expf(m_val - new_m);
        const float wgt = ok ? expf(score - new_m) : 0.0f;   // lanes past the range and skipped entries weigh 0
        l = l * alpha + warp_sum(seg == 0 ? wgt : 0.0f);     // one lane per position counts
        m_val = new_m;
#pragma unroll
        for (int i = 0; i < NI; i++) acc[i] *= alpha;
        const uint32_t nj = min(pw, kv_hi - s0);
        const float* vb = v + (size_t)v_off + (size_t)s0 * v_cs + lane;
#pragma unroll 1
        for (uint32_t j0 = 0; j0 < nj; j0 += 8) {
            float vv[8][NI];
#pragma unroll
            for (int jj = 0; jj < 8; jj++) {
                const bool in = j0 + jj < nj;
#pragma unroll
                for (int i = 0; i < NI; i++) vv[jj][i] = (in && lane + 32 * i < dh) ? vb[(size_t)(j0 + jj) * v_cs + 32 * i] : 0.0f;
            }
#pragma unroll
            for (int jj = 0; jj < 8; jj++) {
                const float wj = __shfl_sync(0xffffffffu, wgt, ((j0 + jj) * lpp) & 31);
#pragma unroll
                for (int i = 0; i < NI; i++) acc[i] = fmaf(wj, vv[jj][i], acc[i]);
            }
        }
    }
    if (lane == 0) { sh_m[warp] = m_val; sh_l[warp] = l; }
#pragma unroll
    for (int i = 0; i < NI; i++) sh_acc[warp][lane + 32 * i] = acc[i];
    __syncthreads();
    float gm = -INFINITY;
    for (int w = 0; w < kAttnFastWarps; w++) gm = fmaxf(gm, sh_m[w]);
    float gl = 0.0f;
    float wscale[kAttnFastWarps];
#pragma unroll
    for (int w = 0; w < kAttnFastWarps; w++) {
        wscale[w] = (sh_m[w] == -INFINITY) ? 0.0f : expf(sh_m[w] - gm);
        gl += sh_l[w] * wscale[w];
    }
    if (splits > 1) {
        // partial state (max, sum, unnormalised accumulator) of this split -> scratch; the last split to arrive merges
        // all of them in split order (deterministic) and writes the output
        __shared__ uint32_t s_last;
        const uint32_t row = blockIdx.y * gridDim.x + qi;
        float* mine = part + ((size_t)row * max_splits + blockIdx.z) * (NI * 32 + 2);
        for (uint32_t r = threadIdx.x; r < dh; r += blockDim.x) {
            float a = 0.0f;
#pragma unroll
            for (int w = 0; w < kAttnFastWarps; w++) a += sh_acc[w][r] * wscale[w];
            mine[2 + r] = a;
        }
        if (threadIdx.x == 0) { mine[0] = gm; mine[1] = gl; }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t old;
            asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(cnt + row) : "memory");
            const uint32_t last = (old == splits - 1) ? 1u : 0u;
            if (last) cnt[row] = 0u;   // re-arm for the next launch
            s_last = last;
        }
        __syncthreads();
        if (s_last) {
            const float* all = part + (size_t)row * max_splits * (NI * 32 + 2);
            float G = -INFINITY;
            for (uint32_t sp = 0; sp < splits; sp++) G = fmaxf(G, __ldcg(all + (size_t)sp * (NI * 32 + 2)));
            float L = 0.0f;
            for (uint32_t sp = 0; sp < splits; sp++) {
                const float ms = __ldcg(all + (size_t)sp * (NI * 32 + 2));
                L += (ms == -INFINITY) ? 0.0f : __ldcg(all + (size_t)sp * (NI * 32 + 2) + 1) * expf(ms - G);
            }
            const float inv_L = L > 0.0f ? 1.0f / L : 0.0f;
            for (uint32_t r = threadIdx.x; r < dh; r += blockDim.x) {
                float a = 0.0f;
                for (uint32_t sp = 0; sp < splits; sp++) {
                    const float ms = __ldcg(all + (size_t)sp * (NI * 32 + 2));
                    if (ms != -INFINITY) a += __ldcg(all + (size_t)sp * (NI * 32 + 2) + 2 + r) * expf(ms - G);
                }
                dst[(size_t)dst_off + (size_t)qi * dst_cs + r] = a * inv_L;
                if (e.dst2) e.dst2[(size_t)e.d2_off + (size_t)r * e.d2_rs + (size_t)qi * e.d2_cs] = a * inv_L;
            }
        }
        ZG_TRACE_MARK(2)
        return;
    }
    const float inv_l = gl > 0.0f ? 1.0f / gl : 0.0f;
    for (uint32_t r = threadIdx.x; r < dh; r += blockDim.x) {
        float a = 0.0f;
#pragma unroll
        for (int w = 0; w < kAttnFastWarps; w++) a += sh_acc[w][r] * wscale[w];
        dst[(size_t)dst_off + (size_t)qi * dst_cs + r] = a * inv_l;
        if (e.dst2) e.dst2[(size_t)e.d2_off + (size_t)r * e.d2_rs + (size_t)qi * e.d2_cs] = a * inv_l;
    }
    ZG_TRACE_MARK(2)
}

// Prefill attention (seq_q >= 16, d_head 64 / 128, unit row strides): flash-attention tiling in fp32.  One CTA owns 64
// query rows of one head and walks the kv positions in tiles of 64: Q, K, V tiles in shared memory, the 64 x 64 score tile
// in registers (thread (ty, tx) of 16 x 16 holds rows 4ty.., columns 4tx..), online softmax per row (row statistics shared by
// the 16 threads of a row through half-warp shuffles), probabilities through shared memory, output accumulators in
// registers (rows 4ty.., DH / 16 columns per thread).  Same skip rules as reference.zig:599-671 (non-finite mask or score
// entries are skipped, an all-masked row gives zeros); a kv tile whose mask is non-finite for every (row, position) pair of
// the CTA — everything above the causal diagonal — is skipped without touching K or V.  The one-CTA-per-query-row kernel
// this replaces re-read every K / V row per query: 466 ms of a 523 ms Llama-3-8B prefill of 2048 tokens.
constexpr int kPfBQ = 64, kPfBK = 64, kPfThreads = 256;
template <int DH>
__global__ void __launch_bounds__(kPfThreads, 2)
k_attention_prefill(const ZgBatchEntry* __restrict__ tab, const uint32_t* __restrict__ d_dyn) {
    ZG_TRACE_BEGIN(7)
    pdl_enter();
    ZG_TRACE_MARK(1)
    extern __shared__ __align__(16) float pf_smem[];
    float* Qs = pf_smem;                         // [64][DH]
    float* Ks = Qs + kPfBQ * DH;                 // [64][DH + 4]  (+4: conflict-free column reads of four consecutive rows)
    float* Vs = Ks + kPfBK * (DH + 4);           // [64][DH]
    float* Ps = Ks;                              // [64][64 + 4]: the probabilities reuse the K tile (two CTAs per SM fit this way)
    constexpr int KP = DH + 4, PP = kPfBK + 4, OC = DH / 16;
    const ZgBatchEntry e = tab[blockIdx.y];
    const uint32_t has_mask = e.u[0], seq_q = e.u[2];
    const float scale = e.f;
    const uint32_t q_off = e.u[3], k_off = e.u[4], v_off = e.u[5], mask_off = e.u[6], dst_off = e.u[7];
    const uint32_t q_cs = e.u[9], k_cs = e.u[11], v_cs = e.u[13], mask_rs = e.u[14], mask_cs = e.u[15], dst_cs = e.u[17];
    const float* __restrict__ q = e.s0;
    const float* __restrict__ k = e.s1;
    const float* __restrict__ v = e.s2;
    const float* __restrict__ mask = e.s3;
    const uint32_t seq_kv = d_dyn[e.dyn];
    const uint32_t tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const uint32_t q0 = blockIdx.x * kPfBQ;
    // Q tile (rows past seq_q: zeros)
    for (uint32_t i = tid; i < kPfBQ * (DH / 4); i += kPfThreads) {
        const uint32_t r = i / (DH / 4), c4 = i % (DH / 4);
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + r < seq_q) val = *reinterpret_cast<const float4*>(q + (size_t)q_off + (size_t)(q0 + r) * q_cs + 4 * c4);
        *reinterpret_cast<float4*>(Qs + r * DH + 4 * c4) = val;
    }
    float o[4][OC];
    float m_run[4], l_run[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        m_run[r] = -INFINITY; l_run[r] = 0.0f;
#pragma unroll
        for (int c = 0; c < OC; c++) o[r][c] = 0.0f;
    }
    for (uint32_t s0 = 0; s0 < seq_kv; s0 += kPfBK) {
        // additive mask of this thread's 4 x 4 entries; out-of-range rows / positions count as masked
        float mk[4][4];
        bool any = false;
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const uint32_t qi = q0 + 4 * ty + r, sj = s0 + 4 * tx + c;
                float mv = -INFINITY;
                if (qi < seq_q && sj < seq_kv) mv = has_mask ? mask[(size_t)mask_off + (size_t)sj * mask_rs + (size_t)qi * mask_cs] : 0.0f;
                mk[r][c] = mv;
                any = any || isfinite(mv);
            }
        if (!__syncthreads_or(any ? 1 : 0)) continue;   // nothing attendable in the tile (also orders the previous tile's reads of Ks / Vs / Ps)
        for (uint32_t i = tid; i < kPfBK * (DH / 4); i += kPfThreads) {
            const uint32_t r = i / (DH / 4), c4 = i % (DH / 4);
            float4 kv4 = make_float4(0.f, 0.f, 0.f, 0.f), vv4 = kv4;
            if (s0 + r < seq_kv) {
                kv4 = *reinterpret_cast<const float4*>(k + (size_t)k_off + (size_t)(s0 + r) * k_cs + 4 * c4);
                vv4 = *reinterpret_cast<const float4*>(v + (size_t)v_off + (size_t)(s0 + r) * v_cs + 4 * c4);
            }
            *reinterpret_cast<float4*>(Ks + r * KP + 4 * c4) = kv4;
            *reinterpret_cast<float4*>(Vs + r * DH + 4 * c4) = vv4;
        }
        __syncthreads();
        // scores: S[4ty + r][4tx + c] = q . k
        float sc[4][4];
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) sc[r][c] = 0.0f;
#pragma unroll 4
        for (int d = 0; d < DH; d += 4) {
            float4 qa[4], ka[4];
#pragma unroll
            for (int r = 0; r < 4; r++) qa[r] = *reinterpret_cast<const float4*>(Qs + (4 * ty + r) * DH + d);
#pragma unroll
            for (int c = 0; c < 4; c++) ka[c] = *reinterpret_cast<const float4*>(Ks + (4 * tx + c) * KP + d);
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    sc[r][c] = fmaf(qa[r].x, ka[c].x, sc[r][c]); sc[r][c] = fmaf(qa[r].y, ka[c].y, sc[r][c]);
                    sc[r][c] = fmaf(qa[r].z, ka[c].z, sc[r][c]); sc[r][c] = fmaf(qa[r].w, ka[c].w, sc[r][c]);
                }
        }
        __syncthreads();   // every thread is done with the K tile: its memory now takes the probabilities
        // online softmax per row (16 threads of a half warp share a row)
#pragma unroll
        for (int r = 0; r < 4; r++) {
            float rmax = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                float sv = -INFINITY;
                if (isfinite(mk[r][c])) { sv = sc[r][c] * scale + mk[r][c]; if (!isfinite(sv)) sv = -INFINITY; }
                sc[r][c] = sv;
                rmax = fmaxf(rmax, sv);
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, off));
            const float new_m = fmaxf(m_run[r], rmax);
            const float alpha = (m_run[r] == -INFINITY || new_m == -INFINITY) ? (new_m == -INFINITY ? 1.0f : 0.0f) : expf(m_run[r] - new_m);
            float rsum = 0.0f;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float pv = (sc[r][c] == -INFINITY) ? 0.0f : expf(sc[r][c] - new_m);
                Ps[(4 * ty + r) * PP + 4 * tx + c] = pv;
                rsum += pv;
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) rsum += __shfl_xor_sync(0xffffffffu, rsum, off);
            l_run[r] = l_run[r] * alpha + rsum;
            m_run[r] = new_m;
#pragma unroll
            for (int c = 0; c < OC; c++) o[r][c] *= alpha;
        }
        __syncthreads();
        // O[4ty + r][tx + 16c] += sum_j P[4ty + r][j] V[j][tx + 16c]   (column = tx + 16c: conflict-free V reads)
#pragma unroll 4
        for (int j = 0; j < kPfBK; j++) {
            float pj[4], vj[OC];
#pragma unroll
            for (int r = 0; r < 4; r++) pj[r] = Ps[(4 * ty + r) * PP + j];
#pragma unroll
            for (int c = 0; c < OC; c++) vj[c] = Vs[j * DH + tx + 16 * c];
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < OC; c++) o[r][c] = fmaf(pj[r], vj[c], o[r][c]);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const uint32_t qi = q0 + 4 * ty + r;
        if (qi >= seq_q) continue;
        const float inv_l = l_run[r] > 0.0f ? 1.0f / l_run[r] : 0.0f;
#pragma unroll
        for (int c = 0; c < OC; c++) {
            const uint32_t col = tx + 16 * c;
            const float val = o[r][c] * inv_l;
            e.dst[(size_t)dst_off + (size_t)qi * dst_cs + col] = val;
            if (e.dst2) e.dst2[(size_t)e.d2_off + (size_t)col * e.d2_rs + (size_t)qi * e.d2_cs] = val;
        }
    }
    ZG_TRACE_MARK(2)
}
template <int DH> constexpr size_t pf_smem_bytes() { return (size_t)(kPfBQ * DH + kPfBK * (DH + 4) + kPfBK * DH) * sizeof(float); }
static_assert(kPfBQ * (kPfBK + 4) <= kPfBK * (64 + 4), "the probability tile fits inside the K tile");
static inline bool attn_prefill_ok(const ZgOp& op);

// ── the attention block of one layer of a single-token program in ONE launch (ZgAttnBlock) ──
// CTA = (query head, kv split).  Prologue: rope of this head's query and of its KV head's key from the projections (the
// rotated query / key buffers, the K-cache row and the V-cache row are stored by the split-0 CTA of the head / of the KV
// head's first query head).  The position written by this step is served from shared memory, so no CTA reads a cache row
// another CTA writes in the same launch.  Main loop, split merge and the second store into the concatenated buffer as in
// k_attention_fast.  Replaces rope x (n_kv + n_heads), slice_assign x (2 n_kv + n_heads) and the attention launch.
template <int NI>
__global__ void __launch_bounds__(kAttnFastWarps * 32)
k_attention_layer(const ZgAttnBlock* __restrict__ blkp, const uint32_t* __restrict__ d_dyn, float* __restrict__ part, uint32_t* __restrict__ cnt,
                  const uint32_t max_splits, const uint32_t cl) {
    ZG_TRACE_BEGIN(7)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const ZgAttnBlock& B = *blkp;
    const uint32_t h = blockIdx.x, sp = blockIdx.y;
    const ZgDecHead hd = B.heads[h];
    const ZgDecKv kv = B.kvs[hd.kv];
    const uint32_t dh = B.d_head, hd2 = dh >> 1;
    const uint32_t seq_kv = d_dyn[hd.dyn];
    const uint32_t splits = min(max_splits, max(1u, seq_kv / B.min_pos));   // every split at least min_pos positions: a short tail split only adds a merge
    if (sp >= splits) {
        // cl: the max_splits CTAs of a head are one cluster (rank = split) and merge through distributed shared memory; the
        // splits this step does not use still take part in the two cluster barriers of the merge
        if (cl) {
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
        return;
    }
    const uint32_t chunk = ((seq_kv + splits - 1) / splits + 31) & ~31u;
    const uint32_t kv_lo = sp * chunk, kv_hi = min(kv_lo + chunk, seq_kv);
    const float* kbase = B.k_cache + hd.k_off;
    const float* vbase = B.v_cache + hd.v_off;
    {   // while the projections are still being computed: pull this CTA's cache rows into L2 (rows below the one written by
        // this step are final since the previous step)
        const char* kb = reinterpret_cast<const char*>(kbase);
        const char* vb = reinterpret_cast<const char*>(vbase);
        for (uint32_t s = kv_lo + threadIdx.x; s < kv_hi; s += blockDim.x)
            for (uint32_t b = 0; b < dh * 4; b += 128) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(kb + (size_t)s * B.k_cs * 4 + b));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(vb + (size_t)s * B.v_cs * 4 + b));
            }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    ZG_TRACE_MARK(1)
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ __align__(16) float sq[NI * 32];
    __shared__ __align__(16) float sk[NI * 32];
    __shared__ __align__(16) float sv[NI * 32];
    __shared__ float sh_m[kAttnFastWarps], sh_l[kAttnFastWarps];
    __shared__ float sh_acc[kAttnFastWarps][NI * 32];
    const uint32_t k_dst = d_dyn[kv.k_dyn], v_dst = d_dyn[kv.v_dyn];
    // the kv position this step writes, as this head's attention op numbers its rows (UINT32_MAX: not one of its rows)
    uint32_t s_new = UINT32_MAX;
    if (k_dst >= hd.k_off && (k_dst - hd.k_off) % B.k_cs == 0 && v_dst >= hd.v_off && (v_dst - hd.v_off) % B.v_cs == 0 &&
        (k_dst - hd.k_off) / B.k_cs == (v_dst - hd.v_off) / B.v_cs)
        s_new = (k_dst - hd.k_off) / B.k_cs;
    const bool store = sp == 0 && ((h == 0) || (B.heads[h - 1].kv != hd.kv));
    if (tid < NI * 32) {
        const uint32_t r = tid;
        float qv = 0.0f, kvv = 0.0f, vv = 0.0f;
        if (r < dh) {
            const bool lo = r < hd2;
            const uint32_t pair = lo ? r : r - hd2, other = lo ? r + hd2 : pair;
            const float c = B.cs[pair], sn = B.cs[pair + hd2];
            const float q_me = B.q_proj[hd.q_src + r], q_pt = B.q_proj[hd.q_src + other];
            const float k_me = B.k_proj[kv.k_src + r], k_pt = B.k_proj[kv.k_src + other];
            vv = B.v_proj[kv.v_src + r];
            // separate roundings like the reference (reference.zig:474-475)
            qv = lo ? __fsub_rn(__fmul_rn(q_me, c), __fmul_rn(q_pt, sn)) : __fadd_rn(__fmul_rn(q_me, c), __fmul_rn(q_pt, sn));
            kvv = lo ? __fsub_rn(__fmul_rn(k_me, c), __fmul_rn(k_pt, sn)) : __fadd_rn(__fmul_rn(k_me, c), __fmul_rn(k_pt, sn));
            if (sp == 0) hd.q_rot[r] = qv;
            if (store) {
                kv.k_rot[r] = kvv;
                B.k_cache[(size_t)k_dst + r] = kvv;
                B.v_cache[(size_t)v_dst + r] = vv;
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
    uint32_t lpp = 1;
    {
        const uint32_t len = kv_hi > kv_lo ? kv_hi - kv_lo : 0;
        while (lpp < 8 && len <= (kAttnFastWarps * 32u) / (2 * lpp) && (dh4 % (2 * lpp)) == 0) lpp *= 2;
    }
    const uint32_t pw = 32 / lpp, seg = lane & (lpp - 1), pos_in_warp = lane / lpp;
    const uint32_t f4 = dh4 / lpp;
    for (uint32_t s0 = kv_lo + warp * pw; s0 < kv_hi; s0 += kAttnFastWarps * pw) {
        const uint32_t s = s0 + pos_in_warp;
        const uint32_t nj = min(pw, kv_hi - s0);
        // every global load of the step is issued before the first one is used: mask, the first eight V rows (all of them when
        // the range is spread over the 16 warps), then K — one memory round trip per step instead of two
        float mask_add = -INFINITY;
        if (s < kv_hi) mask_add = B.has_mask ? B.mask[(size_t)B.mask_off + (size_t)s * B.mask_rs] : 0.0f;
        float vv[8][NI];
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
            const uint32_t sj = s0 + jj;
            const float* vr = (sj == s_new) ? sv : vbase + (size_t)sj * B.v_cs;
#pragma unroll
            for (int i = 0; i < NI; i++) vv[jj][i] = ((uint32_t)jj < nj && lane + 32 * i < dh) ? vr[lane + 32 * i] : 0.0f;
        }
        bool ok = isfinite(mask_add);
        float score = -INFINITY;
        {
            float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f, d3 = 0.0f;
            if (ok) {
                const float4* kr = reinterpret_cast<const float4*>(s == s_new ? sk : kbase + (size_t)s * B.k_cs) + seg * f4;
                const float4* q4 = reinterpret_cast<const float4*>(sq) + seg * f4;
                uint32_t d = 0;
                for (; d + 8 <= f4; d += 8) {
                    float4 kk[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) kk[u] = kr[d + u];
#pragma unroll
                    for (int u = 0; u < 8; u += 4) {
                        const float4 qa = q4[d + u], qb = q4[d + u + 1], qc = q4[d + u + 2], qf = q4[d + u + 3];
                        d0 = fmaf(qa.x, kk[u].x, d0); d0 = fmaf(qa.y, kk[u].y, d0); d0 = fmaf(qa.z, kk[u].z, d0); d0 = fmaf(qa.w, kk[u].w, d0);
                        d1 = fmaf(qb.x, kk[u + 1].x, d1); d1 = fmaf(qb.y, kk[u + 1].y, d1); d1 = fmaf(qb.z, kk[u + 1].z, d1); d1 = fmaf(qb.w, kk[u + 1].w, d1);
                        d2 = fmaf(qc.x, kk[u + 2].x, d2); d2 = fmaf(qc.y, kk[u + 2].y, d2); d2 = fmaf(qc.z, kk[u + 2].z, d2); d2 = fmaf(qc.w, kk[u + 2].w, d2);
                        d3 = fmaf(qf.x, kk[u + 3].x, d3); d3 = fmaf(qf.y, kk[u + 3].y, d3); d3 = fmaf(qf.z, kk[u + 3].z, d3); d3 = fmaf(qf.w, kk[u + 3].w, d3);
                    }
                }
                if (d + 4 <= f4) {
                    float4 kk[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) kk[u] = kr[d + u];
                    const float4 qa = q4[d], qb = q4[d + 1], qc = q4[d + 2], qf = q4[d + 3];
                    d0 = fmaf(qa.x, kk[0].x, d0); d0 = fmaf(qa.y, kk[0].y, d0); d0 = fmaf(qa.z, kk[0].z, d0); d0 = fmaf(qa.w, kk[0].w, d0);
                    d1 = fmaf(qb.x, kk[1].x, d1); d1 = fmaf(qb.y, kk[1].y, d1); d1 = fmaf(qb.z, kk[1].z, d1); d1 = fmaf(qb.w, kk[1].w, d1);
                    d2 = fmaf(qc.x, kk[2].x, d2); d2 = fmaf(qc.y, kk[2].y, d2); d2 = fmaf(qc.z, kk[2].z, d2); d2 = fmaf(qc.w, kk[2].w, d2);
                    d3 = fmaf(qf.x, kk[3].x, d3); d3 = fmaf(qf.y, kk[3].y, d3); d3 = fmaf(qf.z, kk[3].z, d3); d3 = fmaf(qf.w, kk[3].w, d3);
                    d += 4;
                }
                for (; d < f4; d++) {
                    const float4 a = kr[d], qa = q4[d];
                    d0 = fmaf(qa.x, a.x, d0); d0 = fmaf(qa.y, a.y, d0); d0 = fmaf(qa.z, a.z, d0); d0 = fmaf(qa.w, a.w, d0);
                }
            }
            float dot = (d0 + d1) + (d2 + d3);
            for (uint32_t o = 1; o < lpp; o <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            if (ok) {
                score = dot * B.scale + mask_add;
                ok = isfinite(score);
                if (!ok) score = -INFINITY;
            }
        }
        float bm = score;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
        if (bm == -INFINITY) continue;
        const float new_m = fmaxf(m_val, bm);
        const float alpha = (m_val == -INFINITY) ? 0.0f : expf(m_val - new_m);
        const float wgt = ok ? expf(score - new_m) : 0.0f;
        l = l * alpha + warp_sum(seg == 0 ? wgt : 0.0f);
        m_val = new_m;
#pragma unroll
        for (int i = 0; i < NI; i++) acc[i] *= alpha;
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
            const float wj = __shfl_sync(0xffffffffu, wgt, (jj * lpp) & 31);
#pragma unroll
            for (int i = 0; i < NI; i++) acc[i] = fmaf(wj, vv[jj][i], acc[i]);   // rows past nj were loaded as zeros
        }
#pragma unroll 1
        for (uint32_t j0 = 8; j0 < nj; j0 += 8) {
#pragma unroll
            for (int jj = 0; jj < 8; jj++) {
                const bool in = j0 + jj < nj;
                const uint32_t sj = s0 + j0 + jj;
                const float* vr = (sj == s_new) ? sv : vbase + (size_t)sj * B.v_cs;
#pragma unroll
                for (int i = 0; i < NI; i++) vv[jj][i] = (in && lane + 32 * i < dh) ? vr[lane + 32 * i] : 0.0f;
            }
#pragma unroll
            for (int jj = 0; jj < 8; jj++) {
                const float wj = __shfl_sync(0xffffffffu, wgt, ((j0 + jj) * lpp) & 31);
#pragma unroll
                for (int i = 0; i < NI; i++) acc[i] = fmaf(wj, vv[jj][i], acc[i]);
            }
        }
    }
    if (lane == 0) { sh_m[warp] = m_val; sh_l[warp] = l; }
#pragma unroll
    for (int i = 0; i < NI; i++) sh_acc[warp][lane + 32 * i] = acc[i];
    __syncthreads();
    float gm = -INFINITY;
    for (int w = 0; w < kAttnFastWarps; w++) gm = fmaxf(gm, sh_m[w]);
    float gl = 0.0f;
    float wscale[kAttnFastWarps];
#pragma unroll
    for (int w = 0; w < kAttnFastWarps; w++) {
        wscale[w] = (sh_m[w] == -INFINITY) ? 0.0f : expf(sh_m[w] - gm);
        gl += sh_l[w] * wscale[w];
    }
    float* out1 = hd.attn_out;
    float* out2 = B.attn_buf + hd.buf_off;
    if (cl) {
        __shared__ float s_part[NI * 32 + 2];   // {m, l, acc[d_head]} of this split
        if (splits > 1) {
            for (uint32_t r = tid; r < dh; r += blockDim.x) {
                float a = 0.0f;
#pragma unroll
                for (int w = 0; w < kAttnFastWarps; w++) a += sh_acc[w][r] * wscale[w];
                s_part[2 + r] = a;
            }
            if (tid == 0) { s_part[0] = gm; s_part[1] = gl; }
        } else {
            const float inv_l = gl > 0.0f ? 1.0f / gl : 0.0f;
            for (uint32_t r = tid; r < dh; r += blockDim.x) {
                float a = 0.0f;
#pragma unroll
                for (int w = 0; w < kAttnFastWarps; w++) a += sh_acc[w][r] * wscale[w];
                out1[r] = a * inv_l; out2[r] = a * inv_l;
            }
        }
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        if (splits > 1 && sp == 0) {   // same merge, same order as the last arriver of the scratch path
            const uint32_t base = (uint32_t)__cvta_generic_to_shared(s_part);
            auto rd = [&](uint32_t q, uint32_t idx) -> float {
                uint32_t remote; float v;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(base + idx * 4), "r"(q));
                asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
                return v;
            };
            float G = -INFINITY;
            for (uint32_t q = 0; q < splits; q++) G = fmaxf(G, rd(q, 0));
            float L = 0.0f;
            for (uint32_t q = 0; q < splits; q++) {
                const float ms = rd(q, 0);
                L += (ms == -INFINITY) ? 0.0f : rd(q, 1) * expf(ms - G);
            }
            const float inv_L = L > 0.0f ? 1.0f / L : 0.0f;
            for (uint32_t r = tid; r < dh; r += blockDim.x) {
                float a = 0.0f;
                for (uint32_t q = 0; q < splits; q++) {
                    const float ms = rd(q, 0);
                    if (ms != -INFINITY) a += rd(q, 2 + r) * expf(ms - G);
                }
                out1[r] = a * inv_L; out2[r] = a * inv_L;
            }
        }
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        ZG_TRACE_MARK(2)
        return;
    }
    if (splits > 1) {
        __shared__ uint32_t s_last;
        float* mine = part + ((size_t)h * max_splits + sp) * (NI * 32 + 2);
        for (uint32_t r = tid; r < dh; r += blockDim.x) {
            float a = 0.0f;
#pragma unroll
            for (int w = 0; w < kAttnFastWarps; w++) a += sh_acc[w][r] * wscale[w];
            mine[2 + r] = a;
        }
        if (tid == 0) { mine[0] = gm; mine[1] = gl; }
        __syncthreads();
        if (tid == 0) {
            uint32_t old;
            asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(cnt + h) : "memory");
            const uint32_t last = (old == splits - 1) ? 1u : 0u;
            if (last) cnt[h] = 0u;   // re-arm for the next launch
            s_last = last;
        }
        __syncthreads();
        if (s_last) {
            const float* all = part + (size_t)h * max_splits * (NI * 32 + 2);
            float G = -INFINITY;
            for (uint32_t q = 0; q < splits; q++) G = fmaxf(G, __ldcg(all + (size_t)q * (NI * 32 + 2)));
            float L = 0.0f;
            for (uint32_t q = 0; q < splits; q++) {
                const float ms = __ldcg(all + (size_t)q * (NI * 32 + 2));
                L += (ms == -INFINITY) ? 0.0f : __ldcg(all + (size_t)q * (NI * 32 + 2) + 1) * expf(ms - G);
            }
            const float inv_L = L > 0.0f ? 1.0f / L : 0.0f;
            for (uint32_t r = tid; r < dh; r += blockDim.x) {
                float a = 0.0f;
                for (uint32_t q = 0; q < splits; q++) {
                    const float ms = __ldcg(all + (size_t)q * (NI * 32 + 2));
                    if (ms != -INFINITY) a += __ldcg(all + (size_t)q * (NI * 32 + 2) + 2 + r) * expf(ms - G);
                }
                out1[r] = a * inv_L; out2[r] = a * inv_L;
            }
        }
        ZG_TRACE_MARK(2)
        return;
    }
    const float inv_l = gl > 0.0f ? 1.0f / gl : 0.0f;
    for (uint32_t r = tid; r < dh; r += blockDim.x) {
        float a = 0.0f;
#pragma unroll
        for (int w = 0; w < kAttnFastWarps; w++) a += sh_acc[w][r] * wscale[w];
        out1[r] = a * inv_l; out2[r] = a * inv_l;
    }
    ZG_TRACE_MARK(2)
}

static inline bool attn_fast_ok(const ZgOp& op) {
    const auto& a = op.u.attention;
    return a.q_rs == 1 && a.k_rs == 1 && a.v_rs == 1 && a.dst_rs == 1 && (a.d_head % 4) == 0 && a.d_head <= 256 &&
           (a.k_off % 4) == 0 && (a.k_cs % 4) == 0;
}

static inline bool attn_prefill_ok(const ZgOp& op) {
    const auto& a = op.u.attention;
    return attn_fast_ok(op) && a.seq_q >= 16 && (a.d_head == 64 || a.d_head == 128) && (a.q_off % 4) == 0 && (a.q_cs % 4) == 0 &&
           (a.v_off % 4) == 0 && (a.v_cs % 4) == 0;
}

struct MMParams {
    uint32_t M, N, K;
    size_t a_rs, a_cs, b_rs, b_cs, a_off, b_off, d_off, d_rs;
};
// B contiguous along k (b_rs == 1), e.g. the tied LM head x @ token_embed^T
// (src/models/llama.zig:162-165): one warp per output column, float4 loads.
__global__ void k_matmul_kmajor(MMParams p, float* __restrict__ dst, const float* __restrict__ A,
                                const float* __restrict__ B) {
    ZG_TRACE_BEGIN(9)
    pdl_enter();
    ZG_TRACE_MARK(1)
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= p.N) return;
    const uint32_t m = blockIdx.y;
    const float* a = A + p.a_off + (size_t)m * p.a_rs;
    const float* b = B + p.b_off + (size_t)warp * p.b_cs;
    float acc = 0.0f;
    if (p.a_cs == 1 && ((p.b_cs & 3) == 0) && ((p.b_off & 3) == 0) && (((p.a_off + (size_t)m * p.a_rs) & 3) == 0) && ((p.K & 3) == 0)) {
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const float4* b4 = reinterpret_cast<const float4*>(b);
        for (uint32_t k = lane; k < p.K / 4; k += 32) {
            float4 x = a4[k], y = __ldcs(b4 + k);
            acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc);
            acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
        }
    } else {
        for (uint32_t k = lane; k < p.K; k += 32) acc = fmaf(a[(size_t)k * p.a_cs], b[k], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) dst[p.d_off + (size_t)m * p.d_rs + warp] = acc;
}
// general strides: one thread per output (coalesced over n when b_cs == 1)
__global__ void k_matmul_general(MMParams p, float* __restrict__ dst, const float* __restrict__ A,
                                 const float* __restrict__ B) {
    pdl_enter();
    uint32_t n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
    if (n >= p.N) return;
    float acc = 0.0f;
    for (uint32_t k = 0; k < p.K; k++)
        acc = fmaf(A[p.a_off + (size_t)m * p.a_rs + (size_t)k * p.a_cs], B[p.b_off + (size_t)k * p.b_rs + (size_t)n * p.b_cs], acc);
    dst[p.d_off + (size_t)m * p.d_rs + n] = acc;
}

// ── chained small ops ───────────────────────────────────────────────────────────────────────────
// A decode step is a long dependency chain of tiny ops between the matvecs (norms, gamma broadcast, residual adds,
// RoPE, KV-cache stores, the SiLU chain): a few thousand floats each, so a launch per op is pure latency
// (~4 us per dependent kernel, ~1 us of work).  Runs of such ops are executed by ONE CTA instead: ops in table
// order, every op exactly as its DeviceOp defines it (all intermediate buffers are written), `sync` marks the first
// op of a new dependency level (block barrier: CTA-scope visibility of the previous level's global writes).
// Plain (coherent) loads only: data read here may have been written earlier in the same kernel.
constexpr int kChainThreads = 256;   // x 64 registers = 16K: fits next to two resident matvec CTAs (2 x 20K), so PDL can make it resident early
static_assert(sizeof(ZgChainOp) % 4 == 0, "table is copied word-wise");

__device__ __forceinline__ float chain_block_sum(float v, float* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = 0.0f;
#pragma unroll
    for (int i = 0; i < kChainThreads / 32; i++) r += sh[i];
    return r;
}

// The element-parallel op kinds, run by `nt` cooperating threads (a whole CTA, or one warp of it).
__device__ __forceinline__ void chain_small_op(const ZgChainOp* o, uint32_t t, uint32_t nt, const uint32_t* __restrict__ d_dyn) {
    float* dst = o->dst;
    const float* s0 = o->s0;
    const float* s1 = o->s1;
    switch (o->kind) {
        case ZG_OP_ELEMENTWISE: {
            const uint32_t op = o->u[0], n = o->u[1];
            if ((op == ZG_EW_ADD || op == ZG_EW_MUL) && (n & 3u) == 0 && (((size_t)dst | (size_t)s0 | (size_t)s1) & 15u) == 0) {
                // 128-bit accesses, two in flight per thread: these ops are pure load latency
                float4* d4 = reinterpret_cast<float4*>(dst);
                const float4* a4 = reinterpret_cast<const float4*>(s0);
                const float4* b4 = reinterpret_cast<const float4*>(s1);
                const uint32_t n4 = n >> 2;
                for (uint32_t j = t; j < n4; j += 2 * nt) {
                    const uint32_t j2 = j + nt;
                    const float4 a = a4[j], b = b4[j];
                    float4 c = make_float4(0.f, 0.f, 0.f, 0.f), e = c;
                    if (j2 < n4) { c = a4[j2]; e = b4[j2]; }
                    d4[j] = op == ZG_EW_ADD ? make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w) : make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
                    if (j2 < n4) d4[j2] = op == ZG_EW_ADD ? make_float4(c.x + e.x, c.y + e.y, c.z + e.z, c.w + e.w) : make_float4(c.x * e.x, c.y * e.y, c.z * e.z, c.w * e.w);
                }
                break;
            }
            for (uint32_t j = t; j < n; j += nt) {
                const float a = s0[j];
                float r;
                if (op == ZG_EW_ADD) r = a + s1[j];
                else if (op == ZG_EW_MUL) r = a * s1[j];
                else r = apply_unary(op, a);
                dst[j] = r;
            }
            break;
        }
        case ZG_OP_FUSED_ELEMENTWISE: {
            const uint32_t n_steps = o->u[0], n = o->u[1];
            const ZgDevStep* steps = o->steps;
            for (uint32_t j = t; j < n; j += nt) {
                float v = s0[j];
                for (uint32_t s = 0; s < n_steps; s++) {
                    const uint32_t sop = steps[s].op, sw = steps[s].is_swapped;
                    if (sop == ZG_EW_ADD) { const float x = steps[s].sec[j]; v = sw ? x + v : v + x; }
                    else if (sop == ZG_EW_MUL) { const float x = steps[s].sec[j]; v = sw ? x * v : v * x; }
                    else v = apply_unary(sop, v);
                }
                dst[j] = v;
            }
            break;
        }
        case ZG_OP_REPEAT: {
            const uint32_t mode = o->u[0], n = o->u[1], src_n = o->u[2], src_offset = o->u[15];
            if ((mode == 1 || mode == 2) && ((n | src_n) & 3u) == 0 && (((size_t)dst | (size_t)s0) & 15u) == 0) {
                float4* d4 = reinterpret_cast<float4*>(dst);
                const float4* a4 = reinterpret_cast<const float4*>(s0);
                const uint32_t n4 = n >> 2, sn4 = src_n >> 2;
                for (uint32_t j = t; j < n4; j += nt) d4[j] = a4[mode == 1 ? j : j % sn4];
                break;
            }
            for (uint32_t gid = t; gid < n; gid += nt) {
                float v;
                if (mode == 0) v = s0[0];
                else if (mode == 1) v = s0[gid];
                else if (mode == 2) v = s0[gid % src_n];
                else {
                    uint32_t idx = gid, sidx = src_offset;
#pragma unroll
                    for (int dim = 3; dim >= 0; dim--) {
                        const uint32_t ds = o->u[11 + dim];
                        const uint32_t coord = idx / ds;
                        idx = idx % ds;
                        sidx += (coord % o->u[3 + dim]) * o->u[7 + dim];
                    }
                    v = s0[sidx];
                }
                dst[gid] = v;
            }
            break;
        }
        case ZG_OP_SLICE_ASSIGN: {
            const uint32_t rows = o->u[0], cols = o->u[1], drs = o->u[2], dcs = o->u[3], soff = o->u[4], srs = o->u[5], scs = o->u[6];
            const uint32_t doff = d_dyn[o->dyn];
            const uint32_t total = rows * cols;
            for (uint32_t j = t; j < total; j += nt) {
                const uint32_t row = j % rows, col = j / rows;
                dst[(size_t)doff + (size_t)row * drs + (size_t)col * dcs] = s0[(size_t)soff + (size_t)row * srs + (size_t)col * scs];
            }
            break;
        }
        case ZG_OP_ROPE: {
            const uint32_t hd = o->u[0], seq_len = o->u[1], s_off = o->u[2], c_off = o->u[3], d_off = o->u[4], s_rs = o->u[5], s_cs = o->u[6], c_cs = o->u[7];
            const uint32_t total = hd * seq_len;
            for (uint32_t j = t; j < total; j += nt) {
                const uint32_t pair = j % hd, col = j / hd;
                const float x_lo = s0[(size_t)s_off + (size_t)pair * s_rs + (size_t)col * s_cs];
                const float x_hi = s0[(size_t)s_off + (size_t)(pair + hd) * s_rs + (size_t)col * s_cs];
                const float c = s1[(size_t)c_off + pair + (size_t)col * c_cs];
                const float sn = s1[(size_t)c_off + pair + hd + (size_t)col * c_cs];
                dst[(size_t)d_off + pair + (size_t)col * 2 * hd] = __fsub_rn(__fmul_rn(x_lo, c), __fmul_rn(x_hi, sn));
                dst[(size_t)d_off + pair + hd + (size_t)col * 2 * hd] = __fadd_rn(__fmul_rn(x_hi, c), __fmul_rn(x_lo, sn));
            }
            break;
        }
        default: break;
    }
}

__global__ void __launch_bounds__(kChainThreads, 4)
k_chain(const ZgChainOp* __restrict__ tab, uint32_t count, const uint32_t* __restrict__ d_dyn) {
    __shared__ float sh[32];
    __shared__ ZgDevStep s_steps[16];
    __shared__ __align__(16) ZgChainOp s_tab[kZgChainMaxOps];   // the whole op table: one coalesced read instead of a dependent load per op
    const uint32_t tid = threadIdx.x;
    ZG_TRACE_BEGIN(8)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    {   // the op table is immutable while programs run: copy it before waiting for the previous kernel
        const uint32_t words = count * (uint32_t)(sizeof(ZgChainOp) / 4);
        const uint32_t* g = reinterpret_cast<const uint32_t*>(tab);
        uint32_t* l = reinterpret_cast<uint32_t*>(s_tab);
        for (uint32_t j = tid; j < words; j += kChainThreads) l[j] = __ldg(g + j);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    ZG_TRACE_MARK(1)
    __syncthreads();
    uint32_t i = 0;
    while (i < count) {
        const ZgChainOp* o = s_tab + i;
        if (o->sync) __syncthreads();
        if (o->group) {   // a run of tiny independent ops of one dependency level (per-head rope / cache stores): one warp each
            const uint32_t g = o->group;
            for (uint32_t k = tid >> 5; k < g; k += kChainThreads / 32) chain_small_op(o + k, tid & 31, 32, d_dyn);
            i += g;
            continue;
        }
        i++;
        float* dst = o->dst;
        const float* s0 = o->s0;
        switch (o->kind) {
            case ZG_OP_RMSNORM: {   // same arithmetic as k_rmsnorm (rows run one after the other)
                const uint32_t rows = o->u[0], cols = o->u[1];
                const float eps = o->f;
                for (uint32_t r = 0; r < rows; r++) {
                    const float* s = s0 + (size_t)r * cols;
                    float* d = dst + (size_t)r * cols;
                    float ss = 0.0f;
                    if ((cols & 3u) == 0 && (((size_t)s | (size_t)d) & 15u) == 0 && cols <= 16 * kChainThreads) {   // <= 4096
                        // the row stays in registers between the two passes (<= 4 float4 per thread)
                        const float4* s4 = reinterpret_cast<const float4*>(s);
                        const uint32_t c4 = cols >> 2;
                        float4 x[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) { const uint32_t j = tid + u * kChainThreads; x[u] = j < c4 ? s4[j] : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
                        for (int u = 0; u < 4; u++) ss += (x[u].x * x[u].x + x[u].y * x[u].y) + (x[u].z * x[u].z + x[u].w * x[u].w);
                        ss = chain_block_sum(ss, sh);
                        const float inv_rms = 1.0f / sqrtf(ss / (float)cols + eps);
                        float4* d4 = reinterpret_cast<float4*>(d);
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const uint32_t j = tid + u * kChainThreads;
                            if (j < c4) d4[j] = make_float4(x[u].x * inv_rms, x[u].y * inv_rms, x[u].z * inv_rms, x[u].w * inv_rms);
                        }
                        continue;
                    }
                    for (uint32_t j = tid; j < cols; j += kChainThreads) { const float x = s[j]; ss += x * x; }
                    ss = chain_block_sum(ss, sh);
                    const float inv_rms = 1.0f / sqrtf(ss / (float)cols + eps);
                    for (uint32_t j = tid; j < cols; j += kChainThreads) d[j] = s[j] * inv_rms;
                }
                break;
            }
            case kZgChainFusedNorm: {
                // [a + b ->] sum ; rmsnorm(sum) -> bare ; gamma broadcast -> gamma_rep ; bare * gamma_rep -> dst, one pass in
                // registers (cols <= 8192, cols % 4 == 0): one load round, one block reduction, one store round per row.
                const uint32_t rows = o->u[0], cols = o->u[1], c4 = cols >> 2;
                const float eps = o->f;
                const float4* a4 = reinterpret_cast<const float4*>(s0);
                const float4* b4 = reinterpret_cast<const float4*>(o->s1);
                float4* sum4 = reinterpret_cast<float4*>(((uint64_t)o->u[3] << 32) | o->u[2]);
                float4* bare4 = reinterpret_cast<float4*>(((uint64_t)o->u[5] << 32) | o->u[4]);
                const float4* g4 = reinterpret_cast<const float4*>(((uint64_t)o->u[7] << 32) | o->u[6]);
                float4* grep4 = reinterpret_cast<float4*>(((uint64_t)o->u[9] << 32) | o->u[8]);
                float4* n4 = reinterpret_cast<float4*>(dst);
                constexpr int NV = 4;   // float4 per thread held in registers: cols <= 4 * NV * kChainThreads = 4096 in one read
                if (c4 <= NV * kChainThreads) {
                    float4 gam[NV];
#pragma unroll
                    for (int u = 0; u < NV; u++) { const uint32_t j = tid + u * kChainThreads; gam[u] = j < c4 ? g4[j] : make_float4(0.f, 0.f, 0.f, 0.f); }
                    for (uint32_t r = 0; r < rows; r++) {
                        const size_t ro = (size_t)r * c4;
                        float4 x[NV];
#pragma unroll
                        for (int u = 0; u < NV; u++) { const uint32_t j = tid + u * kChainThreads; x[u] = j < c4 ? a4[ro + j] : make_float4(0.f, 0.f, 0.f, 0.f); }
                        if (b4) {
#pragma unroll
                            for (int u = 0; u < NV; u++) {
                                const uint32_t j = tid + u * kChainThreads;
                                if (j >= c4) continue;
                                const float4 t = b4[ro + j];
                                x[u] = make_float4(x[u].x + t.x, x[u].y + t.y, x[u].z + t.z, x[u].w + t.w);
                                sum4[ro + j] = x[u];
                            }
                        }
                        float ss = 0.0f;
#pragma unroll
                        for (int u = 0; u < NV; u++) ss += (x[u].x * x[u].x + x[u].y * x[u].y) + (x[u].z * x[u].z + x[u].w * x[u].w);
                        ss = chain_block_sum(ss, sh);
                        const float inv_rms = 1.0f / sqrtf(ss / (float)cols + eps);
#pragma unroll
                        for (int u = 0; u < NV; u++) {
                            const uint32_t j = tid + u * kChainThreads;
                            if (j >= c4) continue;
                            const float4 bz = make_float4(x[u].x * inv_rms, x[u].y * inv_rms, x[u].z * inv_rms, x[u].w * inv_rms);
                            bare4[ro + j] = bz; grep4[ro + j] = gam[u];
                            n4[ro + j] = make_float4(bz.x * gam[u].x, bz.y * gam[u].y, bz.z * gam[u].z, bz.w * gam[u].w);
                        }
                    }
                } else {   // long rows: sum pass, then a second read of the (L2-resident) inputs
                    for (uint32_t r = 0; r < rows; r++) {
                        const size_t ro = (size_t)r * c4;
                        float ss = 0.0f;
#pragma unroll 4
                        for (uint32_t j = tid; j < c4; j += kChainThreads) {
                            float4 x = a4[ro + j];
                            if (b4) { const float4 t = b4[ro + j]; x = make_float4(x.x + t.x, x.y + t.y, x.z + t.z, x.w + t.w); sum4[ro + j] = x; }
                            ss += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
                        }
                        ss = chain_block_sum(ss, sh);   // its barriers also publish sum4 to the whole CTA
                        const float inv_rms = 1.0f / sqrtf(ss / (float)cols + eps);
                        const float4* x4 = b4 ? sum4 : a4;
#pragma unroll 4
                        for (uint32_t j = tid; j < c4; j += kChainThreads) {
                            const float4 x = x4[ro + j], g = g4[j];
                            const float4 bz = make_float4(x.x * inv_rms, x.y * inv_rms, x.z * inv_rms, x.w * inv_rms);
                            bare4[ro + j] = bz; grep4[ro + j] = g;
                            n4[ro + j] = make_float4(bz.x * g.x, bz.y * g.y, bz.z * g.z, bz.w * g.w);
                        }
                    }
                }
                break;
            }
            case kZgChainEwMul: {
                // fused_elementwise chain -> mid ; mid * other -> dst, four independent elements in flight per thread
                const uint32_t n_steps = o->u[0], n = o->u[1];
                __syncthreads();
                if (tid < n_steps && tid < 16) s_steps[tid] = o->steps[tid];
                __syncthreads();
                const ZgDevStep* steps = s_steps;
                float* mid = reinterpret_cast<float*>(((uint64_t)o->u[3] << 32) | o->u[2]);
                const float* other = o->s1;
                for (uint32_t jb = tid; jb < n; jb += 4 * kChainThreads) {
                    float v[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) { const uint32_t j = jb + u * kChainThreads; v[u] = j < n ? s0[j] : 0.0f; }
                    for (uint32_t st = 0; st < n_steps; st++) {
                        const uint32_t sop = steps[st].op, sw = steps[st].is_swapped;
                        const float* sec = steps[st].sec;
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const uint32_t j = jb + u * kChainThreads;
                            if (j >= n) continue;
                            if (sop == ZG_EW_ADD) { const float t = sec[j]; v[u] = sw ? t + v[u] : v[u] + t; }
                            else if (sop == ZG_EW_MUL) { const float t = sec[j]; v[u] = sw ? t * v[u] : v[u] * t; }
                            else v[u] = apply_unary(sop, v[u]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const uint32_t j = jb + u * kChainThreads;
                        if (j < n) { mid[j] = v[u]; dst[j] = v[u] * other[j]; }
                    }
                }
                break;
            }
            default: chain_small_op(o, tid, kChainThreads, d_dyn); break;
        }
    }
    ZG_TRACE_MARK(2)
}

inline unsigned blocks_for(uint32_t n, unsigned bs, unsigned cap = 4096) {
    unsigned b = (n + bs - 1) / bs;
    if (b < 1) b = 1;
    return b > cap ? cap : b;
}

} // namespace

static uint32_t repeat_mode(const ZgOp& op, size_t* src_n_out) {
    const auto& rp = op.u.repeat;
    const uint32_t* ne = rp.src_ne; const uint32_t* sst = rp.src_strides;
    const size_t src_n = (size_t)ne[0] * ne[1] * ne[2] * ne[3];
    *src_n_out = src_n;
    if (src_n == 1) return 0;
    if (src_n >= rp.n) return 1;
    if (rp.n % src_n == 0 && sst[0] == 1 && (ne[1] <= 1 || sst[1] == ne[0]) && (ne[2] <= 1 || sst[2] == ne[0] * ne[1]) &&
        (ne[3] <= 1 || sst[3] == ne[0] * ne[1] * ne[2])) return 2;
    return 3;
}

bool zg_launch_op(ZgCudaCtx* ctx, const ZgOp& op, float* const* bufs, const uint32_t* d_dyn,
                  uint32_t op_index, const ZgDevStep* d_steps, cudaStream_t st) {
    (void)ctx;
    switch (op.tag) {
        case ZG_OP_ELEMENTWISE: {
            const auto& e = op.u.elementwise;
            if (e.n == 0) return true;
            launch_k(k_elementwise, dim3(blocks_for(e.n, 256)), dim3(256), st, e.op, bufs[e.dst] + e.dst_offset, bufs[e.src0] + e.src0_offset,
                                                                bufs[e.src1] + e.src1_offset, e.n);
            break;
        }
        case ZG_OP_FUSED_ELEMENTWISE: {
            const auto& f = op.u.fused_elementwise;
            if (f.n == 0) return true;
            launch_k(k_fused_elementwise, dim3(blocks_for(f.n, 256)), dim3(256), st, d_steps, (uint32_t)f.n_steps, bufs[f.dst] + f.dst_offset,
                                                                      bufs[f.src] + f.src_offset, f.n);
            break;
        }
        case ZG_OP_SOFTMAX: {
            const auto& s = op.u.softmax;
            if (s.rows == 0 || s.cols == 0) return true;
            launch_k(k_softmax, dim3(s.rows), dim3(256), st, bufs[s.dst] + s.dst_offset, bufs[s.src] + s.src_offset, s.cols);
            break;
        }
        case ZG_OP_LAYERNORM: {
            const auto& l = op.u.layernorm;
            if (l.rows == 0 || l.cols == 0) return true;
            launch_k(k_layernorm, dim3(l.rows), dim3(256), st, bufs[l.dst] + l.dst_offset, bufs[l.src] + l.src_offset, l.cols, l.eps);
            break;
        }
        case ZG_OP_RMSNORM: {
            const auto& r = op.u.rmsnorm;
            if (r.rows == 0 || r.cols == 0) return true;
            launch_k(k_rmsnorm, dim3(r.rows), dim3(1024), st, bufs[r.dst] + r.dst_offset, bufs[r.src] + r.src_offset, r.cols, r.eps);
            break;
        }
        case ZG_OP_REDUCE: {
            const auto& r = op.u.reduce;
            if (r.n_out == 0) return true;
            launch_k(k_reduce, dim3(blocks_for(r.n_out * 32, 256, 65535)), dim3(256), st, r.op == ZG_EW_MAX, bufs[r.dst] + r.dst_offset,
                                                                           bufs[r.src] + r.src_offset, r.n_out, r.reduce_size);
            break;
        }
        case ZG_OP_REPEAT: {
            const auto& rp = op.u.repeat;
            if (rp.n == 0) return true;
            RepeatParams p;
            size_t src_n = 0;
            p.mode = repeat_mode(op, &src_n);
            p.n = rp.n; p.src_n = (uint32_t)src_n; p.src_offset = rp.src_offset;
            for (int i = 0; i < 4; i++) { p.src_ne[i] = rp.src_ne[i]; p.src_strides[i] = rp.src_strides[i]; p.dst_strides[i] = rp.dst_strides[i]; }
            const float* src = (p.mode == 3) ? bufs[rp.src] : bufs[rp.src] + rp.src_offset;
            launch_k(k_repeat, dim3(blocks_for(rp.n, 256)), dim3(256), st, p, bufs[rp.dst] + rp.dst_offset, src);
            break;
        }
        case ZG_OP_SLICE_ASSIGN:
        case ZG_OP_ROPE:
        case ZG_OP_ATTENTION:
            zg_set_error("internal: batched op kind %u reached the single-op launcher", op.tag);
            return false;
        case ZG_OP_MATMUL: {
            const auto& m = op.u.matmul;
            const ZgMatMulGeometry& g = m.geom;
            if (g.M == 0 || g.N == 0) return true;
            MMParams p;
            p.M = (uint32_t)g.M; p.N = (uint32_t)g.N; p.K = (uint32_t)g.K;
            p.a_rs = g.a_row_stride; p.a_cs = g.a_col_stride; p.b_rs = g.b_row_stride; p.b_cs = g.b_col_stride;
            p.a_off = g.a_offset; p.b_off = g.b_offset; p.d_off = g.dst_offset; p.d_rs = g.dst_row_stride;
            if (g.b_row_stride == 1 && g.K >= 32) {
                dim3 grid((unsigned)((g.N * 32 + 255) / 256), (unsigned)g.M);
                launch_k(k_matmul_kmajor, dim3(grid), dim3(256), st, p, bufs[m.dst], bufs[m.a], bufs[m.b]);
            } else {
                dim3 grid((unsigned)((g.N + 127) / 128), (unsigned)g.M);
                launch_k(k_matmul_general, dim3(grid), dim3(128), st, p, bufs[m.dst], bufs[m.a], bufs[m.b]);
            }
            break;
        }
        default:
            zg_set_error("unsupported DeviceOp tag %u", op.tag);
            return false;
    }
    ZG_COUNT_LAUNCH();
    return true;
}

// ── chained small ops: host side ─────────────────────────────────────────────────────────────────
// Elements one CTA would have to walk for this op, or 0 when the op kind cannot be chained.
size_t zg_chain_work(const ZgOp& op) {
    switch (op.tag) {
        case ZG_OP_ELEMENTWISE: return op.u.elementwise.n ? op.u.elementwise.n : 1;
        case ZG_OP_FUSED_ELEMENTWISE: return op.u.fused_elementwise.n ? (size_t)op.u.fused_elementwise.n * (1 + op.u.fused_elementwise.n_steps) : 1;   // transcendental chains: one CTA is slow
        case ZG_OP_RMSNORM: return op.u.rmsnorm.rows <= 8 ? (size_t)op.u.rmsnorm.rows * op.u.rmsnorm.cols + 1 : 0;
        case ZG_OP_REPEAT: return op.u.repeat.n ? op.u.repeat.n : 1;
        case ZG_OP_SLICE_ASSIGN: return (size_t)op.u.slice_assign.rows * op.u.slice_assign.cols + 1;
        case ZG_OP_ROPE: return (size_t)op.u.rope.half_d * op.u.rope.seq_len * 2 + 1;
        default: return 0;
    }
}

bool zg_fill_chain_op(const ZgOp& op, float* const* bufs, uint32_t op_index, const ZgDevStep* d_steps, bool sync, ZgChainOp* c) {
    memset(c, 0, sizeof(*c));
    c->kind = op.tag; c->sync = sync ? 1u : 0u; c->dyn = op_index;
    switch (op.tag) {
        case ZG_OP_ELEMENTWISE: {
            const auto& e = op.u.elementwise;
            c->dst = bufs[e.dst] + e.dst_offset; c->s0 = bufs[e.src0] + e.src0_offset; c->s1 = bufs[e.src1] + e.src1_offset;
            c->u[0] = e.op; c->u[1] = e.n;
            return true;
        }
        case ZG_OP_FUSED_ELEMENTWISE: {
            const auto& f = op.u.fused_elementwise;
            c->dst = bufs[f.dst] + f.dst_offset; c->s0 = bufs[f.src] + f.src_offset; c->steps = d_steps;
            c->u[0] = (uint32_t)f.n_steps; c->u[1] = f.n;
            return true;
        }
        case ZG_OP_RMSNORM: {
            const auto& r = op.u.rmsnorm;
            c->dst = bufs[r.dst] + r.dst_offset; c->s0 = bufs[r.src] + r.src_offset;
            c->u[0] = r.rows; c->u[1] = r.cols; c->f = r.eps;
            return true;
        }
        case ZG_OP_REPEAT: {
            const auto& rp = op.u.repeat;
            size_t src_n = 0;
            const uint32_t mode = repeat_mode(op, &src_n);
            c->dst = bufs[rp.dst] + rp.dst_offset;
            c->s0 = (mode == 3) ? bufs[rp.src] : bufs[rp.src] + rp.src_offset;
            c->u[0] = mode; c->u[1] = rp.n; c->u[2] = (uint32_t)src_n; c->u[15] = rp.src_offset;
            for (int i = 0; i < 4; i++) { c->u[3 + i] = rp.src_ne[i]; c->u[7 + i] = rp.src_strides[i]; c->u[11 + i] = rp.dst_strides[i]; }
            return true;
        }
        case ZG_OP_SLICE_ASSIGN: {
            const auto& sa = op.u.slice_assign;
            c->dst = bufs[sa.dst]; c->s0 = bufs[sa.src];
            c->u[0] = sa.rows; c->u[1] = sa.cols; c->u[2] = sa.dst_row_stride; c->u[3] = sa.dst_col_stride;
            c->u[4] = sa.src_offset; c->u[5] = sa.src_row_stride; c->u[6] = sa.src_col_stride;
            return true;
        }
        case ZG_OP_ROPE: {
            const auto& r = op.u.rope;
            c->dst = bufs[r.dst]; c->s0 = bufs[r.src]; c->s1 = bufs[r.cos_sin];
            c->u[0] = r.half_d; c->u[1] = r.seq_len; c->u[2] = r.src_off; c->u[3] = r.cs_off; c->u[4] = r.dst_off;
            c->u[5] = r.src_rs; c->u[6] = r.src_cs; c->u[7] = r.cs_cs;
            return true;
        }
        default: zg_set_error("internal: op kind %u cannot be chained", op.tag); return false;
    }
}

static inline void put_ptr(ZgChainOp* c, int at, const void* ptr) {
    c->u[at] = (uint32_t)((uint64_t)ptr & 0xFFFFFFFFu); c->u[at + 1] = (uint32_t)((uint64_t)ptr >> 32);
}

bool zg_fill_chain_norm(const ZgNormMacro& m, bool sync, ZgChainOp* c) {
    memset(c, 0, sizeof(*c));
    c->kind = kZgChainFusedNorm; c->sync = sync ? 1u : 0u;
    c->dst = m.norm; c->s0 = m.a; c->s1 = m.b; c->f = m.eps;
    c->u[0] = m.rows; c->u[1] = m.cols;
    put_ptr(c, 2, m.sum); put_ptr(c, 4, m.bare); put_ptr(c, 6, m.gamma); put_ptr(c, 8, m.gamma_rep);
    return true;
}

bool zg_fill_chain_ewmul(const ZgEwMulMacro& m, bool sync, ZgChainOp* c) {
    memset(c, 0, sizeof(*c));
    c->kind = kZgChainEwMul; c->sync = sync ? 1u : 0u;
    c->dst = m.dst; c->s0 = m.src; c->s1 = m.other; c->steps = m.steps;
    c->u[0] = m.n_steps; c->u[1] = m.n;
    put_ptr(c, 2, m.mid);
    return true;
}

bool zg_launch_peer_allreduce(float* buf, size_t n, const ZgPeerComm& pc, cudaStream_t st) {
    if (n == 0) return true;
    launch_k(k_allreduce_peer, dim3(kZgPeerCtas), dim3(256), st, buf, (uint32_t)(n >> 1), pc);
    ZG_COUNT_LAUNCH();
    return true;
}

bool zg_launch_peer_allreduce_norm(float* buf, size_t n, const ZgPeerComm& pc, const ZgNormMacro& m, cudaStream_t st) {
    if (n == 0) return true;
    launch_k(k_allreduce_norm, dim3(kZgPeerCtas), dim3(256), st, buf, (uint32_t)(n >> 1), pc, m);
    ZG_COUNT_LAUNCH();
    return true;
}

bool zg_launch_norm_macro(const ZgNormMacro& m, cudaStream_t st) {
    if (m.rows == 0) return true;
    static const bool use_cluster = [] { const char* e = getenv("ZG_CUDA_NORM_CLUSTER"); return !(e && e[0] == '0'); }();
    if (use_cluster && (m.cols >> 2) <= kNormClusterCtas * 256) {
        uint32_t n_cta = 1;
        while (n_cta * 256 < (m.cols >> 2)) n_cta *= 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(m.rows * n_cta); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = n_cta; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = g_zg_pdl ? 2 : 1;
        if (cudaLaunchKernelEx(&cfg, k_norm_macro_cluster, m) == cudaSuccess) { ZG_COUNT_LAUNCH(); return true; }
        cudaGetLastError();   // fall through to the single-CTA form
    }
    launch_k(k_norm_macro, dim3(m.rows), dim3(1024), st, m);
    ZG_COUNT_LAUNCH();
    return true;
}

bool zg_launch_ewmul(const ZgEwMulMacro& m, cudaStream_t st) {
    if (m.n == 0) return true;
    launch_k(k_fused_ew_mul, dim3(blocks_for(m.n, 128)), dim3(128), st, m.steps, m.n_steps, m.mid, m.src, m.n, m.other, m.dst);
    ZG_COUNT_LAUNCH();
    return true;
}

bool zg_launch_chain(const ZgChainOp* d_ops, uint32_t count, const uint32_t* d_dyn, cudaStream_t st) {
    if (count == 0) return true;
    launch_k(k_chain, dim3(1), dim3(kChainThreads), st, d_ops, count, d_dyn);
    ZG_COUNT_LAUNCH();
    return true;
}

// ── batched per-head ops ────────────────────────────────────────────────────────────────────────
bool zg_op_is_batched(uint32_t tag) { return tag == ZG_OP_SLICE_ASSIGN || tag == ZG_OP_ROPE || tag == ZG_OP_ATTENTION; }

// Two ops may share a launch when their kind and work shape agree (grid dimensions are common to the batch).
uint64_t zg_batch_signature(const ZgOp& op) {
    switch (op.tag) {
        case ZG_OP_SLICE_ASSIGN: return ((uint64_t)op.tag << 56) | ((uint64_t)(op.u.slice_assign.rows & 0xFFFFFFF) << 28) | (op.u.slice_assign.cols & 0xFFFFFFF);
        case ZG_OP_ROPE: return ((uint64_t)op.tag << 56) | ((uint64_t)(op.u.rope.half_d & 0xFFFFFFF) << 28) | (op.u.rope.seq_len & 0xFFFFFFF);
        case ZG_OP_ATTENTION: return ((uint64_t)op.tag << 56) | ((uint64_t)attn_fast_ok(op) << 55) | ((uint64_t)attn_prefill_ok(op) << 54) | ((uint64_t)(op.u.attention.d_head & 0x7FFFFFF) << 28) | (op.u.attention.seq_q & 0xFFFFFFF);
        default: return 0;
    }
}

bool zg_fill_batch_entry(const ZgOp& op, float* const* bufs, uint32_t op_index, ZgBatchEntry* e) {
    memset(e, 0, sizeof(*e));
    e->dyn = op_index;
    switch (op.tag) {
        case ZG_OP_SLICE_ASSIGN: {
            const auto& sa = op.u.slice_assign;
            e->dst = bufs[sa.dst]; e->s0 = bufs[sa.src];
            e->u[0] = sa.rows; e->u[1] = sa.cols; e->u[2] = sa.dst_row_stride; e->u[3] = sa.dst_col_stride;
            e->u[4] = sa.src_offset; e->u[5] = sa.src_row_stride; e->u[6] = sa.src_col_stride;
            return true;
        }
        case ZG_OP_ROPE: {
            const auto& r = op.u.rope;
            e->dst = bufs[r.dst]; e->s0 = bufs[r.src]; e->s1 = bufs[r.cos_sin];
            e->u[0] = r.half_d; e->u[1] = r.seq_len; e->u[2] = r.src_off; e->u[3] = r.cs_off; e->u[4] = r.dst_off;
            e->u[5] = r.src_rs; e->u[6] = r.src_cs; e->u[7] = r.cs_cs;
            return true;
        }
        case ZG_OP_ATTENTION: {
            const auto& a = op.u.attention;
            if (a.d_head > 512) { zg_set_error("attention: d_head %u > 512", a.d_head); return false; }
            e->dst = bufs[a.dst]; e->s0 = bufs[a.q]; e->s1 = bufs[a.k]; e->s2 = bufs[a.v]; e->s3 = bufs[a.mask];
            e->f = a.scale;
            e->u[0] = a.has_mask; e->u[1] = a.d_head; e->u[2] = a.seq_q;
            e->u[3] = a.q_off; e->u[4] = a.k_off; e->u[5] = a.v_off; e->u[6] = a.mask_off; e->u[7] = a.dst_off;
            e->u[8] = a.q_rs; e->u[9] = a.q_cs; e->u[10] = a.k_rs; e->u[11] = a.k_cs; e->u[12] = a.v_rs; e->u[13] = a.v_cs;
            e->u[14] = a.mask_rs; e->u[15] = a.mask_cs; e->u[16] = a.dst_rs; e->u[17] = a.dst_cs;
            return true;
        }
        default: zg_set_error("internal: op kind %u has no batch entry", op.tag); return false;
    }
}

bool zg_launch_attention_layer(const ZgAttnBlock* d_blk, uint32_t n_heads, uint32_t d_head, uint32_t max_splits, const uint32_t* d_dyn, float* part,
                               uint32_t* cnt, cudaStream_t st) {
    if (n_heads == 0) return true;
    const dim3 grid(n_heads, max_splits), block(kAttnFastWarps * 32);
    static const bool want_cl = [] { const char* e = getenv("ZG_CUDA_ATTN_CLUSTER"); return !(e && e[0] == '0'); }();
    // the splits of a head = one cluster, merged in DSMEM.  Clusters of up to 4 only: 16 clusters of 8 CTAs (a 4-way 70B shard)
    // were measured 13 % slower per layer than the scratch merge — a cluster needs all its slots in one GPC at once, and the
    // neighbouring kernels' resident CTAs delay that — while clusters of 2 and 4 gain 1-3 %.
    const uint32_t cl = (want_cl && max_splits >= 2 && max_splits <= 4) ? 1u : 0u;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    unsigned na = 0;
    if (g_zg_pdl) { attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; na++; }
    if (cl) { attr[na].id = cudaLaunchAttributeClusterDimension; attr[na].val.clusterDim.x = 1; attr[na].val.clusterDim.y = max_splits; attr[na].val.clusterDim.z = 1; na++; }
    cfg.attrs = attr; cfg.numAttrs = na;
    cudaError_t le;
    if (d_head <= 64) le = cudaLaunchKernelEx(&cfg, k_attention_layer<2>, d_blk, d_dyn, part, cnt, max_splits, cl);
    else if (d_head <= 128) le = cudaLaunchKernelEx(&cfg, k_attention_layer<4>, d_blk, d_dyn, part, cnt, max_splits, cl);
    else le = cudaLaunchKernelEx(&cfg, k_attention_layer<8>, d_blk, d_dyn, part, cnt, max_splits, cl);
    if (le != cudaSuccess) { zg_set_error("attention layer launch failed: %s", cudaGetErrorString(le)); return false; }
    ZG_COUNT_LAUNCH();
    return true;
}

// `first` = any op of the batch (all share the work shape); `d_entries` = `count` consecutive table entries.
// Decode attention: how many CTAs share one head's kv range (1 = no split).  `count` ops x seq_q rows are in the launch.
uint32_t zg_attention_splits(const ZgOp& op, size_t k_buffer_elems, uint32_t count, int sm_count) {
    const auto& a = op.u.attention;
    if (!attn_fast_ok(op) || a.seq_q == 0 || a.seq_q > 8 || a.k_cs == 0) return 1;
    const size_t max_kv = (k_buffer_elems > a.k_off ? k_buffer_elems - a.k_off : 0) / a.k_cs;   // the cache cannot hold more rows
    uint32_t s = (uint32_t)((max_kv + 127) / 128);
    const uint32_t rows = count * a.seq_q;
    while (s > 1 && rows * s > (uint32_t)sm_count) s--;   // at most one wave of CTAs
    return s < 1 ? 1 : (s > 8 ? 8 : s);
}
size_t zg_attention_part_elems(const ZgOp& op, uint32_t count, uint32_t splits) {
    const uint32_t ni = op.u.attention.d_head <= 64 ? 2 : (op.u.attention.d_head <= 128 ? 4 : 8);
    return splits > 1 ? (size_t)count * op.u.attention.seq_q * splits * (ni * 32 + 2) : 0;
}

bool zg_launch_batch(const ZgOp& first, const ZgBatchEntry* d_entries, uint32_t count, const uint32_t* d_dyn, cudaStream_t st,
                     float* attn_part, uint32_t* attn_cnt, uint32_t attn_splits) {
    if (count == 0) return true;
    switch (first.tag) {
        case ZG_OP_SLICE_ASSIGN: {
            const auto& sa = first.u.slice_assign;
            if (sa.rows == 0 || sa.cols == 0) return true;
            launch_k(k_slice_assign, dim3(dim3(blocks_for(sa.rows * sa.cols, 256), count)), dim3(256), st, d_entries, d_dyn);
            break;
        }
        case ZG_OP_ROPE: {
            const auto& r = first.u.rope;
            if (r.half_d == 0 || r.seq_len == 0) return true;
            launch_k(k_rope, dim3(dim3(blocks_for(r.half_d * r.seq_len, 128), count)), dim3(128), st, d_entries);
            break;
        }
        case ZG_OP_ATTENTION: {
            const auto& a = first.u.attention;
            if (a.seq_q == 0 || a.d_head == 0) return true;
            if (attn_prefill_ok(first)) {
                const dim3 grid((a.seq_q + kPfBQ - 1) / kPfBQ, count);
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = grid; cfg.blockDim = dim3(kPfThreads); cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = attr; cfg.numAttrs = g_zg_pdl ? 1 : 0;
                cudaError_t le;
                if (a.d_head == 64) {
                    static bool once64 = false;
                    if (!once64) { cudaFuncSetAttribute(k_attention_prefill<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pf_smem_bytes<64>()); once64 = true; }
                    cfg.dynamicSmemBytes = pf_smem_bytes<64>();
                    le = cudaLaunchKernelEx(&cfg, k_attention_prefill<64>, d_entries, d_dyn);
                } else {
                    static bool once128 = false;
                    if (!once128) { cudaFuncSetAttribute(k_attention_prefill<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pf_smem_bytes<128>()); once128 = true; }
                    cfg.dynamicSmemBytes = pf_smem_bytes<128>();
                    le = cudaLaunchKernelEx(&cfg, k_attention_prefill<128>, d_entries, d_dyn);
                }
                if (le != cudaSuccess) { zg_set_error("prefill attention launch failed: %s", cudaGetErrorString(le)); return false; }
            } else if (attn_fast_ok(first)) {
                const uint32_t sp = (attn_part && attn_cnt && attn_splits > 1) ? attn_splits : 1u;
                const dim3 grid(a.seq_q, count, sp);
                if (a.d_head <= 64) launch_k(k_attention_fast<2>, dim3(grid), dim3(kAttnFastWarps * 32), st, d_entries, d_dyn, attn_part, attn_cnt, sp);
                else if (a.d_head <= 128) launch_k(k_attention_fast<4>, dim3(grid), dim3(kAttnFastWarps * 32), st, d_entries, d_dyn, attn_part, attn_cnt, sp);
                else launch_k(k_attention_fast<8>, dim3(grid), dim3(kAttnFastWarps * 32), st, d_entries, d_dyn, attn_part, attn_cnt, sp);
            } else launch_k(k_attention, dim3(dim3(a.seq_q, count)), dim3(kAttnWarps * 32), st, d_entries, d_dyn);
            break;
        }
        default: zg_set_error("internal: op kind %u is not batched", first.tag); return false;
    }
    ZG_COUNT_LAUNCH();
    return true;
}
