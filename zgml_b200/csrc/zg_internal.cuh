// Internal declarations shared by the CUDA translation units of libzgml_cuda.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <atomic>
#include <vector>

// The library is built with -fvisibility=hidden; only the C-ABI of the public header is exported.
#pragma GCC visibility push(default)
#include "../../include/zgml_cuda.h"
#pragma GCC visibility pop

void zg_set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_zg_launches;
extern std::atomic<uint64_t> g_zg_stream_launches;
#define ZG_COUNT_LAUNCH() (g_zg_launches.fetch_add(1, std::memory_order_relaxed))

#define ZG_CUDA_OK(expr)                                                                     \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            zg_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                         __LINE__);                                                          \
            return false;                                                                    \
        }                                                                                    \
    } while (0)

// ── Packed, GPU-resident quantized weight ─────────────────────────────────────
//
// Fast formats (block_size == 32 and N % 32 == 0) are stored as RECORDS.  One
// record covers one quant-block column group nb (ZG_TN = 32 output columns) x
// ZG_KR = 32 k-rows, and is laid out so that a warp's 128-bit global loads ARE the
// A fragments of mma.sync.m16n8k32 (weights as the 16-row operand: 16 output
// columns x 32 k), with no shared-memory staging and no unpack for int8:
//
//   q area   int8 formats: 2 column tiles (16 columns each) x 512 B.  Inside a column
//            tile lane L = 4*g + t owns 16 B = regs r = 0..3, bytes b = 0..3 with
//            column n = 16*ct + g + 8*(r & 1), row k = 4*t + b + 16*(r >> 1); the
//            byte is q itself (two's complement).
//            int4 format : 512 B.  Lane L owns 16 B = {w0, w1 of ct = 0, w0, w1 of
//            ct = 1}; byte b of w0 holds rows k = 4*t + b: low nibble = q + 8 of
//            column n = 16*ct + g, high nibble = q + 8 of column n = 16*ct + g + 8
//            (the GGUF Q4_0 biased nibbles); w1 is k + 16.
//   s area   [t 4][i 8] scales (f16 or f32): lane t's eight scales, i < 4: row
//            k = 4*t + i, i >= 4: row k = 16 + 4*t + (i - 4).
//
// Records are laid out [nb][k_chunk]: a warp walking K for one column group reads one
// contiguous run.  Rows >= K are zero-padded (q = 0, scale = 0).
// `smax[n_nb]` = max scale of each 32-column quant block over all k (used to choose
// the fixed-point exponent of the activation*scale products, see qgemv.cu).
#define ZG_TN 32
#define ZG_KR 32

struct ZgCudaQWeight {
    int fmt = 0;                 // ZG_QFMT_*
    size_t K = 0, N = 0, bs = 0; // logical [K, N], block size
    // fast formats
    uint32_t n_nb = 0, n_kc = 0, rec_bytes = 0, q_bytes = 0;
    uint8_t* recs = nullptr;
    float* smax = nullptr;
    // generic format: flat copies
    int8_t* g_data = nullptr;
    float* g_scales = nullptr;
    // W8A8 form (prepareTransposed, src/quant.zig:274-317): [N, K] int8 + [N, ceil(K / bs)] f32, plus the quantized
    // activation scratch of gemv (qw8a8.cu); null until zg_cuda_qweight_prepare_transposed
    int8_t* t_data = nullptr;
    float* t_scales = nullptr;
    int8_t* x_q = nullptr;
    float* x_s = nullptr;
    size_t device_bytes = 0;
};

static inline uint32_t zg_rec_q_bytes(int fmt) { return fmt == ZG_QFMT_I4_F16 ? 512u : 1024u; }
static inline uint32_t zg_rec_s_bytes(int fmt) { return fmt == ZG_QFMT_I8_F32 ? 128u : 64u; }

// split-K scratch: per-split partial sums of split column groups + one arrival counter per column group
struct ZgGemvWs {
    float* partials = nullptr;
    size_t partials_elems = 0;
    uint32_t* counters = nullptr;
    size_t counters_n = 0;
    float* gemm_scratch = nullptr;   // TF32-rounded activations of the prefill GEMM (one op at a time uses it)
    size_t gemm_scratch_elems = 0;
};

// One-shot all-reduce over NVLink peer memory (comm.cu, ops.cu k_allreduce_peer): every rank owns
// `slots[set][src_rank][max_n / 2]` 16-byte cells {x, epoch, y, epoch}; peers store their vector straight into them.
// Passed to the kernel by value.
constexpr int kZgMaxRanks = 8;
constexpr uint32_t kZgPeerSets = 2;
constexpr uint32_t kZgPeerCtas = 16;   // CTAs of one all-reduce (each owns a slice of the vector and its own counter)
struct ZgPeerComm {
    int rank = 0, world = 1;
    uint32_t max_n = 0;                 // floats per slot; 0 = peer path unavailable (NCCL is used instead)
    float* slots[kZgMaxRanks] = {};     // slot base of every rank (own entry = local pointer)
    uint32_t* seq = nullptr;            // local: [1] = timeout marker, [2 + c] = all-reduces CTA c has completed
    unsigned long long* cells = nullptr;   // local: [set][CTA] {partial sum of squares, epoch} of the fused all-reduce + norm kernel
};

struct ZgCudaCtx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool owns_stream = true;
    bool graph_mode = true;
    bool profiling = false;
    bool pdl = true; // programmatic dependent launch between consecutive qgemv kernels (ZG_CUDA_PDL=0 disables)
    int tune_s = 0, tune_p = 0, tune_u = 0, tune_g = 0, tune_smax = 0, tune_rows = 0; // ZG_GEMV_S / _P / _NS / _G overrides (kernel tuning only)
    ZgGemvWs ws; // split-K workspace for the direct zg_cuda_qmatmul_* calls
    int gemv_batch = 8;          // independent same-shape matvecs of a dependency level per launch (ZG_CUDA_GEMV_BATCH, 1 = off)
    int gemv_fuse = 0;           // evaluate norm (bit 0) / SiLU*up (bit 1) blocks inside the consuming matvecs' prologues (ZG_CUDA_GEMV_FUSE).
                                 // Off: measured SLOWER in-graph (the prologue's extra dependent L2 round trips cost what the removed kernel did)
    bool attn_split = true;      // decode attention: several CTAs per head over the kv range (ZG_CUDA_ATTN_SPLIT=0: one)
    int gemv_stream = 1, stream_min_chunks = 0, stream_early = 0, stream_ns = 0, stream_waves = 0, stream_chunks = 0, stream_align = 0;   // ZG_GEMV_STREAM / ZG_GEMV_STREAM_MIN (qgemv_stream.cu)
    bool ar_norm = false;        // sharded programs: all-reduce + the [add,] rmsnorm, gamma, mul block that consumes it in ONE launch (ZG_CUDA_AR_NORM=1; measured 2 % slower)
    bool gemv_cluster = true;    // k-split matvec: S in {2, 4, 8} splits of a column group form one cluster, partial sums meet in DSMEM (ZG_GEMV_CLUSTER=0: global scratch)
    bool gemv_pair = true;       // single-token programs: gate | up matvecs + SiLU * up chain in ONE launch (ZG_CUDA_GEMV_PAIR=0: off)
    bool attn_layer = true;      // single-token programs: rope + KV-cache stores + attention + concat of a layer in ONE launch (ZG_CUDA_ATTN_LAYER=0: off)
    bool decode_fused = false;   // single-token LLaMA layers run in the persistent fused decode kernel (decode.cu; ZG_CUDA_DECODE=1: on)
    bool fuse = true;            // evaluate the lowering's fixed op patterns (norm+gamma, SiLU*up, attention+store) in one pass
    size_t chain_max = 8200;     // small ops up to this many element visits join single-CTA chains (0 = off, ZG_CUDA_CHAIN)
    ZgPeerComm peer;             // NVLink peer-memory all-reduce state (max_n == 0: not available)
    void* peer_mem = nullptr;    // this rank's slots + counters (cudaMalloc, exported by cudaIpc)
    uint32_t* h_peer_err = nullptr;   // pinned copy of the peer-timeout word (ZgPeerComm::seq[1])
    void* peer_mapped[kZgMaxRanks] = {};   // cudaIpcOpenMemHandle mappings to close
    void* nccl_comm = nullptr;   // ncclComm_t (comm.cu), null unless zg_cuda_comm_init ran
    int rank = 0, world = 1;
    std::vector<cudaStream_t> branch; // extra capture streams: independent ops of a program become concurrent graph branches
};

// Grow-only (re)allocation; counters are zero-filled.  Never call between a graph
// capture and its replays: programs own a workspace sized once at compile time.
bool zg_gemv_ws_reserve(ZgGemvWs* ws, size_t partial_elems, size_t counters, cudaStream_t st, size_t gemm_scratch_elems = 0);
void zg_gemv_ws_free(ZgGemvWs* ws);

// qweight.cu
bool zg_qweight_dequant_to_device(ZgCudaCtx* ctx, const ZgCudaQWeight* w, float* d_out);   // dequantizeTo into K * N device floats (async)
ZgCudaQWeight* zg_qweight_from_device_flat(ZgCudaCtx* ctx, const int8_t* d_data, const float* d_scales,
                                           size_t K, size_t N, size_t bs, int fmt_hint);
// qgemv.cu
// Work split of one launch (<= 8 activation rows): P column groups per CTA, S k-splits per column group.
struct ZgGemvPlan {
    uint32_t grid = 0, threads = 256, P = 1, S = 1, mp = 1, G = 2, NS = 3, lcap = 1, xs_stride = 32, smem_bytes = 0;
};
ZgGemvPlan zg_qgemv_plan(const ZgCudaCtx* ctx, const ZgCudaQWeight* w, uint32_t M, uint32_t count = 1);
// dense_head.cu: bf16 copy of a dense matmul B operand (the tied LM head) with the argmax-safe exact recompute
struct ZgDenseHead { int format = 0; void* w16 = nullptr; float* err = nullptr; float* part_max = nullptr; float* glob = nullptr; uint32_t N = 0, K = 0, n_part = 0; };
ZgDenseHead* zg_dense_head_create(ZgCudaCtx* ctx, int format, const float* d_w, size_t w_off, size_t row_stride, uint32_t N, uint32_t K);
bool zg_dense_head_refresh(ZgCudaCtx* ctx, ZgDenseHead* h, const float* d_w, size_t w_off, size_t row_stride, cudaStream_t st);
void zg_dense_head_free(ZgDenseHead* h);
bool zg_dense_head_launch(ZgCudaCtx* ctx, const ZgDenseHead* h, const float* d_x, const float* d_w, size_t w_off, size_t row_stride,
                          float* d_dst, cudaStream_t st);
// qgemv_stream.cu: the large-launch form of the single-row matvec (one column group per warp, evenly sliced work)
struct ZgGemvStreamPlan { bool use = false; uint32_t grid = 0, GB = 0, nq = 0, TQ = 0, per = 0, Lq = 0, NS = 2, slots = 0, smem_bytes = 0; };
ZgGemvStreamPlan zg_qgemv_stream_plan(const ZgCudaCtx* ctx, const ZgCudaQWeight* w, uint32_t count);
bool zg_qgemv_stream_init(ZgCudaCtx* ctx);
bool zg_qgemv_stream_launch(ZgCudaCtx* ctx, const ZgGemvStreamPlan& pl, uint32_t count, const ZgCudaQWeight* const* w, const float* const* d_in,
                            float* const* d_out, const ZgGemvWs* ws, cudaStream_t st);
void zg_trace_set_gemv_stream(unsigned long long* d_buf);   // count: matvecs sharing the launch
void zg_qgemv_ws_need(const ZgCudaCtx* ctx, const ZgCudaQWeight* w, uint32_t M, size_t* partial_elems,
                      size_t* counters);
bool zg_qmatmul_launch(ZgCudaCtx* ctx, const ZgCudaQWeight* w, const float* d_in, float* d_out,
                       uint32_t M, uint32_t in_rs, uint32_t out_rs, const ZgGemvWs* ws, cudaStream_t st);
// How a decode matvec obtains its activation vector.  kind 0: read x.  Otherwise the small ops that PRODUCE x in the
// program (the lowering's norm block, or the SiLU chain times `up`) are evaluated in the matvec's own prologue, every
// CTA redundantly on its k-range, and the CTAs of column-group block 0 of the op flagged `write` also store the absorbed
// ops' output buffers (every DeviceOp result stays observable).  One kernel less per block on the decode critical path.
struct ZgDevStepC { uint32_t op, is_swapped; const float* sec; };
struct ZgGemvPrologue {
    uint32_t kind = 0;      // 1: [a + b ->] sum ; rmsnorm(sum) -> bare ; gamma -> gamma_rep ; bare * gamma_rep -> x
                            // 2: steps(a) -> mid ; mid * b -> x
    uint32_t write = 0;
    uint32_t n_steps = 0;
    float eps = 0.0f;
    const float* a = nullptr; const float* b = nullptr; const float* gamma = nullptr;
    float* o_sum = nullptr; float* o_mid = nullptr; float* o_grep = nullptr; float* o_x = nullptr;   // o_mid: bare (1) / mid (2)
    ZgDevStepC steps[6] = {};
};
// Epilogue of a gate | up matvec PAIR (M == 1): both matvecs run in one launch, every CTA computes the same column groups of
// both weights, and the activation chain that consumes them — fused_elementwise(mid = steps(gate)) ; mul(dst = mid * up), the
// SiLU * up of src/nn.zig:38-44 — is evaluated on the finished sums: one launch less on the decode critical path
// (src/backend/metal.zig:2533 fuses the same pair).  Every DeviceOp's buffer is still written.
struct ZgGemvEpilogue {
    uint32_t n_steps = 0, _pad = 0;
    ZgDevStepC steps[6] = {};
    uint32_t sec_kind[6] = {};     // 0: steps[i].sec[n] ; 1: the gate value ; 2: the up value
    float* o_mid = nullptr; float* o_dst = nullptr;
};
bool zg_qgemv_launch_pair(ZgCudaCtx* ctx, const ZgCudaQWeight* wa, const ZgCudaQWeight* wb, const float* d_in, float* d_out_a, float* d_out_b,
                          const ZgGemvWs* ws_a, const ZgGemvWs* ws_b, const ZgGemvEpilogue& epi, cudaStream_t st);
constexpr uint32_t kZgGemvBatch = 8;   // independent same-shape matvecs of one dependency level per launch
bool zg_qgemv_launch_batch(ZgCudaCtx* ctx, uint32_t count, const ZgCudaQWeight* const* w, const float* const* d_in,
                           float* const* d_out, uint32_t M, const uint32_t* in_rs, const uint32_t* out_rs,
                           const ZgGemvWs* ws, cudaStream_t st, const ZgGemvPrologue* pro = nullptr);   // ws, pro: one per op
bool zg_qgemv_init(ZgCudaCtx* ctx);
// qgemm.cu : M > 8 on tcgen05 tensor cores
bool zg_qgemm_init(ZgCudaCtx* ctx);
size_t zg_qgemm_scratch_elems(const ZgCudaQWeight* w, uint32_t M);
bool zg_qgemm_launch(ZgCudaCtx* ctx, const ZgCudaQWeight* w, const float* d_in, float* d_out, uint32_t M, uint32_t in_rs,
                     uint32_t out_rs, float* scratch, cudaStream_t st);

// In-kernel timeline (debug aid, ZG tracing off by default): block 0 / thread 0 of every program kernel stamps
// %globaltimer at entry (after the PDL wait) and exit into a device buffer: [0] = slot counter, then per slot
// {kind << 56 | t_entry_before_wait, t_after_wait, t_exit}.  Each translation unit keeps its own __constant__ pointer.
#define ZG_TRACE_DECL static __constant__ unsigned long long* c_zg_trace = nullptr;
#define ZG_TRACE_BEGIN(kind)                                                                                          \
    unsigned long long* zt_slot = nullptr;                                                                            \
    if (c_zg_trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {                                       \
        unsigned long long t;                                                                                         \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));                                                         \
        const unsigned long long s = atomicAdd(c_zg_trace, 1ull);                                                     \
        if (s < 16000) { zt_slot = c_zg_trace + 1 + 3 * s; zt_slot[0] = ((unsigned long long)(kind) << 56) | (t & 0xFFFFFFFFFFFFFFull); } \
    }
#define ZG_TRACE_MARK(idx)                                                                                            \
    if (zt_slot) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); zt_slot[idx] = t; }
void zg_trace_set_ops(unsigned long long* d_buf);
void zg_trace_set_gemv(unsigned long long* d_buf);
extern bool g_zg_pdl;   // ops.cu: launch op kernels with the programmatic-dependent-launch attribute
// ops.cu : one launcher per DeviceOp tag (buffers = device pointer table)
struct ZgDevStep { uint32_t op, is_swapped; const float* sec; };
// one op's parameters in the device table of the batched per-head kernels (rope, slice_assign, attention)
// dst2 != null (attention only): the output is ALSO stored at dst2[d2_off + r * d2_rs + qi * d2_cs] — the slice_assign that
// copies a head's output into the concatenated buffer (device_inference.zig:695-701), absorbed into the attention launch.
struct ZgBatchEntry { float* dst; const float* s0; const float* s1; const float* s2; const float* s3; uint32_t u[18]; float f; uint32_t dyn;
                      float* dst2; uint32_t d2_off, d2_rs, d2_cs, _pad; };
// one op of a chained run of small ops (ops.cu k_chain): executed by a single CTA in table order
struct ZgChainOp { float* dst; const float* s0; const float* s1; const ZgDevStep* steps; uint32_t u[18]; float f; uint32_t dyn, kind, sync, group, _pad; };
constexpr uint32_t kZgChainMaxOps = 160;
// macro ops of a chain (several consecutive DeviceOps evaluated in registers, every op's output still written):
//   FUSED_NORM: [elementwise add] -> rmsnorm -> repeat(gamma over rows) -> elementwise mul      (llama_transformer.zig:118-125)
//   EW_MUL:     fused_elementwise -> elementwise mul by another vector                          (SiLU(gate) * up, nn.zig:38-44)
constexpr uint32_t kZgChainFusedNorm = 100, kZgChainEwMul = 101;
struct ZgNormMacro { const float* a; const float* b; float* sum; float* bare; const float* gamma; float* gamma_rep; float* norm; uint32_t rows, cols; float eps; };
struct ZgEwMulMacro { const float* src; float* mid; const float* other; float* dst; const ZgDevStep* steps; uint32_t n_steps, n; };
bool zg_fill_chain_norm(const ZgNormMacro& m, bool sync, ZgChainOp* c);
bool zg_fill_chain_ewmul(const ZgEwMulMacro& m, bool sync, ZgChainOp* c);
bool zg_launch_ewmul(const ZgEwMulMacro& m, cudaStream_t st);
bool zg_launch_peer_allreduce_norm(float* buf, size_t n, const ZgPeerComm& pc, const ZgNormMacro& m, cudaStream_t st);
bool zg_launch_norm_macro(const ZgNormMacro& m, cudaStream_t st);   // rows longer than 4096: one 1024-thread CTA per row   // the same pair as one multi-CTA launch   // ops per chain launch (the table lives in shared memory)
size_t zg_chain_work(const ZgOp& op);
bool zg_fill_chain_op(const ZgOp& op, float* const* bufs, uint32_t op_index, const ZgDevStep* d_steps, bool sync, ZgChainOp* c);
bool zg_launch_chain(const ZgChainOp* d_ops, uint32_t count, const uint32_t* d_dyn, cudaStream_t st);
bool zg_launch_peer_allreduce(float* buf, size_t n, const ZgPeerComm& pc, cudaStream_t st);
bool zg_peer_allreduce_ok(const ZgCudaCtx* ctx, size_t n);
void zg_peer_check_enqueue(ZgCudaCtx* ctx, cudaStream_t st);   // comm.cu: copy the peer-timeout word behind the queued work ...
bool zg_peer_check_result(ZgCudaCtx* ctx);                      // ... and, after the sync: false (+ error string) when a peer all-reduce gave up   // comm.cu: the peer path can take an n-float all-reduce
bool zg_op_is_batched(uint32_t tag);
uint64_t zg_batch_signature(const ZgOp& op);
bool zg_fill_batch_entry(const ZgOp& op, float* const* bufs, uint32_t op_index, ZgBatchEntry* e);
bool zg_launch_batch(const ZgOp& first, const ZgBatchEntry* d_entries, uint32_t count, const uint32_t* d_dyn, cudaStream_t st,
                     float* attn_part = nullptr, uint32_t* attn_cnt = nullptr, uint32_t attn_splits = 1);
uint32_t zg_attention_splits(const ZgOp& op, size_t k_buffer_elems, uint32_t count, int sm_count);
size_t zg_attention_part_elems(const ZgOp& op, uint32_t count, uint32_t splits);
bool zg_launch_op(ZgCudaCtx* ctx, const ZgOp& op, float* const* bufs, const uint32_t* d_dyn,
                  uint32_t op_index, const ZgDevStep* d_steps, cudaStream_t st);

// ── decode.cu : whole LLaMA layers of a single-token (T == 1) program in ONE persistent kernel ─────────────────────
// The per-layer op pattern of the lowering (src/models/llama_transformer.zig:192-253 through src/device_inference.zig) is
// recognised at compile time (backend.cu match_decode_layers) and executed as five phases per layer separated by grid
// barriers: [norm + q|k|v matvecs] [rope + KV store + split-KV attention] [merge + o matvec] [residual + norm + gate|up
// matvecs] [SiLU*up + down matvec] (+ an NVLink peer all-reduce phase after o / down when sharded).  Every DeviceOp's
// output buffer is still written.  Weights stream through per-warp TMA rings that run AHEAD across phase and layer
// boundaries (weights are immutable), so barrier and prologue latencies overlap with HBM traffic.
constexpr uint32_t kZgDecMaxHeads = 64;   // query heads / KV heads of one layer (per rank)
constexpr uint32_t kZgDecMaxMv = 3;       // matvecs per phase (q|k|v, gate|up)
constexpr uint32_t kZgDecMaxSteps = 8;    // fused_elementwise steps of the activation chain
constexpr uint32_t kZgDecMaxD = 8192;     // floats a CTA stages per phase (norm phases: the whole d_model vector)
struct ZgDecVec {           // a vector produced by a matvec phase: complete in `full` (S == 0) or S partial sums part[s * n + i]
    float* full; const float* part; uint32_t S, n;
};
struct ZgDecMv {
    const uint8_t* recs; const float* smax;
    float* out;             // S == 1: results; S > 1: part[split * N + n] (the consumer phase sums the splits and stores `out`)
    float* part;
    uint32_t n_nb, first_item, N, _pad;
};
struct ZgDecPhase {
    ZgDecMv mv[kZgDecMaxMv];       // absent entries: first_item = UINT32_MAX
    uint32_t n_mv, fmt, n_kc, K, S, n_slots, n_items, rec_bytes;
    uint32_t lS, _pad;             // S = 1 << lS k-splits (powers of two keep the device-side geometry to shifts)
};
struct ZgDecHead { float* q_rot; float* attn_out; uint32_t q_src, k_off, v_off, kv, buf_off, dyn; };
struct ZgDecKv { float* k_rot; uint32_t k_src, v_src, k_dyn, v_dyn, k_base, v_base; };
struct ZgDecStep { uint32_t op, is_swapped, sec_kind, _pad; const float* sec; };   // sec_kind 0: sec[i]; 1: the gate value; 2: the up value
struct ZgDecLayer {
    // phase 1: x = a (+ b -> sum) ; rmsnorm -> bare ; gamma -> grep ; bare * grep -> norm ; q|k|v matvecs
    const float* x1_a; ZgDecVec x1_b; float* x1_sum;
    const float* gamma1; float* bare1; float* grep1; float* norm1; float eps1; uint32_t D;
    // phase 2: rope(q), rope(k) + KV store, attention over the cache -> partial states
    ZgDecVec q, k, v;
    const float* cs; const float* mask; float* k_cache; float* v_cache; float* attn_buf;
    uint32_t n_heads, n_kv, d_head, has_mask, mask_off, mask_rs, head0, kv0, k_cs, v_cs;
    float scale;
    // phase 3: merge -> attn_out / attn_buf ; o matvec (+ all-reduce)
    ZgDecVec o_local, o;    // as the matvec leaves it / as phase 4 reads it (complete after an all-reduce)
    uint32_t ar_o, ar_down;
    // phase 4: x = a + o -> sum ; norm ; gate|up matvecs
    const float* x2_a; float* x2_sum; const float* gamma2; float* bare2; float* grep2; float* norm2; float eps2;
    // phase 5: act = steps(gate) -> silu ; silu * up -> hidden ; down matvec (+ all-reduce)
    ZgDecVec gate, up; float* silu; float* hidden; uint32_t n_steps, F;
    uint32_t act_silu, _pad2;   // the steps are exactly neg, exp, add(ext), recip, mul(gate): evaluated inline
    ZgDecStep steps[kZgDecMaxSteps];
    ZgDecVec down_local, down;
};
constexpr uint32_t kZgDecMaxItems = 32;   // column groups one CTA may own in one matvec phase
// Per-layer descriptor block in device memory, copied into shared memory one layer ahead of its use:
//   [ZgDecLayer][ZgDecPhase x 4 : qkv, o, gate|up, down][ZgDecHead x cap_heads][ZgDecKv x cap_kv]
struct ZgDecodePlan {       // passed to the kernel by value
    const uint8_t* blocks; uint32_t blk_bytes, cap_heads, cap_kv, n_layers;
    float* attn_part;       // [head][max_splits][2 + part_dh]
    uint32_t* sync;         // [0] grid barrier counter, [32] exit counter, [64] sticky error flag
    const uint32_t* dyn;
    uint32_t max_splits, part_dh, grid, _pad;
    ZgPeerComm pc;
};
struct ZgDecodeHost {       // owned by a compiled program
    ZgDecodePlan plan = {};
    void* d_blocks = nullptr;
    float* d_part = nullptr; float* d_attn_part = nullptr; uint32_t* d_sync = nullptr;
    uint32_t* h_err = nullptr;   // pinned copy of the error flag, read after every execute
    bool valid = false;
};
// ── ops.cu k_attention_layer: the whole attention block of one layer of a single-token program in ONE launch ──
// rope(k) -> K-cache store, V-cache store, rope(q), split-KV attention, merge, copy into the concatenated buffer — 3 * (n_kv +
// n_heads) DeviceOps (src/device_inference.zig lowering of llama_transformer.zig:192-253).  Every op's buffer is still written.
struct ZgAttnBlock {
    const float* q_proj; const float* k_proj; const float* v_proj; const float* cs; const float* mask;
    float* k_cache; float* v_cache; float* attn_buf;
    uint32_t n_heads, n_kv, d_head, has_mask, mask_off, mask_rs, k_cs, v_cs;
    float scale; uint32_t min_pos;   // KV positions per split at least (64)
    ZgDecHead heads[kZgDecMaxHeads];
    ZgDecKv kvs[kZgDecMaxHeads];
};
bool zg_launch_attention_layer(const ZgAttnBlock* d_blk, uint32_t n_heads, uint32_t d_head, uint32_t max_splits, const uint32_t* d_dyn, float* part,
                               uint32_t* cnt, cudaStream_t st);   // part: [n_heads][max_splits][2 + pad32(d_head)], cnt: [n_heads] zeroed
bool zg_decode_init(ZgCudaCtx* ctx);
uint32_t zg_decode_grid(const ZgCudaCtx* ctx);
uint32_t zg_decode_block_bytes();
void zg_decode_block_fill(uint8_t* block, const ZgDecLayer& ly, const ZgDecPhase* ph4, const ZgDecHead* heads, uint32_t n_heads,
                          const ZgDecKv* kvs, uint32_t n_kv);
bool zg_decode_launch(ZgCudaCtx* ctx, const ZgDecodeHost& d, cudaStream_t st);
void zg_decode_free(ZgDecodeHost* d);
void zg_trace_set_decode(unsigned long long* d_buf);

// qkv.cu : quantized KV cache ops as program ops (zg_cuda_program_quantize_kv)
struct ZgKvqStore { const float* src; int8_t* q; float* s; uint32_t src_cs, n_write, dyn_idx, d_head, bs, bpc; };
struct ZgKvqAttn {
    float* dst; size_t dst_cs;
    const float* q; size_t q_cs;
    uint32_t d_head, seq_kv, bs, nb;
    const int8_t* k_q; const float* k_s; size_t k_col_start;
    const int8_t* v_q; const float* v_s; size_t v_col_start;
    const float* mask; size_t mask_rs, mask_cs;
    float scale;
    int int8_query;
    uint32_t dyn_idx, _pad;
};
bool zg_kvq_cache_arrays(const ZgCudaKVCache* c, int8_t** q, float** s, uint32_t* d_head, uint32_t* bs, uint32_t* bpc, size_t* n_cols);
bool zg_kvq_launch_stores(const ZgKvqStore* d_tab, uint32_t count, uint32_t max_warps, const uint32_t* d_dyn, cudaStream_t st);
bool zg_kvq_launch_attention(const ZgKvqAttn* d_tab, uint32_t count, uint32_t seq_q, const uint32_t* d_dyn, float* part, uint32_t* cnt,
                             uint32_t splits_max, cudaStream_t st);   // part: [count * seq_q][splits_max][2 + d_head], cnt: [count * seq_q] zeroed

// comm.cu : NCCL through dlopen (no link-time dependency)
bool zg_comm_allreduce(ZgCudaCtx* ctx, float* buf, size_t n, cudaStream_t st);
bool zg_comm_allgather(ZgCudaCtx* ctx, const float* src, float* dst, size_t n_per_rank, cudaStream_t st);
