// C-ABI entry points: backend context, compiled DevicePrograms (device-resident
// buffers + packed weights + CUDA-graph replay) and the direct quantized-matmul
// calls.  Mirrors the Backend vtable contract of src/backend.zig:330-352 the way
// src/backend/cpu.zig:55-147 implements it for the CPU.
#include "zg_internal.cuh"

#include <algorithm>
#include <initializer_list>
#include <map>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <string>

std::atomic<uint64_t> g_zg_launches{0};
std::atomic<uint64_t> g_zg_stream_launches{0};   // launches of the streamed matvec kernel (qgemv_stream.cu)
static char g_err[1024] = "";

void zg_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* zg_cuda_last_error(void) { return g_err; }
extern "C" uint64_t zg_cuda_launch_count(void) { return g_zg_launches.load(); }

bool zg_gemv_ws_reserve(ZgGemvWs* ws, size_t partial_elems, size_t counters, cudaStream_t st, size_t gemm_scratch_elems) {
    if (gemm_scratch_elems > ws->gemm_scratch_elems) {
        if (ws->gemm_scratch) { cudaStreamSynchronize(st); cudaFree(ws->gemm_scratch); ws->gemm_scratch = nullptr; }
        ZG_CUDA_OK(cudaMalloc(&ws->gemm_scratch, gemm_scratch_elems * sizeof(float)));
        ws->gemm_scratch_elems = gemm_scratch_elems;
    }
    if (partial_elems > ws->partials_elems) {
        if (ws->partials) { cudaStreamSynchronize(st); cudaFree(ws->partials); ws->partials = nullptr; }
        ZG_CUDA_OK(cudaMalloc(&ws->partials, partial_elems * sizeof(float)));
        ws->partials_elems = partial_elems;
    }
    if (counters > ws->counters_n) {
        if (ws->counters) { cudaStreamSynchronize(st); cudaFree(ws->counters); ws->counters = nullptr; }
        ZG_CUDA_OK(cudaMalloc(&ws->counters, counters * sizeof(uint32_t)));
        ZG_CUDA_OK(cudaMemsetAsync(ws->counters, 0, counters * sizeof(uint32_t), st));
        ZG_CUDA_OK(cudaStreamSynchronize(st));
        ws->counters_n = counters;
    }
    return true;
}
void zg_gemv_ws_free(ZgGemvWs* ws) {
    cudaFree(ws->partials); cudaFree(ws->counters); cudaFree(ws->gemm_scratch);
    *ws = ZgGemvWs();
}

// stride != 0: only the windows [lo + j * stride, + width) inside [lo, hi) are touched (a head's rows of a [T, D] matrix)
struct ZgRange { uint32_t buf; size_t lo, hi; bool write; uint32_t dyn; uint32_t stride = 0, width = 0; };
struct ZgCudaProgram {
    ZgCudaCtx* ctx = nullptr;
    std::vector<ZgOp> ops;
    std::vector<std::vector<ZgFusedEwStep>> steps; // owned copies, per op (empty unless fused)
    std::vector<uint32_t> step_off;                // offset into d_steps per op
    std::vector<float*> buffers;
    std::vector<size_t> buffer_elems;
    std::vector<ZgCudaQWeight*> qweights;
    std::vector<bool> qweight_owned;   // false: resident weight borrowed from the caller (ZG_QWEIGHT_RESIDENT)
    ZgDevStep* d_steps = nullptr;
    uint32_t* d_dyn = nullptr;
    uint32_t* h_dyn = nullptr; // pinned
    bool dyn_dirty = true;     // h_dyn differs from d_dyn
    ZgGemvWs ws;                       // split-K scratch: one private slice per qmatmul op (ops may run concurrently)
    std::vector<size_t> ws_part_off, ws_cnt_off;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool graph_valid = false;
    uint64_t graph_kernels = 0; // kernel nodes captured in `graph` (added to the launch counter per replay)
    uint64_t graph_streamed = 0; // of those: streamed matvec launches
    ZgProfile profile;
    std::vector<cudaEvent_t> prof_events;
    std::vector<cudaEvent_t> dep_events; // capture-time fork/join markers of the concurrent graph branches
    // launch schedule: ops sorted by dependency level, same-level per-head ops (rope / slice_assign / attention)
    // of equal shape merged into one batched launch
    // ... and runs of small ops (norms, broadcasts, residual adds, rope, cache stores, SiLU chains) spanning
    // consecutive levels chained into one single-CTA launch (ops.cu k_chain)
    struct Unit {
        std::vector<uint32_t> ops;         // every DeviceOp the launch executes (its dependency footprint)
        std::vector<uint32_t> entry_ops;   // batched: the ops that own a table entry (absorbed slice_assigns do not)
        std::map<uint32_t, uint32_t> store_of;   // batched attention op -> the slice_assign absorbed into it
        uint32_t first_entry = 0, n_entries = 0;
        std::vector<ZgGemvPrologue> pros;  // matvec batch: how each op obtains its activations
        std::vector<ZgRange> ranges;       // dependency footprint when it differs from the union of the ops' own ranges
        bool batched = false, chain = false, ewmul = false, gemv_batch = false, norm = false, decode = false;
        uint32_t kvq = 0, kvq_max_warps = 0, kvq_seq_q = 0, kvq_splits = 1; size_t kvq_part_off = 0, kvq_cnt_off = 0;
        bool ar_norm = false;   // peer all-reduce + the norm block that consumes it (ops.cu k_allreduce_norm); nm holds the block
        bool gemv_pair = false; ZgGemvEpilogue epi;   // gate | up matvec pair + activation epilogue (qgemv.cu qgemv_pair_kernel)
        int attn_blk = -1; uint32_t ab_splits = 1; size_t ab_part_off = 0, ab_cnt_off = 0;   // fused attention block of one layer   // 1: batch of cache stores, 2: batch of cache-backed attentions
        ZgNormMacro nm = {};
        uint32_t attn_splits = 1; size_t attn_part_off = 0, attn_cnt_off = 0;   // split-KV decode attention scratch (per unit)
        ZgEwMulMacro em = {};
    };
    std::vector<Unit> units;
    std::vector<uint32_t> entry_of_op;   // index into d_batch for batched op kinds
    std::vector<uint32_t> single_entry;  // index into d_batch of a batched-kind op's own plain entry (per-op launches)
    ZgBatchEntry* d_batch = nullptr;
    ZgChainOp* d_chain = nullptr;
    // input staging: the per-step inputs of a decode program are many small host buffers (one RoPE leaf per layer);
    // they are packed into ONE pinned buffer, copied once and scattered by a kernel instead of one pageable copy each
    uint8_t* h_in_stage = nullptr; uint8_t* d_in_stage = nullptr; size_t in_stage_bytes = 0;
    struct InSeg { float* dst; uint32_t src_off, words; };
    InSeg* d_in_tab = nullptr; size_t in_tab_cap = 0;
    std::vector<InSeg> in_tab_host;   // what d_in_tab holds
    float* d_attn_part = nullptr;      // split-KV partial states, one slice per attention unit
    uint32_t* d_attn_cnt = nullptr;    // arrival counters (self re-arming)
    bool uniform_pos = true;   // every patched slice_assign sits at the same position (checked per refresh)
    // quantized KV cache mode (zg_cuda_program_quantize_kv; LlamaInferenceSession.quantizeKV, src/llama_inference.zig:648-679):
    // Q8 caches stand in for the f32 cache buffers; the patched slice_assigns into them run storeColumn, the attention ops
    // over them attentionQuantized (src/llama_inference.zig:336-377)
    bool kvq = false; size_t kvq_bs = 32; int kvq_int8 = 0;
    std::map<uint32_t, ZgCudaKVCache*> kv_caches;   // program buffer -> the cache standing in for it
    std::map<uint32_t, ZgDenseHead*> dense;         // matmul op index -> bf16 copy of its B operand (zg_cuda_program_promote_dense)
    std::vector<char> kvq_role;                     // per op: 0 none, 1 cache store, 2 attention over the caches
    std::vector<uint32_t> kvq_entry;                // per op: its entry in d_kvq_store / d_kvq_attn
    ZgKvqStore* d_kvq_store = nullptr; ZgKvqAttn* d_kvq_attn = nullptr;
    float* d_kvq_part = nullptr; uint32_t* d_kvq_cnt = nullptr;   // split-KV scratch of the cache-backed attention units
    std::vector<ZgAttnBlock> attn_blocks;            // single-token attention blocks fused into one launch each (ops.cu k_attention_layer)
    ZgAttnBlock* d_attn_blocks = nullptr; float* d_attn_blk_part = nullptr; uint32_t* d_attn_blk_cnt = nullptr;
    ZgDecodeHost dec;          // fused decode kernel (decode.cu) when the program's layers match the single-token LLaMA pattern
    uint32_t dec_first = 0, dec_count = 0;   // the ops it covers
};

// ── context ──────────────────────────────────────────────────────────────────
extern "C" ZgCudaCtx* zg_cuda_create(int device_ordinal) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { zg_set_error("no CUDA device available"); return nullptr; }
    if (device_ordinal < 0 || device_ordinal >= n) { zg_set_error("device ordinal %d out of range (%d devices)", device_ordinal, n); return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_ordinal) != cudaSuccess) { zg_set_error("cudaGetDeviceProperties failed"); return nullptr; }
    if (prop.major != 10) { // kernels are built for sm_100a only; no other code path exists
        zg_set_error("device %d is sm_%d%d; this backend is built for sm_100a (B200) only", device_ordinal, prop.major, prop.minor);
        return nullptr;
    }
    if (cudaSetDevice(device_ordinal) != cudaSuccess) { zg_set_error("cudaSetDevice failed"); return nullptr; }
    ZgCudaCtx* ctx = new ZgCudaCtx();
    ctx->device = device_ordinal;
    ctx->sm_count = prop.multiProcessorCount;
    if (const char* e = getenv("ZG_CUDA_PDL")) ctx->pdl = (e[0] != '0');
    g_zg_pdl = ctx->pdl;
    if (const char* e = getenv("ZG_CUDA_GEMV_BATCH")) { ctx->gemv_batch = atoi(e); if (ctx->gemv_batch < 1) ctx->gemv_batch = 1; if (ctx->gemv_batch > (int)kZgGemvBatch) ctx->gemv_batch = kZgGemvBatch; }
    if (const char* e = getenv("ZG_CUDA_GEMV_FUSE")) ctx->gemv_fuse = atoi(e);   // bit 0: norm block, bit 1: SiLU*up pair
    if (const char* e = getenv("ZG_CUDA_ATTN_SPLIT")) ctx->attn_split = (e[0] != '0');
    if (const char* e = getenv("ZG_CUDA_FUSE")) ctx->fuse = (e[0] != '0');   // 0: no macro patterns (one chain / batch entry per DeviceOp)
    if (const char* e = getenv("ZG_CUDA_CHAIN")) ctx->chain_max = (size_t)atol(e);   // 0: one launch per small op
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        zg_set_error("cudaStreamCreate failed"); delete ctx; return nullptr;
    }
    if (const char* e = getenv("ZG_CUDA_AR_NORM")) ctx->ar_norm = (e[0] != '0');       // 1: the all-reduce and the norm block behind it as ONE launch (measured slower, off by default)
    if (const char* e = getenv("ZG_CUDA_GEMV_PAIR")) ctx->gemv_pair = (e[0] != '0');   // 0: gate | up as a plain batch, the activation chain as its own launch
    if (const char* e = getenv("ZG_CUDA_ATTN_LAYER")) ctx->attn_layer = (e[0] != '0');   // 0: rope / cache stores / attention as separate launches
    if (const char* e = getenv("ZG_CUDA_DECODE")) ctx->decode_fused = (e[0] != '0');   // 0: never use the fused decode kernel
    if (!zg_qgemv_init(ctx) || !zg_qgemm_init(ctx) || !zg_decode_init(ctx)) { cudaStreamDestroy(ctx->stream); delete ctx; return nullptr; }
    int n_branch = 7; // capture streams for independent ops of a program (ZG_CUDA_BRANCH=0: strictly serial graphs)
    if (const char* e = getenv("ZG_CUDA_BRANCH")) n_branch = atoi(e);
    for (int i = 0; i < n_branch && i < 31; i++) {
        cudaStream_t b;
        if (cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking) != cudaSuccess) break;
        ctx->branch.push_back(b);
    }
    return ctx;
}

extern "C" void zg_cuda_destroy(ZgCudaCtx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    zg_cuda_comm_destroy(ctx);
    zg_gemv_ws_free(&ctx->ws);
    for (cudaStream_t b : ctx->branch) cudaStreamDestroy(b);
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" void zg_cuda_capabilities(ZgCapabilities* c) {
    // modelled on Capabilities.reference_cpu (src/backend.zig:60-70) minus host-visible memory
    memset(c, 0, sizeof(*c));
    c->compiled_programs = 1; c->host_visible_program_memory = 0;
    c->dense_matmul_f32 = 1; c->qmatmul = 1; c->fused_elementwise = 1; c->max_fused_elementwise_steps = 0;
    c->dynamic_program_refresh = 1; c->prefill_attention = 1; c->decode_attention = 1; c->quantized_kv = 1;
    c->attention_supported = 1; c->attention_max_seq_kv = 0; c->attention_max_d_head = 512;
}

extern "C" int zg_cuda_dense_matmul_f32(ZgCudaCtx*, float*, const float*, const float*, const ZgMatMulGeometry*) {
    return 0; // decline: host tensors are not device resident
}

extern "C" void zg_cuda_set_stream(ZgCudaCtx* ctx, void* s) {
    if (!ctx) return;
    cudaStreamSynchronize(ctx->stream);
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)s;
    ctx->owns_stream = false;
}
extern "C" void zg_cuda_sync(ZgCudaCtx* ctx) {
    if (!ctx) return;
    zg_peer_check_enqueue(ctx, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    zg_peer_check_result(ctx);
}
extern "C" void zg_cuda_set_graph_mode(ZgCudaCtx* ctx, int e) { if (ctx) ctx->graph_mode = e != 0; }
extern "C" void zg_cuda_set_profiling(ZgCudaCtx* ctx, int e) { if (ctx) ctx->profiling = e != 0; }

// ── program ──────────────────────────────────────────────────────────────────
static bool op_buffers_valid(const ZgOp& op, size_t nb) { // src/backend.zig:303-325
    auto ok = [&](uint32_t i) { return (size_t)i < nb; };
    switch (op.tag) {
        case ZG_OP_ELEMENTWISE: return ok(op.u.elementwise.dst) && ok(op.u.elementwise.src0) && ok(op.u.elementwise.src1);
        case ZG_OP_MATMUL: return ok(op.u.matmul.dst) && ok(op.u.matmul.a) && ok(op.u.matmul.b);
        case ZG_OP_QMATMUL: return ok(op.u.qmatmul.dst) && ok(op.u.qmatmul.input);
        case ZG_OP_SOFTMAX: return ok(op.u.softmax.dst) && ok(op.u.softmax.src);
        case ZG_OP_LAYERNORM: return ok(op.u.layernorm.dst) && ok(op.u.layernorm.src);
        case ZG_OP_RMSNORM: return ok(op.u.rmsnorm.dst) && ok(op.u.rmsnorm.src);
        case ZG_OP_REDUCE: return ok(op.u.reduce.dst) && ok(op.u.reduce.src);
        case ZG_OP_REPEAT: return ok(op.u.repeat.dst) && ok(op.u.repeat.src);
        case ZG_OP_SLICE_ASSIGN: return ok(op.u.slice_assign.dst) && ok(op.u.slice_assign.src);
        case ZG_OP_ROPE: return ok(op.u.rope.dst) && ok(op.u.rope.src) && ok(op.u.rope.cos_sin);
        case ZG_OP_ATTENTION:
            return ok(op.u.attention.dst) && ok(op.u.attention.q) && ok(op.u.attention.k) && ok(op.u.attention.v) && ok(op.u.attention.mask);
        case ZG_OP_FUSED_ELEMENTWISE: {
            if (!ok(op.u.fused_elementwise.dst) || !ok(op.u.fused_elementwise.src)) return false;
            for (size_t s = 0; s < op.u.fused_elementwise.n_steps; s++) {
                const ZgFusedEwStep& st = op.u.fused_elementwise.steps[s];
                if ((st.op == ZG_EW_ADD || st.op == ZG_EW_MUL) && !ok(st.secondary_buf)) return false;
            }
            return true;
        }
        case ZG_OP_ALLREDUCE: return ok(op.u.allreduce.buf);
        case ZG_OP_ALLGATHER: return ok(op.u.allgather.dst) && ok(op.u.allgather.src);
        default: return false;
    }
}

static uint32_t op_dyn_value(const ZgOp& op) {
    if (op.tag == ZG_OP_SLICE_ASSIGN) return op.u.slice_assign.dst_offset;
    if (op.tag == ZG_OP_ATTENTION) return op.u.attention.seq_kv;
    return 0;
}

static void free_program(ZgCudaProgram* p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    if (p->exec) cudaGraphExecDestroy(p->exec);
    if (p->graph) cudaGraphDestroy(p->graph);
    for (float* b : p->buffers) cudaFree(b);
    for (size_t i = 0; i < p->qweights.size(); i++)
        if (p->qweight_owned[i]) zg_cuda_qweight_free(p->ctx, p->qweights[i]);
    cudaFree(p->d_steps); cudaFree(p->d_dyn); cudaFree(p->d_batch); cudaFree(p->d_chain); cudaFree(p->d_attn_part); cudaFree(p->d_attn_cnt);
    cudaFree(p->d_in_stage); cudaFree(p->d_in_tab);
    if (p->h_in_stage) cudaFreeHost(p->h_in_stage);
    if (p->h_dyn) cudaFreeHost(p->h_dyn);
    zg_gemv_ws_free(&p->ws);
    zg_decode_free(&p->dec);
    for (auto& kv : p->kv_caches) zg_cuda_kvcache_free(p->ctx, kv.second);
    for (auto& d : p->dense) zg_dense_head_free(d.second);
    cudaFree(p->d_kvq_store); cudaFree(p->d_kvq_attn); cudaFree(p->d_kvq_part); cudaFree(p->d_kvq_cnt);
    cudaFree(p->d_attn_blocks); cudaFree(p->d_attn_blk_part); cudaFree(p->d_attn_blk_cnt);
    for (cudaEvent_t e : p->prof_events) cudaEventDestroy(e);
    for (cudaEvent_t e : p->dep_events) cudaEventDestroy(e);
    delete p;
}

// (Re)build the device-side fused-step table from p->ops / p->steps.
static bool upload_steps(ZgCudaProgram* p) {
    std::vector<ZgDevStep> flat;
    p->step_off.assign(p->ops.size(), 0);
    for (size_t i = 0; i < p->ops.size(); i++) {
        if (p->ops[i].tag != ZG_OP_FUSED_ELEMENTWISE) continue;
        p->step_off[i] = (uint32_t)flat.size();
        for (const ZgFusedEwStep& st : p->steps[i]) {
            ZgDevStep d;
            d.op = st.op; d.is_swapped = st.is_swapped; d.sec = nullptr;
            if (st.op == ZG_EW_ADD || st.op == ZG_EW_MUL) d.sec = p->buffers[st.secondary_buf] + st.secondary_offset;
            flat.push_back(d);
        }
    }
    cudaFree(p->d_steps); p->d_steps = nullptr;
    if (!flat.empty()) {
        ZG_CUDA_OK(cudaMalloc(&p->d_steps, flat.size() * sizeof(ZgDevStep)));
        ZG_CUDA_OK(cudaMemcpy(p->d_steps, flat.data(), flat.size() * sizeof(ZgDevStep), cudaMemcpyHostToDevice));
    }
    return true;
}

static bool adopt_ops(ZgCudaProgram* p, const ZgOp* ops, size_t n_ops) {
    p->ops.assign(ops, ops + n_ops);
    p->steps.assign(n_ops, {});
    for (size_t i = 0; i < n_ops; i++) {
        if (ops[i].tag == ZG_OP_FUSED_ELEMENTWISE) {
            const auto& f = ops[i].u.fused_elementwise;
            p->steps[i].assign(f.steps, f.steps + f.n_steps);
            p->ops[i].u.fused_elementwise.steps = p->steps[i].data();
        }
    }
    return true;
}

static bool validate_ops(const ZgCudaProgram* p, const ZgOp* ops, size_t n_ops) {
    for (size_t i = 0; i < n_ops; i++) {
        if (!op_buffers_valid(ops[i], p->buffers.size())) { zg_set_error("op %zu: invalid tag or buffer index", i); return false; }
        if (ops[i].tag == ZG_OP_QMATMUL) { // src/backend.zig:284-292
            const auto& q = ops[i].u.qmatmul;
            if ((size_t)q.weight_idx >= p->qweights.size()) { zg_set_error("op %zu: weight_idx out of range", i); return false; }
            const ZgCudaQWeight* w = p->qweights[q.weight_idx];
            if (w->K != q.K || w->N != q.N) { zg_set_error("op %zu: qweight is [%zu,%zu], op wants [%u,%u]", i, w->K, w->N, q.K, q.N); return false; }
        }
        if (ops[i].tag == ZG_OP_ATTENTION && ops[i].u.attention.d_head > 512) { zg_set_error("op %zu: d_head > 512", i); return false; }
    }
    return true;
}

static bool build_schedule(ZgCudaProgram* p);

static bool reserve_workspace(ZgCudaProgram* p) {
    size_t pe = 0, nc = 0, gs = 0;
    p->ws_part_off.assign(p->ops.size(), 0);
    p->ws_cnt_off.assign(p->ops.size(), 0);
    for (size_t i = 0; i < p->ops.size(); i++) {
        const ZgOp& op = p->ops[i];
        if (op.tag != ZG_OP_QMATMUL) continue;
        size_t a = 0, b = 0;
        zg_qgemv_ws_need(p->ctx, p->qweights[op.u.qmatmul.weight_idx], op.u.qmatmul.M, &a, &b);
        p->ws_part_off[i] = pe; p->ws_cnt_off[i] = nc;
        pe += a; nc += b;
        const size_t g = zg_qgemm_scratch_elems(p->qweights[op.u.qmatmul.weight_idx], op.u.qmatmul.M);
        if (g > gs) gs = g;
    }
    return zg_gemv_ws_reserve(&p->ws, pe, nc, p->ctx->stream, gs);
}

// ── op dependencies (element ranges per buffer) for concurrent graph branches ─────────────────
// dyn != 0: the range is shifted at run time by pos * dyn (slice_assign with patch_stride, i.e. KV-cache writes);
// all such ops of a program share one pos (src/device_inference.zig:240-247), checked in zg_cuda_refresh.


static void op_ranges(const ZgCudaProgram* p, const ZgOp& op, std::vector<ZgRange>& out) {
    out.clear();
    auto whole = [&](uint32_t b, bool w) { out.push_back({b, 0, p->buffer_elems[b], w, 0}); };
    auto span = [&](uint32_t b, size_t lo, size_t n, bool w) { out.push_back({b, lo, lo + n, w, 0}); };
    switch (op.tag) {
        case ZG_OP_ELEMENTWISE: {
            const auto& e = op.u.elementwise;
            span(e.dst, e.dst_offset, e.n, true); span(e.src0, e.src0_offset, e.n, false); span(e.src1, e.src1_offset, e.n, false);
            break;
        }
        case ZG_OP_MATMUL: whole(op.u.matmul.dst, true); whole(op.u.matmul.a, false); whole(op.u.matmul.b, false); break;
        case ZG_OP_QMATMUL: {
            const auto& q = op.u.qmatmul;
            const size_t irs = q.input_row_stride ? q.input_row_stride : q.K, drs = q.dst_row_stride ? q.dst_row_stride : q.N;
            if (q.M == 0) break;
            span(q.dst, q.dst_offset, (size_t)(q.M - 1) * drs + q.N, true);
            span(q.input, q.input_offset, (size_t)(q.M - 1) * irs + q.K, false);
            // the tensor-core path stages its rounded activations in ONE shared scratch: a virtual buffer orders its users
            if (zg_qgemm_scratch_elems(p->qweights[q.weight_idx], q.M)) span((uint32_t)p->buffers.size(), 0, 1, true);
            break;
        }
        case ZG_OP_SOFTMAX: span(op.u.softmax.dst, op.u.softmax.dst_offset, (size_t)op.u.softmax.rows * op.u.softmax.cols, true);
                            span(op.u.softmax.src, op.u.softmax.src_offset, (size_t)op.u.softmax.rows * op.u.softmax.cols, false); break;
        case ZG_OP_LAYERNORM: span(op.u.layernorm.dst, op.u.layernorm.dst_offset, (size_t)op.u.layernorm.rows * op.u.layernorm.cols, true);
                              span(op.u.layernorm.src, op.u.layernorm.src_offset, (size_t)op.u.layernorm.rows * op.u.layernorm.cols, false); break;
        case ZG_OP_RMSNORM: span(op.u.rmsnorm.dst, op.u.rmsnorm.dst_offset, (size_t)op.u.rmsnorm.rows * op.u.rmsnorm.cols, true);
                            span(op.u.rmsnorm.src, op.u.rmsnorm.src_offset, (size_t)op.u.rmsnorm.rows * op.u.rmsnorm.cols, false); break;
        case ZG_OP_REDUCE: whole(op.u.reduce.dst, true); whole(op.u.reduce.src, false); break;
        case ZG_OP_REPEAT: whole(op.u.repeat.dst, true); whole(op.u.repeat.src, false); break;
        case ZG_OP_SLICE_ASSIGN: {
            const auto& sa = op.u.slice_assign;
            if (sa.rows == 0 || sa.cols == 0) break;
            const size_t dext = (size_t)(sa.rows - 1) * sa.dst_row_stride + (size_t)(sa.cols - 1) * sa.dst_col_stride + 1;
            const size_t sext = (size_t)(sa.rows - 1) * sa.src_row_stride + (size_t)(sa.cols - 1) * sa.src_col_stride + 1;
            if (sa.patch_stride == 0 && sa.dst_row_stride == 1 && sa.cols > 1 && sa.dst_col_stride >= sa.rows) {
                ZgRange r{sa.dst, sa.dst_offset, sa.dst_offset + dext, true, 0};
                r.stride = sa.dst_col_stride; r.width = sa.rows;   // per-head column window: heads do not overlap
                out.push_back(r);
            } else if (sa.patch_stride == 0) span(sa.dst, sa.dst_offset, dext, true);
            else if (p->uniform_pos) out.push_back({sa.dst, sa.dst_base_offset, sa.dst_base_offset + dext, true, sa.patch_stride});
            else whole(sa.dst, true);
            span(sa.src, sa.src_offset, sext, false);
            break;
        }
        case ZG_OP_ROPE: whole(op.u.rope.dst, true); whole(op.u.rope.src, false); whole(op.u.rope.cos_sin, false); break;
        case ZG_OP_ATTENTION: {
            const auto& a = op.u.attention;
            whole(a.dst, true); whole(a.q, false); whole(a.k, false); whole(a.v, false);
            if (a.has_mask) whole(a.mask, false);
            break;
        }
        case ZG_OP_FUSED_ELEMENTWISE: {
            const auto& f = op.u.fused_elementwise;
            span(f.dst, f.dst_offset, f.n, true); span(f.src, f.src_offset, f.n, false);
            for (size_t k = 0; k < f.n_steps; k++)
                if (f.steps[k].op == ZG_EW_ADD || f.steps[k].op == ZG_EW_MUL) span(f.steps[k].secondary_buf, f.steps[k].secondary_offset, f.n, false);
            break;
        }
        case ZG_OP_ALLREDUCE:
            span(op.u.allreduce.buf, op.u.allreduce.offset, op.u.allreduce.n, true);
            span((uint32_t)p->buffers.size() + 1, 0, 1, true);   // one communicator: collectives stay in program order
            break;
        case ZG_OP_ALLGATHER:
            span(op.u.allgather.dst, op.u.allgather.dst_offset, (size_t)op.u.allgather.n * p->ctx->world, true);
            span(op.u.allgather.src, op.u.allgather.src_offset, op.u.allgather.n, false);
            span((uint32_t)p->buffers.size() + 1, 0, 1, true);
            break;
        default: break;
    }
}

static inline bool range_conflict(const ZgRange& x, const ZgRange& y) {
    if (x.buf != y.buf || !(x.write || y.write)) return false;
    if (x.dyn && x.dyn == y.dyn) return x.lo < y.hi && y.lo < x.hi;   // same run-time shift: compare the bases
    if (x.stride && x.stride == y.stride && !x.dyn && !y.dyn) {       // same row pitch: disjoint column windows never meet
        const size_t ax = x.lo % x.stride, ay = y.lo % y.stride;
        if (ax + x.width <= x.stride && ay + y.width <= y.stride && (ax + x.width <= ay || ay + y.width <= ax)) return false;
    }
    const size_t xhi = x.dyn ? (size_t)-1 : x.hi, yhi = y.dyn ? (size_t)-1 : y.hi;   // shifted range: anywhere above its base
    return x.lo < yhi && y.lo < xhi;
}
static bool ranges_conflict(const std::vector<ZgRange>& a, const std::vector<ZgRange>& b) {
    for (const ZgRange& x : a)
        for (const ZgRange& y : b)
            if (range_conflict(x, y)) return true;
    return false;
}

// Offsets / extents of the run-time patched fields against the buffer sizes (the reference's CPU slices would trap on an
// out-of-range position, src/backend/reference.zig:436-458,599-671; here it would be a silent out-of-bounds access).
static bool dyn_extent_ok(const ZgCudaProgram* p, const ZgOp& op, size_t i) {
    if (op.tag == ZG_OP_SLICE_ASSIGN) {
        const auto& sa = op.u.slice_assign;
        if (sa.rows == 0 || sa.cols == 0) return true;
        const size_t dext = (size_t)(sa.rows - 1) * sa.dst_row_stride + (size_t)(sa.cols - 1) * sa.dst_col_stride + 1;
        if ((size_t)sa.dst_offset + dext > p->buffer_elems[sa.dst]) {
            zg_set_error("op %zu: slice_assign writes [%u, %zu) of a %zu-element buffer", i, sa.dst_offset, (size_t)sa.dst_offset + dext, p->buffer_elems[sa.dst]);
            return false;
        }
    } else if (op.tag == ZG_OP_ATTENTION) {
        const auto& a = op.u.attention;
        if (a.seq_kv == 0 || a.d_head == 0 || a.seq_q == 0) return true;
        const size_t kext = (size_t)a.k_off + (size_t)(a.seq_kv - 1) * a.k_cs + (size_t)(a.d_head - 1) * a.k_rs + 1;
        const size_t vext = (size_t)a.v_off + (size_t)(a.seq_kv - 1) * a.v_cs + (size_t)(a.d_head - 1) * a.v_rs + 1;
        const size_t mext = a.has_mask ? (size_t)a.mask_off + (size_t)(a.seq_kv - 1) * a.mask_rs + (size_t)(a.seq_q - 1) * a.mask_cs + 1 : 0;
        if (kext > p->buffer_elems[a.k] || vext > p->buffer_elems[a.v] || (a.has_mask && mext > p->buffer_elems[a.mask])) {
            zg_set_error("op %zu: attention over seq_kv = %u reads past its k / v / mask buffer", i, a.seq_kv);
            return false;
        }
    }
    return true;
}
// every op's static element ranges inside its buffers (compile time and structural refresh)
static bool validate_extents(const ZgCudaProgram* p) {
    std::vector<ZgRange> rr;
    for (size_t i = 0; i < p->ops.size(); i++) {
        op_ranges(p, p->ops[i], rr);
        for (const ZgRange& r : rr) {
            if (r.buf >= p->buffers.size() || r.dyn) continue;   // virtual ordering buffers; patched ranges are checked with their run-time value
            if (r.hi > p->buffer_elems[r.buf] && !(p->ops[i].tag == ZG_OP_ALLGATHER)) {
                zg_set_error("op %zu (tag %u): touches [%zu, %zu) of buffer %u which has %zu elements", i, p->ops[i].tag, r.lo, r.hi, r.buf, p->buffer_elems[r.buf]);
                return false;
            }
        }
        if (!dyn_extent_ok(p, p->ops[i], i)) return false;
    }
    return true;
}

extern "C" ZgCudaProgram* zg_cuda_compile(ZgCudaCtx* ctx, const ZgProgram* prog) {
    if (!ctx || !prog) { zg_set_error("compile: null argument"); return nullptr; }
    if (prog->n_buffers > 65535) { zg_set_error("compile: more than 65535 buffers"); return nullptr; }
    cudaSetDevice(ctx->device);
    ZgCudaProgram* p = new ZgCudaProgram();
    p->ctx = ctx;
    memset(&p->profile, 0, sizeof(p->profile));
    // buffers: max(size,1) f32, zero-filled (src/backend/reference.zig:89-93)
    p->buffers.assign(prog->n_buffers, nullptr);
    p->buffer_elems.assign(prog->n_buffers, 0);
    for (size_t i = 0; i < prog->n_buffers; i++) {
        size_t n = prog->buffer_sizes[i] > 1 ? prog->buffer_sizes[i] : 1;
        size_t bytes = (n * sizeof(float) + 15) & ~(size_t)15;
        if (cudaMalloc(&p->buffers[i], bytes) != cudaSuccess) {
            zg_set_error("compile: cudaMalloc(%zu B) for buffer %zu failed", bytes, i);
            free_program(p); return nullptr;
        }
        cudaMemsetAsync(p->buffers[i], 0, bytes, ctx->stream);
        p->buffer_elems[i] = n;
    }
    for (size_t i = 0; i < prog->n_uploads; i++) {
        const ZgIO& io = prog->initial_uploads[i];
        if (io.buf_idx >= prog->n_buffers || (size_t)io.offset + io.size > p->buffer_elems[io.buf_idx] * sizeof(float)) {
            zg_set_error("compile: initial upload %zu out of range", i);
            free_program(p); return nullptr;
        }
        cudaMemcpyAsync((uint8_t*)p->buffers[io.buf_idx] + io.offset, io.host_ptr, io.size, cudaMemcpyHostToDevice, ctx->stream);
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { zg_set_error("compile: buffer initialisation failed"); free_program(p); return nullptr; }
    for (size_t i = 0; i < prog->n_qweights; i++) {
        const ZgQWeight& d = prog->qweights[i];
        if (d.block_size == ZG_QWEIGHT_RESIDENT) {   // already packed in HBM by zg_cuda_qweight_upload*: borrow it
            ZgCudaQWeight* w = (ZgCudaQWeight*)d.data;
            if (!w || w->K != d.rows || w->N != d.cols) { zg_set_error("compile: resident qweight %zu is null or not [%zu,%zu]", i, d.rows, d.cols); free_program(p); return nullptr; }
            p->qweights.push_back(w); p->qweight_owned.push_back(false);
            continue;
        }
        ZgCudaQWeight* w = zg_cuda_qweight_upload(ctx, &d, ZG_QFMT_AUTO);
        if (!w) { free_program(p); return nullptr; }
        p->qweights.push_back(w); p->qweight_owned.push_back(true);
    }
    if (!validate_ops(p, prog->ops, prog->n_ops) || !adopt_ops(p, prog->ops, prog->n_ops) || !validate_extents(p) || !upload_steps(p) ||
        !reserve_workspace(p) || !build_schedule(p)) {
        free_program(p); return nullptr;
    }
    size_t nd = prog->n_ops ? prog->n_ops : 1;
    if (cudaMalloc(&p->d_dyn, nd * 4) != cudaSuccess || cudaMallocHost(&p->h_dyn, nd * 4) != cudaSuccess) {
        zg_set_error("compile: dyn table allocation failed"); free_program(p); return nullptr;
    }
    for (size_t i = 0; i < prog->n_ops; i++) p->h_dyn[i] = op_dyn_value(p->ops[i]);
    return p;
}

extern "C" void zg_cuda_refresh(ZgCudaCtx* ctx, ZgCudaProgram* p, const ZgOp* ops, size_t n_ops) {
    if (!ctx || !p || !ops) return;
    bool structural = (n_ops != p->ops.size());
    if (!structural) {
        for (size_t i = 0; i < n_ops && !structural; i++) {
            const ZgOp& old = p->ops[i];
            if (ops[i].tag != old.tag) { structural = true; break; }
            // everything but the per-step field must be unchanged; compared in place (no copy: an 80-layer program has ~9 k ops
            // and this runs every token)
            auto same_except = [&](size_t off, size_t len) {
                const char* a = reinterpret_cast<const char*>(&ops[i]);
                const char* b = reinterpret_cast<const char*>(&old);
                return memcmp(a, b, off) == 0 && memcmp(a + off + len, b + off + len, sizeof(ZgOp) - off - len) == 0;
            };
            if (ops[i].tag == ZG_OP_SLICE_ASSIGN) {
                if (!same_except(offsetof(ZgOp, u.slice_assign.dst_offset), sizeof(old.u.slice_assign.dst_offset))) structural = true;
            } else if (ops[i].tag == ZG_OP_ATTENTION) {
                if (!same_except(offsetof(ZgOp, u.attention.seq_kv), sizeof(old.u.attention.seq_kv))) structural = true;
            } else if (ops[i].tag == ZG_OP_FUSED_ELEMENTWISE) {
                const auto& f = ops[i].u.fused_elementwise;
                if (f.n_steps != p->steps[i].size() ||
                    (f.n_steps && memcmp(f.steps, p->steps[i].data(), f.n_steps * sizeof(ZgFusedEwStep)) != 0)) { structural = true; break; }
                if (!same_except(offsetof(ZgOp, u.fused_elementwise.steps), sizeof(old.u.fused_elementwise.steps))) structural = true;
            } else if (memcmp(&ops[i], &old, sizeof(ZgOp)) != 0) structural = true;
        }
    }
    if (structural) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        if (!validate_ops(p, ops, n_ops)) return; // keep the previous op list; error string set
        adopt_ops(p, ops, n_ops);
        if (!validate_extents(p)) { p->ops.clear(); p->steps.clear(); n_ops = 0; }   // out-of-range op list: the program becomes a no-op (error string set)
        upload_steps(p);
        reserve_workspace(p);
        build_schedule(p);
        cudaFree(p->d_dyn); cudaFreeHost(p->h_dyn); p->d_dyn = nullptr; p->h_dyn = nullptr;
        size_t nd = n_ops ? n_ops : 1;
        cudaMalloc(&p->d_dyn, nd * 4); cudaMallocHost(&p->h_dyn, nd * 4);
        memset(p->h_dyn, 0xFF, nd * 4);
        p->dyn_dirty = true;
        p->graph_valid = false;
    }
    if (p->uniform_pos) {   // the schedule assumes one KV write position per step; otherwise fall back to conservative ranges
        long pos = -1;
        for (size_t i = 0; i < n_ops; i++) {
            if (ops[i].tag != ZG_OP_SLICE_ASSIGN || ops[i].u.slice_assign.patch_stride == 0) continue;
            const auto& sa = ops[i].u.slice_assign;
            const long d = (long)sa.dst_offset - (long)sa.dst_base_offset;
            const long q = d / (long)sa.patch_stride;
            if (d < 0 || d % (long)sa.patch_stride != 0 || (pos >= 0 && q != pos)) {
                cudaStreamSynchronize(ctx->stream);
                p->uniform_pos = false;
                build_schedule(p);
                p->graph_valid = false;
                break;
            }
            pos = q;
        }
    }
    for (size_t i = 0; i < n_ops; i++) {   // patched positions must stay inside their buffers: reject the refresh otherwise (old values stay)
        if ((ops[i].tag == ZG_OP_SLICE_ASSIGN || ops[i].tag == ZG_OP_ATTENTION) && op_dyn_value(ops[i]) != p->h_dyn[i] && !dyn_extent_ok(p, ops[i], i)) return;
    }
    for (size_t i = 0; i < n_ops; i++) {
        uint32_t v = op_dyn_value(ops[i]);
        if (p->h_dyn[i] != v) p->dyn_dirty = true;
        p->h_dyn[i] = v;
        if (ops[i].tag == ZG_OP_SLICE_ASSIGN) p->ops[i].u.slice_assign.dst_offset = v;
        else if (ops[i].tag == ZG_OP_ATTENTION) p->ops[i].u.attention.seq_kv = v;
    }
}

// Dependency levels (level = 1 + max level of every earlier op it conflicts with), then units in (level, program
// order); any such order is a valid topological order of the program's dependency DAG.
// ── macro patterns: consecutive DeviceOps that the lowering always emits together ────────────────────────────
// Each is evaluated in ONE pass (values kept in registers, every op's output buffer still written), so the result of
// every op is what the op-by-op execution gives.  Matching is deliberately strict: consecutive in program order, whole
// contiguous vectors, all buffers distinct.
struct ZgItem { uint32_t first = 0, count = 1, kind = 0; };   // kind 0: one op; 1: [add,] rmsnorm, repeat, mul; 2: fused_elementwise, mul; 3: attention, slice_assign
enum { ITEM_OP = 0, ITEM_NORM = 1, ITEM_EWMUL = 2, ITEM_ATTN_STORE = 3, ITEM_DECODE = 4, ITEM_KVQ = 5, ITEM_ATTN_LAYER = 6 };

static bool distinct(std::initializer_list<uint32_t> bufs) {
    std::vector<uint32_t> v(bufs);
    std::sort(v.begin(), v.end());
    return std::adjacent_find(v.begin(), v.end()) == v.end();
}

// [elementwise add(sum = a + b)] ; rmsnorm(bare = norm(sum)) ; repeat(gamma_rep = gamma tiled over rows) ; mul(dst = bare * gamma_rep)
static uint32_t match_norm(const ZgCudaProgram* p, size_t i, ZgNormMacro* m) {
    const size_t n = p->ops.size();
    const bool has_add = p->ops[i].tag == ZG_OP_ELEMENTWISE && p->ops[i].u.elementwise.op == ZG_EW_ADD;
    const size_t r0 = i + (has_add ? 1 : 0);
    if (r0 + 2 >= n) return 0;
    const ZgOp &rn = p->ops[r0], &rp = p->ops[r0 + 1], &ml = p->ops[r0 + 2];
    if (rn.tag != ZG_OP_RMSNORM || rp.tag != ZG_OP_REPEAT || ml.tag != ZG_OP_ELEMENTWISE || ml.u.elementwise.op != ZG_EW_MUL) return 0;
    const auto& r = rn.u.rmsnorm;
    const auto& q = rp.u.repeat;
    const auto& e = ml.u.elementwise;
    const uint32_t rows = r.rows, cols = r.cols, tot = rows * cols;
    if (rows == 0 || rows > 8 || cols == 0 || (cols & 3u) || cols > 8192 || tot > p->ctx->chain_max) return 0;   // 4 float4 x 512 threads per row
    if (r.src_offset || r.dst_offset || q.src_offset || q.dst_offset || e.dst_offset || e.src0_offset || e.src1_offset) return 0;
    if (q.n != tot || e.n != tot) return 0;
    // gamma: a contiguous [cols] vector tiled over the rows
    if (q.src_ne[0] != cols || q.src_ne[1] != 1 || q.src_ne[2] != 1 || q.src_ne[3] != 1 || q.src_strides[0] != 1) return 0;
    if (q.dst_ne[0] != cols || q.dst_ne[1] != rows || q.dst_ne[2] != 1 || q.dst_ne[3] != 1 || q.dst_strides[0] != 1 || (rows > 1 && q.dst_strides[1] != cols)) return 0;
    if (e.src0 != r.dst || e.src1 != q.dst) return 0;
    uint32_t a_buf = r.src, b_buf = UINT32_MAX, sum_buf = UINT32_MAX;
    if (has_add) {
        const auto& ad = p->ops[i].u.elementwise;
        if (ad.n != tot || ad.dst_offset || ad.src0_offset || ad.src1_offset || ad.dst != r.src) return 0;
        a_buf = ad.src0; b_buf = ad.src1; sum_buf = ad.dst;
        if (!distinct({ad.src0, ad.src1, ad.dst, r.dst, q.src, q.dst, e.dst})) return 0;
    } else if (!distinct({r.src, r.dst, q.src, q.dst, e.dst})) return 0;
    for (uint32_t b : {a_buf, r.dst, q.dst, e.dst}) if (p->buffer_elems[b] < tot) return 0;
    if (p->buffer_elems[q.src] < cols) return 0;
    m->a = p->buffers[a_buf]; m->b = has_add ? p->buffers[b_buf] : nullptr; m->sum = has_add ? p->buffers[sum_buf] : nullptr;
    m->bare = p->buffers[r.dst]; m->gamma = p->buffers[q.src]; m->gamma_rep = p->buffers[q.dst]; m->norm = p->buffers[e.dst];
    m->rows = rows; m->cols = cols; m->eps = r.eps;
    return has_add ? 4 : 3;
}

// fused_elementwise(mid = chain(src)) ; mul(dst = mid * other)
static uint32_t match_ewmul(const ZgCudaProgram* p, size_t i, ZgEwMulMacro* m) {
    if (i + 1 >= p->ops.size()) return 0;
    const ZgOp &fo = p->ops[i], &mo = p->ops[i + 1];
    if (fo.tag != ZG_OP_FUSED_ELEMENTWISE || mo.tag != ZG_OP_ELEMENTWISE || mo.u.elementwise.op != ZG_EW_MUL) return 0;
    const auto& f = fo.u.fused_elementwise;
    const auto& e = mo.u.elementwise;
    if (f.n == 0 || f.n != e.n || f.dst_offset || f.src_offset || e.dst_offset || e.src0_offset || e.src1_offset) return 0;
    if (e.src0 != f.dst || e.src1 == f.dst || e.dst == f.dst || e.dst == f.src || e.dst == e.src1 || f.dst == f.src) return 0;
    for (const ZgFusedEwStep& st : p->steps[i])
        if ((st.op == ZG_EW_ADD || st.op == ZG_EW_MUL) && (st.secondary_buf == f.dst || st.secondary_buf == e.dst)) return 0;
    m->src = p->buffers[f.src]; m->mid = p->buffers[f.dst]; m->other = p->buffers[e.src1]; m->dst = p->buffers[e.dst];
    m->steps = p->d_steps + p->step_off[i]; m->n_steps = (uint32_t)f.n_steps; m->n = f.n;
    return 2;
}

// attention(out) ; slice_assign(concat[...] = out): the copy of a head's whole output into the concatenated buffer
static uint32_t match_attn_store(const ZgCudaProgram* p, size_t i) {
    if (i + 1 >= p->ops.size()) return 0;
    const ZgOp &ao = p->ops[i], &so = p->ops[i + 1];
    if (ao.tag != ZG_OP_ATTENTION || so.tag != ZG_OP_SLICE_ASSIGN) return 0;
    const auto& a = ao.u.attention;
    const auto& sa = so.u.slice_assign;
    if (sa.patch_stride != 0 || sa.src != a.dst || sa.dst == a.dst || sa.dst == a.q || sa.dst == a.k || sa.dst == a.v || sa.dst == a.mask) return 0;
    if (sa.rows != a.d_head || sa.cols != a.seq_q || sa.src_offset != a.dst_off || sa.src_row_stride != a.dst_rs || sa.src_col_stride != a.dst_cs) return 0;
    return 2;
}

// The attention block of one layer of a single-token program, starting at op i:
//   per KV head: rope(k_rot <- k_proj) ; slice_assign(K cache <- k_rot, patched) ; slice_assign(V cache <- v_proj, patched)
//   per head:    rope(q_rot <- q_proj) ; attention ; slice_assign(concat <- attn_out)
// Returns the number of ops matched (0: no match) and fills the launch descriptor.  As strict as the decode matcher.
static uint32_t match_attention_block(const ZgCudaProgram* p, size_t i, ZgAttnBlock* B) {
    const size_t n = p->ops.size();
    auto elems = [&](uint32_t b) { return p->buffer_elems[b]; };
    size_t c = i;
    memset(B, 0, sizeof(*B));
    if (c + 6 > n || p->ops[c].tag != ZG_OP_ROPE || p->ops[c + 1].tag != ZG_OP_SLICE_ASSIGN || p->ops[c + 1].u.slice_assign.patch_stride == 0) return 0;
    const uint32_t k_buf = p->ops[c].u.rope.src;
    uint32_t dh = 0, cs_buf = UINT32_MAX, kc_buf = UINT32_MAX, vc_buf = UINT32_MAX, v_buf = UINT32_MAX, patch = 0, n_kv = 0;
    std::vector<uint32_t> written, readonly;
    while (c + 3 <= n && n_kv < kZgDecMaxHeads && p->ops[c].tag == ZG_OP_ROPE && p->ops[c].u.rope.src == k_buf && p->ops[c + 1].tag == ZG_OP_SLICE_ASSIGN &&
           p->ops[c + 1].u.slice_assign.patch_stride != 0) {
        const auto& ro = p->ops[c].u.rope;
        const ZgOp &ks = p->ops[c + 1], &vs = p->ops[c + 2];
        if (vs.tag != ZG_OP_SLICE_ASSIGN) return 0;
        const auto& ka = ks.u.slice_assign; const auto& va = vs.u.slice_assign;
        const uint32_t d = 2 * ro.half_d;
        if (dh == 0) { dh = d; cs_buf = ro.cos_sin; kc_buf = ka.dst; vc_buf = va.dst; v_buf = va.src; patch = ka.patch_stride; }
        if (d != dh || d == 0 || ro.seq_len != 1 || ro.cos_sin != cs_buf || ro.cs_off != 0 || ro.dst_off != 0 || ro.src_rs != 1 ||
            (size_t)ro.src_off + dh > elems(k_buf) || elems(ro.dst) < dh || elems(cs_buf) < dh) return 0;
        if (ka.dst != kc_buf || ka.src != ro.dst || ka.rows != dh || ka.cols != 1 || ka.dst_row_stride != 1 || ka.src_offset != 0 ||
            ka.src_row_stride != 1 || ka.patch_stride != patch) return 0;
        if (va.dst != vc_buf || va.src != v_buf || va.rows != dh || va.cols != 1 || va.dst_row_stride != 1 || va.src_row_stride != 1 ||
            va.patch_stride != patch || (size_t)va.src_offset + dh > elems(v_buf)) return 0;
        ZgDecKv& kv = B->kvs[n_kv++];
        kv.k_rot = p->buffers[ro.dst]; kv.k_src = ro.src_off; kv.v_src = va.src_offset; kv.k_dyn = (uint32_t)(c + 1); kv.v_dyn = (uint32_t)(c + 2);
        kv.k_base = ka.dst_base_offset; kv.v_base = va.dst_base_offset;
        written.push_back(ro.dst);
        c += 3;
    }
    if (n_kv == 0 || kc_buf == vc_buf) return 0;
    if (c + 3 > n || p->ops[c].tag != ZG_OP_ROPE) return 0;
    const uint32_t q_buf = p->ops[c].u.rope.src;
    uint32_t n_heads = 0, mask_buf = UINT32_MAX, cat_buf = UINT32_MAX;
    while (c + 3 <= n && n_heads < kZgDecMaxHeads && p->ops[c].tag == ZG_OP_ROPE && p->ops[c].u.rope.src == q_buf && p->ops[c + 1].tag == ZG_OP_ATTENTION) {
        const auto& ro = p->ops[c].u.rope;
        const ZgOp &ao = p->ops[c + 1], &so = p->ops[c + 2];
        if (so.tag != ZG_OP_SLICE_ASSIGN) return 0;
        const auto& a = ao.u.attention; const auto& sa = so.u.slice_assign;
        const uint32_t h = n_heads;
        if (2 * ro.half_d != dh || ro.seq_len != 1 || ro.cos_sin != cs_buf || ro.cs_off != 0 || ro.dst_off != 0 || ro.src_rs != 1 ||
            (size_t)ro.src_off + dh > elems(q_buf) || elems(ro.dst) < dh) return 0;
        if (a.q != ro.dst || a.k != kc_buf || a.v != vc_buf || a.d_head != dh || a.seq_q != 1 || a.q_off != 0 || a.q_rs != 1 || a.k_rs != 1 ||
            a.v_rs != 1 || a.dst_off != 0 || a.dst_rs != 1 || (dh % 4) != 0 || dh > 256 || (a.k_off % 4) != 0 || (a.k_cs % 4) != 0 ||
            a.k_cs != patch || a.v_cs != patch || a.k_cs == 0 || elems(a.dst) < dh) return 0;
        if (h == 0) {
            B->has_mask = a.has_mask; mask_buf = a.mask; B->mask_off = a.mask_off; B->mask_rs = a.mask_rs; B->scale = a.scale;
            { static const uint32_t mp = [] { const char* e = getenv("ZG_CUDA_ATTN_MIN_POS"); const int v = e ? atoi(e) : 64; return (uint32_t)(v >= 32 ? v : 64); }(); B->min_pos = mp; }   // 64: an 8-head shard at 512 positions uses 8 splits (56.1 -> 54.4 us per layer); more CTAs than SMs is slower (ZG_CUDA_ATTN_SPLIT_MULT)
            B->k_cs = a.k_cs; B->v_cs = a.v_cs; cat_buf = sa.dst;
        } else if (a.has_mask != B->has_mask || (a.has_mask && (a.mask != mask_buf || a.mask_off != B->mask_off || a.mask_rs != B->mask_rs)) ||
                   a.scale != B->scale || a.k_cs != B->k_cs || a.v_cs != B->v_cs) return 0;
        uint32_t kvi = UINT32_MAX;
        for (uint32_t g = 0; g < n_kv; g++) if (B->kvs[g].k_base == a.k_off && B->kvs[g].v_base == a.v_off) kvi = g;
        if (kvi == UINT32_MAX || (h > 0 && kvi < B->heads[h - 1].kv)) return 0;
        if (sa.dst != cat_buf || sa.src != a.dst || sa.patch_stride != 0 || sa.rows != dh || sa.cols != 1 || sa.dst_row_stride != 1 ||
            sa.src_offset != 0 || sa.src_row_stride != 1 || (size_t)sa.dst_offset + dh > elems(cat_buf)) return 0;
        ZgDecHead& hd = B->heads[n_heads++];
        hd.q_rot = p->buffers[ro.dst]; hd.attn_out = p->buffers[a.dst]; hd.q_src = ro.src_off; hd.k_off = a.k_off; hd.v_off = a.v_off; hd.kv = kvi;
        hd.buf_off = sa.dst_offset; hd.dyn = (uint32_t)(c + 1);
        written.push_back(ro.dst); written.push_back(a.dst);
        c += 3;
    }
    if (n_heads == 0) return 0;
    for (uint32_t g = 0; g < n_kv; g++) {   // every KV head is read by some query head (its first one stores the cache rows)
        bool used = false;
        for (uint32_t h = 0; h < n_heads; h++) used = used || B->heads[h].kv == g;
        if (!used) return 0;
    }
    // distinct buffers: projections, rope table, mask, caches, concat and the per-head temporaries
    written.push_back(kc_buf); written.push_back(vc_buf); written.push_back(cat_buf);
    readonly = {q_buf, k_buf, v_buf, cs_buf};
    if (B->has_mask) readonly.push_back(mask_buf);
    std::vector<uint32_t> all(written);
    all.insert(all.end(), readonly.begin(), readonly.end());
    std::sort(all.begin(), all.end());
    if (std::adjacent_find(all.begin(), all.end()) != all.end()) {
        // q / k / v may legitimately be distinct windows of one buffer?  The lowering uses three buffers: anything else keeps the general path
        return 0;
    }
    B->q_proj = p->buffers[q_buf]; B->k_proj = p->buffers[k_buf]; B->v_proj = p->buffers[v_buf]; B->cs = p->buffers[cs_buf];
    B->mask = B->has_mask ? p->buffers[mask_buf] : nullptr; B->k_cache = p->buffers[kc_buf]; B->v_cache = p->buffers[vc_buf];
    B->attn_buf = p->buffers[cat_buf]; B->n_heads = n_heads; B->n_kv = n_kv; B->d_head = dh;
    return (uint32_t)(c - i);
}

// ── fused decode kernel (decode.cu): recognise the single-token layer pattern of the LLaMA lowering ───────────────
// (src/device_inference.zig:61-238 over src/models/llama_transformer.zig:192-253; mirrored by host/llama.py build_program)
//   [add,] rmsnorm, repeat, mul ; qmatmul q, k, v ; per KV head: rope, slice_assign(K cache), slice_assign(V cache) ;
//   per head: rope, attention, slice_assign(concat) ; qmatmul o [; allreduce] ; add, rmsnorm, repeat, mul ;
//   qmatmul gate, up ; fused_elementwise, mul ; qmatmul down [; allreduce]
// Matching is strict — every offset, stride and buffer role is checked — and anything else keeps the general schedule.
struct DecHostLayer {
    ZgDecLayer ly; ZgDecPhase ph[4];
    std::vector<ZgDecHead> heads; std::vector<ZgDecKv> kvs;
    uint32_t part_elems[4];   // split-K scratch floats per phase
};

static bool plan_decode_phase(uint32_t grid, uint32_t n_kc, uint32_t n_items, bool whole_vector, uint32_t* S_out, uint32_t* lS_out, uint32_t* slots_out) {
    // S = 2^lS k-splits x (grid >> lS) column-group slots; cost = rounds x (records per warp + per-item overhead)
    double best = 1e30; uint32_t bl = UINT32_MAX;
    for (uint32_t lS = 0; lS <= 6; lS++) {
        const uint32_t S = 1u << lS;
        if (S > n_kc || S > grid) break;
        const uint32_t recs = (n_kc + S - 1) / S;
        if (!whole_vector && (size_t)recs * ZG_KR > kZgDecMaxD) continue;
        const uint32_t slots = grid >> lS, rounds = (n_items + slots - 1) / slots, per_warp = (recs + 15) / 16;
        if (rounds > kZgDecMaxItems) continue;
        const double cost = (double)rounds * ((double)per_warp + 1.5) + 0.02 * S;
        if (cost < best) { best = cost; bl = lS; }
    }
    if (bl == UINT32_MAX) return false;
    *S_out = 1u << bl; *lS_out = bl; *slots_out = grid >> bl;
    return true;
}

static bool match_decode_layers(ZgCudaProgram* p, size_t i0, std::vector<DecHostLayer>& out, size_t* end_out) {
    const size_t n = p->ops.size();
    const uint32_t grid = zg_decode_grid(p->ctx);
    auto bufp = [&](uint32_t b) { return p->buffers[b]; };
    auto elems = [&](uint32_t b) { return p->buffer_elems[b]; };
    size_t i = i0;
    uint32_t prev_after = UINT32_MAX, prev_down = UINT32_MAX;   // buffers of the previous matched layer
    ZgDecVec prev_down_vec = {};
    std::vector<uint32_t> written, external;   // buffer indices: written inside the range / required to be untouched by it
    while (i < n) {
        DecHostLayer L; memset(&L.ly, 0, sizeof(L.ly)); memset(L.ph, 0, sizeof(L.ph)); memset(L.part_elems, 0, sizeof(L.part_elems));
        ZgDecLayer& ly = L.ly;
        size_t c = i;
        std::vector<uint32_t> roles;   // role buffers of this layer that must be pairwise distinct
        // ── norm block 1 ──
        ZgNormMacro nm = {};
        const uint32_t cnt1 = match_norm(p, c, &nm);
        if (!cnt1 || nm.rows != 1 || nm.cols > kZgDecMaxD) break;
        const bool has_add = cnt1 == 4;
        const uint32_t D = nm.cols;
        uint32_t x_buf;   // the layer input as later ops name it
        {
            const size_t r0 = c + (has_add ? 1 : 0);
            const auto& r = p->ops[r0].u.rmsnorm; const auto& q = p->ops[r0 + 1].u.repeat; const auto& e = p->ops[r0 + 2].u.elementwise;
            if (has_add) {
                const auto& ad = p->ops[c].u.elementwise;
                if (out.empty() || ad.src0 != prev_after || ad.src1 != prev_down) break;   // only the previous layer's closing add is absorbed
                ly.x1_a = bufp(ad.src0); ly.x1_b = prev_down_vec; ly.x1_sum = bufp(ad.dst);
                x_buf = ad.dst;
                roles.push_back(ad.dst);   // src0 / src1 are the previous layer's after_attn / down, rewritten by this layer's phases 4 / 5
                written.push_back(ad.dst);
            } else {
                if (!out.empty()) break;
                ly.x1_a = bufp(r.src); x_buf = r.src;
                roles.push_back(r.src);
                external.push_back(r.src);
            }
            ly.gamma1 = bufp(q.src); ly.bare1 = bufp(r.dst); ly.grep1 = bufp(q.dst); ly.norm1 = bufp(e.dst); ly.eps1 = r.eps; ly.D = D;
            roles.insert(roles.end(), {r.dst, q.src, q.dst, e.dst});
            written.insert(written.end(), {r.dst, q.dst, e.dst});
            external.push_back(q.src);
            c += cnt1;
            const uint32_t norm_buf = e.dst;
            // ── q, k, v ──
            if (c + 3 > n) break;
            const ZgOp* qo[3] = {&p->ops[c], &p->ops[c + 1], &p->ops[c + 2]};
            bool ok = true;
            for (int k = 0; k < 3 && ok; k++) {
                if (qo[k]->tag != ZG_OP_QMATMUL) { ok = false; break; }
                const auto& m = qo[k]->u.qmatmul;
                const ZgCudaQWeight* w = p->qweights[m.weight_idx];
                ok = m.M == 1 && m.input == norm_buf && m.K == D && m.input_offset == 0 && m.dst_offset == 0 && w->fmt != ZG_QFMT_GENERIC &&
                     w->fmt == p->qweights[qo[0]->u.qmatmul.weight_idx]->fmt && elems(m.dst) >= m.N;
            }
            if (!ok) break;
            c += 3;
        }
        const ZgOp *op_q = &p->ops[c - 3], *op_k = &p->ops[c - 2], *op_v = &p->ops[c - 1];
        const auto& mq = op_q->u.qmatmul; const auto& mk = op_k->u.qmatmul; const auto& mv = op_v->u.qmatmul;
        roles.insert(roles.end(), {mq.dst, mk.dst, mv.dst});
        written.insert(written.end(), {mq.dst, mk.dst, mv.dst});
        // ── per KV head: rope(k) ; K store ; V store ──
        uint32_t dh = 0, cs_buf = UINT32_MAX, kc_buf = UINT32_MAX, vc_buf = UINT32_MAX;
        while (c + 3 <= n && p->ops[c].tag == ZG_OP_ROPE && p->ops[c].u.rope.src == mk.dst) {
            const auto& ro = p->ops[c].u.rope;
            const ZgOp &ks = p->ops[c + 1], &vs = p->ops[c + 2];
            if (ks.tag != ZG_OP_SLICE_ASSIGN || vs.tag != ZG_OP_SLICE_ASSIGN) break;
            const auto& ka = ks.u.slice_assign; const auto& va = vs.u.slice_assign;
            const uint32_t d = 2 * ro.half_d;
            if (dh == 0) { dh = d; cs_buf = ro.cos_sin; kc_buf = ka.dst; vc_buf = va.dst; }
            if (d != dh || d == 0 || ro.seq_len != 1 || ro.cos_sin != cs_buf || ro.cs_off != 0 || ro.dst_off != 0 || ro.src_rs != 1 ||
                ro.src_off + dh > mk.N || elems(ro.dst) < dh || elems(cs_buf) < dh) break;
            if (ka.dst != kc_buf || ka.src != ro.dst || ka.rows != dh || ka.cols != 1 || ka.dst_row_stride != 1 || ka.src_offset != 0 ||
                ka.src_row_stride != 1 || ka.patch_stride == 0) break;
            if (va.dst != vc_buf || va.src != mv.dst || va.rows != dh || va.cols != 1 || va.dst_row_stride != 1 || va.src_row_stride != 1 ||
                va.patch_stride == 0 || va.src_offset + dh > mv.N) break;
            ZgDecKv kv;
            kv.k_rot = bufp(ro.dst); kv.k_src = ro.src_off; kv.v_src = va.src_offset; kv.k_dyn = (uint32_t)(c + 1); kv.v_dyn = (uint32_t)(c + 2);
            kv.k_base = ka.dst_base_offset; kv.v_base = va.dst_base_offset;
            if (ka.patch_stride != va.patch_stride) break;
            L.kvs.push_back(kv);
            roles.push_back(ro.dst);
            written.push_back(ro.dst);
            c += 3;
        }
        if (L.kvs.empty() || kc_buf == vc_buf) break;
        roles.insert(roles.end(), {cs_buf, kc_buf, vc_buf});
        written.insert(written.end(), {kc_buf, vc_buf});
        external.push_back(cs_buf);
        const uint32_t patch_stride = p->ops[L.kvs[0].k_dyn].u.slice_assign.patch_stride;
        // ── per head: rope(q) ; attention ; slice_assign into the concatenated buffer ──
        uint32_t mask_buf = UINT32_MAX, cat_buf = UINT32_MAX;
        bool heads_ok = true;
        while (c + 3 <= n && p->ops[c].tag == ZG_OP_ROPE && p->ops[c].u.rope.src == mq.dst) {
            const auto& ro = p->ops[c].u.rope;
            const ZgOp &ao = p->ops[c + 1], &so = p->ops[c + 2];
            if (ao.tag != ZG_OP_ATTENTION || so.tag != ZG_OP_SLICE_ASSIGN) { heads_ok = false; break; }
            const auto& a = ao.u.attention; const auto& sa = so.u.slice_assign;
            const uint32_t h = (uint32_t)L.heads.size();
            if (2 * ro.half_d != dh || ro.seq_len != 1 || ro.cos_sin != cs_buf || ro.cs_off != 0 || ro.dst_off != 0 || ro.src_rs != 1 ||
                ro.src_off + dh > mq.N || elems(ro.dst) < dh) { heads_ok = false; break; }
            if (a.q != ro.dst || a.k != kc_buf || a.v != vc_buf || a.d_head != dh || a.seq_q != 1 || a.q_off != 0 || a.q_rs != 1 || a.k_rs != 1 ||
                a.v_rs != 1 || a.dst_off != 0 || a.dst_rs != 1 || (dh % 4) != 0 || dh > 256 || (a.k_off % 4) != 0 || (a.k_cs % 4) != 0 ||
                a.k_cs != patch_stride || a.v_cs != patch_stride || a.k_cs == 0 || elems(a.dst) < dh) { heads_ok = false; break; }
            if (h == 0) {
                ly.has_mask = a.has_mask; mask_buf = a.mask; ly.mask_off = a.mask_off; ly.mask_rs = a.mask_rs; ly.scale = a.scale;
                ly.k_cs = a.k_cs; ly.v_cs = a.v_cs; cat_buf = sa.dst;
            } else if (a.has_mask != ly.has_mask || (a.has_mask && (a.mask != mask_buf || a.mask_off != ly.mask_off || a.mask_rs != ly.mask_rs)) ||
                       a.scale != ly.scale || a.k_cs != ly.k_cs || a.v_cs != ly.v_cs) { heads_ok = false; break; }
            // the KV head whose cache slab this head reads
            uint32_t kvi = UINT32_MAX;
            for (uint32_t g = 0; g < L.kvs.size(); g++) if (L.kvs[g].k_base == a.k_off && L.kvs[g].v_base == a.v_off) kvi = g;
            if (kvi == UINT32_MAX || (!L.heads.empty() && kvi < L.heads.back().kv)) { heads_ok = false; break; }
            if (sa.dst != cat_buf || sa.src != a.dst || sa.patch_stride != 0 || sa.rows != dh || sa.cols != 1 || sa.dst_row_stride != 1 ||
                sa.src_offset != 0 || sa.src_row_stride != 1 || sa.dst_offset != h * dh) { heads_ok = false; break; }
            ZgDecHead hd;
            hd.q_rot = bufp(ro.dst); hd.attn_out = bufp(a.dst); hd.q_src = ro.src_off; hd.k_off = a.k_off; hd.v_off = a.v_off; hd.kv = kvi;
            hd.buf_off = sa.dst_offset; hd.dyn = (uint32_t)(c + 1);
            L.heads.push_back(hd);
            roles.insert(roles.end(), {ro.dst, a.dst});
            written.insert(written.end(), {ro.dst, a.dst});
            c += 3;
        }
        if (!heads_ok || L.heads.empty()) break;
        for (uint32_t g = 0; g < L.kvs.size(); g++) {   // every KV head is read by at least one query head (its first one stores the cache rows)
            bool used = false;
            for (const ZgDecHead& hd : L.heads) used = used || hd.kv == g;
            if (!used) { heads_ok = false; break; }
        }
        if (!heads_ok) break;
        const uint32_t n_heads = (uint32_t)L.heads.size(), Dq = n_heads * dh;
        if (elems(cat_buf) < Dq) break;
        roles.push_back(cat_buf);
        written.push_back(cat_buf);
        if (ly.has_mask) { roles.push_back(mask_buf); external.push_back(mask_buf); }
        // ── o projection [+ all-reduce] ──
        if (c >= n || p->ops[c].tag != ZG_OP_QMATMUL) break;
        const ZgOp* op_o = &p->ops[c];
        const auto& mo = op_o->u.qmatmul;
        {
            const ZgCudaQWeight* w = p->qweights[mo.weight_idx];
            if (mo.M != 1 || mo.input != cat_buf || mo.K != Dq || mo.N != D || mo.input_offset != 0 || mo.dst_offset != 0 || w->fmt == ZG_QFMT_GENERIC) break;
        }
        c++;
        roles.push_back(mo.dst); written.push_back(mo.dst);
        bool ar_o = false;
        if (c < n && p->ops[c].tag == ZG_OP_ALLREDUCE) {
            const auto& ar = p->ops[c].u.allreduce;
            if (ar.buf != mo.dst || ar.offset != 0 || ar.n != D || (p->ctx->world > 1 && !zg_peer_allreduce_ok(p->ctx, D))) break;
            ar_o = p->ctx->world > 1;
            c++;
        }
        // ── norm block 2 (with the residual add) ──
        ZgNormMacro nm2 = {};
        if (c >= n || match_norm(p, c, &nm2) != 4 || nm2.rows != 1 || nm2.cols != D) break;
        const auto& ad2 = p->ops[c].u.elementwise;
        if (ad2.src0 != x_buf || ad2.src1 != mo.dst) break;
        {
            const auto& r = p->ops[c + 1].u.rmsnorm; const auto& q = p->ops[c + 2].u.repeat; const auto& e = p->ops[c + 3].u.elementwise;
            ly.x2_a = bufp(ad2.src0); ly.x2_sum = bufp(ad2.dst); ly.gamma2 = bufp(q.src); ly.bare2 = bufp(r.dst); ly.grep2 = bufp(q.dst);
            ly.norm2 = bufp(e.dst); ly.eps2 = r.eps;
            roles.insert(roles.end(), {ad2.dst, q.src});
            if (bufp(r.dst) != ly.bare1) roles.push_back(r.dst);
            if (bufp(q.dst) != ly.grep1) roles.push_back(q.dst);
            if (bufp(e.dst) != ly.norm1) roles.push_back(e.dst);
            written.insert(written.end(), {ad2.dst, r.dst, q.dst, e.dst});
            external.push_back(q.src);
        }
        const uint32_t after_buf = ad2.dst, norm2_buf = p->ops[c + 3].u.elementwise.dst;
        c += 4;
        // ── gate, up ──
        if (c + 2 > n || p->ops[c].tag != ZG_OP_QMATMUL || p->ops[c + 1].tag != ZG_OP_QMATMUL) break;
        const ZgOp *op_g = &p->ops[c], *op_u = &p->ops[c + 1];
        const auto& mg = op_g->u.qmatmul; const auto& mu = op_u->u.qmatmul;
        {
            const ZgCudaQWeight *wg = p->qweights[mg.weight_idx], *wu = p->qweights[mu.weight_idx];
            if (mg.M != 1 || mu.M != 1 || mg.input != norm2_buf || mu.input != norm2_buf || mg.K != D || mu.K != D || mg.N != mu.N ||
                mg.input_offset || mu.input_offset || mg.dst_offset || mu.dst_offset || wg->fmt == ZG_QFMT_GENERIC || wg->fmt != wu->fmt ||
                mg.dst == mu.dst) break;
        }
        const uint32_t F = mg.N;
        c += 2;
        roles.insert(roles.end(), {mg.dst, mu.dst}); written.insert(written.end(), {mg.dst, mu.dst});
        // ── activation chain * up ──
        ZgEwMulMacro em = {};
        if (c + 2 > n || match_ewmul(p, c, &em) != 2) break;
        const auto& fe = p->ops[c].u.fused_elementwise; const auto& me = p->ops[c + 1].u.elementwise;
        if (fe.src != mg.dst || me.src1 != mu.dst || fe.n != F || fe.n_steps > kZgDecMaxSteps) break;
        ly.n_steps = (uint32_t)fe.n_steps; ly.F = F; ly.silu = bufp(fe.dst); ly.hidden = bufp(me.dst);
        for (size_t k = 0; k < fe.n_steps; k++) {
            const ZgFusedEwStep& st = p->steps[c][k];
            ZgDecStep& d = ly.steps[k];
            d.op = st.op; d.is_swapped = st.is_swapped; d.sec_kind = 0; d.sec = nullptr;
            if (st.op == ZG_EW_ADD || st.op == ZG_EW_MUL) {
                if (st.secondary_buf == mg.dst && st.secondary_offset == 0) d.sec_kind = 1;
                else if (st.secondary_buf == mu.dst && st.secondary_offset == 0) d.sec_kind = 2;
                else {
                    if ((size_t)st.secondary_offset + F > elems(st.secondary_buf)) { heads_ok = false; break; }
                    d.sec = bufp(st.secondary_buf) + st.secondary_offset; external.push_back(st.secondary_buf);
                }
            }
        }
        if (!heads_ok) break;
        {   // nn.silu as the lowering emits it (src/nn.zig:38-44): neg, exp, + ones, recip, * gate
            const ZgDecStep* st = ly.steps;
            ly.act_silu = (ly.n_steps == 5 && st[0].op == ZG_EW_NEG && st[1].op == ZG_EW_EXP && st[2].op == ZG_EW_ADD && st[2].sec_kind == 0 &&
                           st[3].op == ZG_EW_RECIP && st[4].op == ZG_EW_MUL && st[4].sec_kind == 1) ? 1u : 0u;
        }
        roles.insert(roles.end(), {fe.dst, me.dst}); written.insert(written.end(), {fe.dst, me.dst});
        const uint32_t hidden_buf = me.dst;
        c += 2;
        // ── down projection [+ all-reduce] ──
        if (c >= n || p->ops[c].tag != ZG_OP_QMATMUL) break;
        const ZgOp* op_d = &p->ops[c];
        const auto& md = op_d->u.qmatmul;
        {
            const ZgCudaQWeight* w = p->qweights[md.weight_idx];
            if (md.M != 1 || md.input != hidden_buf || md.K != F || md.N != D || md.input_offset || md.dst_offset || w->fmt == ZG_QFMT_GENERIC) break;
        }
        c++;
        roles.push_back(md.dst); written.push_back(md.dst);
        bool ar_down = false;
        if (c < n && p->ops[c].tag == ZG_OP_ALLREDUCE) {
            const auto& ar = p->ops[c].u.allreduce;
            if (ar.buf != md.dst || ar.offset != 0 || ar.n != D || (p->ctx->world > 1 && !zg_peer_allreduce_ok(p->ctx, D))) break;
            ar_down = p->ctx->world > 1;
            c++;
        }
        {   // role buffers pairwise distinct (the kernel reorders the stores of a layer's small ops inside a phase)
            std::vector<uint32_t> v(roles);
            std::sort(v.begin(), v.end());
            if (std::adjacent_find(v.begin(), v.end()) != v.end()) break;
        }
        if (D % 4 || F % 4 || (D & 1)) break;
        // ── phases ──
        auto fill_mv = [&](ZgDecPhase& ph, std::initializer_list<const ZgOp*> mops, uint32_t K, bool whole, uint32_t* part_elems) {
            ph.n_mv = (uint32_t)mops.size(); ph.K = K;
            for (uint32_t m2 = 0; m2 < kZgDecMaxMv; m2++) ph.mv[m2].first_item = UINT32_MAX;   // absent entries never match an item
            uint32_t items = 0, k = 0;
            for (const ZgOp* mop : mops) {
                const auto& qm = mop->u.qmatmul;
                const ZgCudaQWeight* w = p->qweights[qm.weight_idx];
                ZgDecMv& m = ph.mv[k++];
                m.recs = w->recs; m.smax = w->smax; m.out = bufp(qm.dst); m.part = nullptr; m.n_nb = w->n_nb; m.first_item = items; m.N = (uint32_t)w->N;
                items += w->n_nb;
                ph.fmt = (uint32_t)w->fmt; ph.n_kc = w->n_kc; ph.rec_bytes = w->rec_bytes;
            }
            ph.n_items = items;
            if (!plan_decode_phase(grid, ph.n_kc, items, whole, &ph.S, &ph.lS, &ph.n_slots)) return false;
            *part_elems = 0;
            if (ph.S > 1) for (uint32_t m2 = 0; m2 < ph.n_mv; m2++) *part_elems += ph.S * ph.mv[m2].N;
            return true;
        };
        if (!fill_mv(L.ph[0], {op_q, op_k, op_v}, D, true, &L.part_elems[0]) || !fill_mv(L.ph[1], {op_o}, Dq, false, &L.part_elems[1]) ||
            !fill_mv(L.ph[2], {op_g, op_u}, D, true, &L.part_elems[2]) || !fill_mv(L.ph[3], {op_d}, F, false, &L.part_elems[3])) break;
        if (n_heads > kZgDecMaxHeads || L.kvs.size() > kZgDecMaxHeads) break;
        if ((size_t)L.ph[0].n_kc * ZG_KR > kZgDecMaxD) break;
        // vectors as their consumers see them (partial pointers are patched in once the scratch is allocated)
        auto vec_of = [&](const ZgDecPhase& ph, uint32_t k) { ZgDecVec v; v.full = ph.mv[k].out; v.part = nullptr; v.S = ph.S > 1 ? ph.S : 0; v.n = ph.mv[k].N; return v; };
        ly.q = vec_of(L.ph[0], 0); ly.k = vec_of(L.ph[0], 1); ly.v = vec_of(L.ph[0], 2);
        ly.o_local = vec_of(L.ph[1], 0); ly.o = ly.o_local;
        ly.gate = vec_of(L.ph[2], 0); ly.up = vec_of(L.ph[2], 1);
        ly.down_local = vec_of(L.ph[3], 0); ly.down = ly.down_local;
        ly.ar_o = ar_o ? 1u : 0u; ly.ar_down = ar_down ? 1u : 0u;
        if (ar_o) { ly.o.S = 0; ly.o.part = nullptr; }
        if (ar_down) { ly.down.S = 0; ly.down.part = nullptr; }
        ly.cs = bufp(cs_buf); ly.mask = ly.has_mask ? bufp(mask_buf) : nullptr; ly.k_cache = bufp(kc_buf); ly.v_cache = bufp(vc_buf);
        ly.attn_buf = bufp(cat_buf); ly.n_heads = n_heads; ly.n_kv = (uint32_t)L.kvs.size(); ly.d_head = dh;
        out.push_back(L);
        prev_after = after_buf; prev_down = md.dst; prev_down_vec = ly.down;
        i = c;
    }
    if (out.empty()) return false;
    // nothing the kernel treats as constant input may be written inside the range
    std::sort(written.begin(), written.end());
    for (uint32_t b : external) if (std::binary_search(written.begin(), written.end(), b)) { out.clear(); return false; }
    *end_out = i;
    return true;
}

// Allocate and upload the device tables of the fused decode kernel for the matched layers.
static bool build_decode_plan(ZgCudaProgram* p, std::vector<DecHostLayer>& layers) {
    ZgDecodeHost& d = p->dec;
    zg_decode_free(&d);
    const uint32_t grid = zg_decode_grid(p->ctx);
    const uint32_t L = (uint32_t)layers.size();
    uint32_t max_heads = 0, max_dh = 0; size_t part_max[4] = {0, 0, 0, 0};
    for (DecHostLayer& h : layers) {
        max_heads = std::max(max_heads, h.ly.n_heads); max_dh = std::max(max_dh, h.ly.d_head);
        for (int k = 0; k < 4; k++) part_max[k] = std::max(part_max[k], (size_t)h.part_elems[k]);
    }
    // split-K scratch: one region per phase kind, shared by all layers (a phase's partial sums are consumed before the same
    // phase of the next layer runs: at least two grid barriers lie in between)
    size_t part_off[4], part_total = 0;
    for (int k = 0; k < 4; k++) { part_off[k] = part_total; part_total += (part_max[k] + 31) & ~(size_t)31; }
    if (part_total) ZG_CUDA_OK(cudaMalloc(&d.d_part, part_total * sizeof(float)));
    uint32_t max_splits = grid / std::max(max_heads, 1u);
    max_splits = std::max(1u, std::min(max_splits, 16u));
    const uint32_t part_dh = (max_dh + 31) & ~31u;
    ZG_CUDA_OK(cudaMalloc(&d.d_attn_part, (size_t)max_heads * max_splits * (2 + part_dh) * sizeof(float)));
    ZG_CUDA_OK(cudaMalloc(&d.d_sync, 128 * sizeof(uint32_t)));
    ZG_CUDA_OK(cudaMemset(d.d_sync, 0, 128 * sizeof(uint32_t)));
    ZG_CUDA_OK(cudaMallocHost(&d.h_err, sizeof(uint32_t)));
    *d.h_err = 0;
    std::vector<ZgDecLayer> lys;
    for (DecHostLayer& h : layers) {
        for (int k = 0; k < 4; k++) {
            ZgDecPhase& ph = h.ph[k];
            if (ph.S > 1) {
                float* base = d.d_part + part_off[k];
                for (uint32_t m = 0; m < ph.n_mv; m++) { ph.mv[m].part = base; base += (size_t)ph.S * ph.mv[m].N; }
            }
        }
        ZgDecLayer& ly = h.ly;
        auto patch = [&](ZgDecVec& v, const ZgDecPhase& ph, uint32_t m) { if (v.S) v.part = ph.mv[m].part; };
        patch(ly.q, h.ph[0], 0); patch(ly.k, h.ph[0], 1); patch(ly.v, h.ph[0], 2);
        patch(ly.o_local, h.ph[1], 0); patch(ly.o, h.ph[1], 0);
        patch(ly.gate, h.ph[2], 0); patch(ly.up, h.ph[2], 1);
        patch(ly.down_local, h.ph[3], 0); patch(ly.down, h.ph[3], 0);
        ly.head0 = 0; ly.kv0 = 0;   // head / KV tables are per layer inside its descriptor block
        lys.push_back(ly);
    }
    for (uint32_t l = 1; l < L; l++) lys[l].x1_b = lys[l - 1].down;   // the previous layer's down projection, partial pointers now resolved
    const uint32_t blk = zg_decode_block_bytes();
    std::vector<uint8_t> blocks((size_t)L * blk);
    uint32_t cap_heads = 0, cap_kv = 0;
    for (uint32_t l = 0; l < L; l++) {
        zg_decode_block_fill(blocks.data() + (size_t)l * blk, lys[l], layers[l].ph, layers[l].heads.data(), (uint32_t)layers[l].heads.size(),
                             layers[l].kvs.data(), (uint32_t)layers[l].kvs.size());
        cap_heads = std::max(cap_heads, (uint32_t)layers[l].heads.size()); cap_kv = std::max(cap_kv, (uint32_t)layers[l].kvs.size());
    }
    if (cudaMalloc(&d.d_blocks, blocks.size()) != cudaSuccess || cudaMemcpy(d.d_blocks, blocks.data(), blocks.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        zg_set_error("decode plan: device table allocation failed"); zg_decode_free(&d); return false;
    }
    ZgDecodePlan& P = d.plan;
    P.blocks = (const uint8_t*)d.d_blocks; P.blk_bytes = blk; P.cap_heads = cap_heads; P.cap_kv = cap_kv; P.n_layers = L;
    P.attn_part = d.d_attn_part; P.sync = d.d_sync; P.dyn = nullptr;   // dyn: set at launch (the table is allocated after the schedule)
    P.max_splits = max_splits; P.part_dh = part_dh; P.grid = grid; P._pad = 0;
    P.pc = p->ctx->peer;
    d.valid = true;
    return true;
}

// Dependency levels over ITEMS (level = 1 + max level of every earlier item it conflicts with), then units in (level,
// program order); any such order is a valid topological order of the program's dependency DAG.
static bool build_schedule(ZgCudaProgram* p) {
    const size_t n = p->ops.size();
    const size_t chain_max = p->ctx->chain_max;
    std::vector<ZgItem> items;
    std::vector<ZgNormMacro> norm_of;     // per item (kind ITEM_NORM)
    std::vector<ZgEwMulMacro> ewmul_of;   // per item (kind ITEM_EWMUL)
    std::map<size_t, int> attn_of_item;   // item (kind ITEM_ATTN_LAYER) -> index into p->attn_blocks
    p->attn_blocks.clear();
    // single-token LLaMA layers: ONE persistent kernel for all of them (decode.cu)
    std::vector<DecHostLayer> dec_layers;
    size_t dec_end = 0;
    zg_decode_free(&p->dec);
    p->dec_first = 0; p->dec_count = 0;
    if (p->ctx->decode_fused && !p->kvq && p->ctx->fuse && chain_max && p->uniform_pos && match_decode_layers(p, 0, dec_layers, &dec_end)) {
        if (!build_decode_plan(p, dec_layers)) return false;
        p->dec_count = (uint32_t)dec_end;
    }
    for (size_t i = 0; i < n;) {
        if (p->dec_count && i == 0) {
            ZgItem it; it.first = 0; it.count = p->dec_count; it.kind = ITEM_DECODE;
            items.push_back(it); norm_of.push_back(ZgNormMacro{}); ewmul_of.push_back(ZgEwMulMacro{});
            i = p->dec_count;
            continue;
        }
        ZgItem it; it.first = (uint32_t)i;
        ZgNormMacro nm = {}; ZgEwMulMacro em = {};
        uint32_t c = 0;
        ZgAttnBlock ab;
        if (p->kvq && i < p->kvq_role.size() && p->kvq_role[i]) { it.kind = ITEM_KVQ; c = 1; }
        else if (p->ctx->fuse && p->ctx->attn_layer && !p->kvq && (c = match_attention_block(p, i, &ab))) { it.kind = ITEM_ATTN_LAYER; attn_of_item[items.size()] = (int)p->attn_blocks.size(); p->attn_blocks.push_back(ab); }
        else if (chain_max && p->ctx->fuse && (c = match_norm(p, i, &nm))) it.kind = ITEM_NORM;
        else if (p->ctx->fuse && (c = match_ewmul(p, i, &em))) it.kind = ITEM_EWMUL;
        else if (p->ctx->fuse && (c = match_attn_store(p, i))) it.kind = ITEM_ATTN_STORE;
        else c = 1;
        it.count = c;
        items.push_back(it); norm_of.push_back(nm); ewmul_of.push_back(em);
        i += c;
    }
    const size_t ni = items.size();
    // ── matvec prologue fusion: a norm block / SiLU*up pair whose result is consumed by the matvecs that follow it
    //    directly (q|k|v, gate|up, down) is evaluated inside those matvecs' prologues (ZgGemvPrologue); the first consumer
    //    also stores the absorbed ops' buffers.  Decode rows only (M <= 2). ──
    std::vector<ZgGemvPrologue> pro_of(ni);      // per consumer item
    std::vector<int> absorbed_by(ni, -1);        // macro item -> first consumer item
    std::vector<int> macro_of(ni, -1);           // consumer item -> macro item
    // ── gate | up pair + activation epilogue: qmatmul(gate), qmatmul(up), fused_elementwise(steps(gate)), mul(. * up) of a
    //    single-token program become ONE launch (the pair kernel evaluates the chain on the finished sums) ──
    std::vector<char> pair_lead(ni, 0), pair_absorbed(ni, 0), arn_lead(ni, 0);
    std::vector<ZgGemvEpilogue> epi_of(ni);
    // ── all-reduce + the norm block right behind it (every sharded layer has two): ONE launch ──
    if (p->ctx->fuse && p->ctx->ar_norm && p->ctx->world > 1) {
        for (size_t k = 0; k + 1 < ni; k++) {
            if (items[k].kind != ITEM_OP || items[k + 1].kind != ITEM_NORM) continue;
            const ZgOp& ao = p->ops[items[k].first];
            if (ao.tag != ZG_OP_ALLREDUCE) continue;
            const auto& ar = ao.u.allreduce;
            const ZgNormMacro& nm = norm_of[k + 1];
            float* ptr = p->buffers[ar.buf];
            if (ar.offset || (ar.n & 1) || !zg_peer_allreduce_ok(p->ctx, ar.n) || (((size_t)ptr) & 15) || nm.rows != 1 || nm.cols != ar.n) continue;
            if (nm.b ? (nm.a != ptr && nm.b != ptr) || nm.a == nm.b : nm.a != ptr) continue;
            arn_lead[k] = 1; pair_absorbed[k + 1] = 1;
        }
    }
    if (p->ctx->fuse && p->ctx->gemv_pair && !p->ctx->gemv_fuse) {
        for (size_t k = 0; k + 2 < ni; k++) {
            const ZgItem &ia = items[k], &ib = items[k + 1], &ie = items[k + 2];
            if (ia.kind != ITEM_OP || ib.kind != ITEM_OP || ie.kind != ITEM_EWMUL) continue;
            const ZgOp &oa = p->ops[ia.first], &ob = p->ops[ib.first];
            if (oa.tag != ZG_OP_QMATMUL || ob.tag != ZG_OP_QMATMUL) continue;
            const auto& qa = oa.u.qmatmul; const auto& qb = ob.u.qmatmul;
            const ZgCudaQWeight *wa = p->qweights[qa.weight_idx], *wb = p->qweights[qb.weight_idx];
            if (qa.M != 1 || qb.M != 1 || qa.input != qb.input || qa.input_offset != qb.input_offset || qa.dst_offset || qb.dst_offset || qa.dst == qb.dst ||
                wa->fmt == ZG_QFMT_GENERIC || wa->fmt != wb->fmt || wa->K != wb->K || wa->N != wb->N) continue;
            if (zg_qgemv_stream_plan(p->ctx, wa, 2).use) continue;   // large pairs: the streamed batch + a chain launch is faster than the pair kernel
            const ZgEwMulMacro& em = ewmul_of[k + 2];
            if (em.src != p->buffers[qa.dst] || em.other != p->buffers[qb.dst] || em.n != wa->N || em.n_steps > 6 ||
                em.mid == p->buffers[qa.input] || em.dst == p->buffers[qa.input]) continue;   // other CTAs still read the input vector
            ZgGemvEpilogue ep;
            ep.n_steps = em.n_steps; ep.o_mid = em.mid; ep.o_dst = em.dst;
            const auto& st = p->steps[ie.first];
            bool ok = true;
            for (uint32_t q = 0; q < em.n_steps && ok; q++) {
                ep.steps[q].op = st[q].op; ep.steps[q].is_swapped = st[q].is_swapped; ep.steps[q].sec = nullptr; ep.sec_kind[q] = 0;
                if (st[q].op == ZG_EW_ADD || st[q].op == ZG_EW_MUL) {
                    if (st[q].secondary_buf == qa.dst) { ep.sec_kind[q] = 1; ok = st[q].secondary_offset == 0; }
                    else if (st[q].secondary_buf == qb.dst) { ep.sec_kind[q] = 2; ok = st[q].secondary_offset == 0; }
                    else { ep.steps[q].sec = p->buffers[st[q].secondary_buf] + st[q].secondary_offset; ok = (size_t)st[q].secondary_offset + em.n <= p->buffer_elems[st[q].secondary_buf]; }
                }
            }
            if (!ok) continue;
            pair_lead[k] = 1; pair_absorbed[k + 1] = 1; pair_absorbed[k + 2] = 1; epi_of[k] = ep;
            k += 2;
        }
    }
    if (p->ctx->fuse && p->ctx->gemv_fuse) {
        for (size_t k = 0; k + 1 < ni; k++) {
            const ZgItem& it = items[k];
            if (it.kind != ITEM_NORM && it.kind != ITEM_EWMUL) continue;
            if (!(p->ctx->gemv_fuse & (it.kind == ITEM_NORM ? 1 : 2))) continue;
            const ZgOp& last = p->ops[it.first + it.count - 1];           // the closing elementwise mul: its dst is the matvec input
            const uint32_t out_buf = last.u.elementwise.dst;
            ZgGemvPrologue pr;
            uint32_t rows, cols;
            std::vector<uint32_t> touched;                                  // buffers the recipe reads or writes
            if (it.kind == ITEM_NORM) {
                const ZgNormMacro& m = norm_of[k];
                if (m.rows > 2) continue;
                rows = m.rows; cols = m.cols;
                pr.kind = 1; pr.eps = m.eps; pr.a = m.a; pr.b = m.b; pr.gamma = m.gamma;
                pr.o_sum = m.sum; pr.o_mid = m.bare; pr.o_grep = m.gamma_rep; pr.o_x = m.norm;
            } else {
                const ZgEwMulMacro& m = ewmul_of[k];
                if (m.n_steps > 6) continue;
                rows = 0; cols = m.n;                                       // rows fixed by the consumer: n = M * K
                pr.kind = 2; pr.n_steps = m.n_steps; pr.a = m.src; pr.b = m.other; pr.o_mid = m.mid; pr.o_x = m.dst;
                const auto& st = p->steps[it.first];
                for (uint32_t q = 0; q < m.n_steps; q++) {
                    pr.steps[q].op = st[q].op; pr.steps[q].is_swapped = st[q].is_swapped;
                    pr.steps[q].sec = (st[q].op == ZG_EW_ADD || st[q].op == ZG_EW_MUL) ? p->buffers[st[q].secondary_buf] + st[q].secondary_offset : nullptr;
                }
            }
            for (uint32_t j = 0; j < it.count; j++) {
                std::vector<ZgRange> rr; op_ranges(p, p->ops[it.first + j], rr);
                for (const ZgRange& r : rr) touched.push_back(r.buf);
            }
            size_t j = k + 1;
            for (; j < ni; j++) {
                const ZgItem& c = items[j];
                if (c.kind != ITEM_OP || p->ops[c.first].tag != ZG_OP_QMATMUL) break;
                const auto& q = p->ops[c.first].u.qmatmul;
                const ZgCudaQWeight* w = p->qweights[q.weight_idx];
                if (q.input != out_buf || q.input_offset != 0 || (q.input_row_stride != 0 && q.input_row_stride != q.K) || q.M == 0 || q.M > 2 ||
                    w->fmt == ZG_QFMT_GENERIC) break;
                if (it.kind == ITEM_NORM ? (q.M != rows || q.K != cols) : ((size_t)q.M * q.K != cols)) break;
                if (std::find(touched.begin(), touched.end(), q.dst) != touched.end()) break;   // the matvec must not overwrite a recipe buffer
                pro_of[j] = pr;
                pro_of[j].write = (absorbed_by[k] < 0) ? 1u : 0u;
                macro_of[j] = (int)k;
                if (absorbed_by[k] < 0) absorbed_by[k] = (int)j;
            }
        }
    }
    // element ranges of every item (absorbed macro: none of its own; first consumer: macro + matvec; other consumers:
    // the matvec with its activation read replaced by reads of the recipe's inputs, so they do not wait for the writer)
    std::vector<std::vector<ZgRange>> item_rng(ni);
    {
        std::vector<ZgRange> tmp;
        for (size_t k = 0; k < ni; k++) {
            if (absorbed_by[k] >= 0 || pair_absorbed[k]) continue;
            for (uint32_t j = 0; j < items[k].count; j++) { op_ranges(p, p->ops[items[k].first + j], tmp); item_rng[k].insert(item_rng[k].end(), tmp.begin(), tmp.end()); }
            if (pair_lead[k] || arn_lead[k])   // the launch also runs the up matvec and the activation chain / the norm block
                for (size_t k2 = k + 1; k2 <= k + (pair_lead[k] ? 2 : 1); k2++)
                    for (uint32_t j = 0; j < items[k2].count; j++) { op_ranges(p, p->ops[items[k2].first + j], tmp); item_rng[k].insert(item_rng[k].end(), tmp.begin(), tmp.end()); }
            if (macro_of[k] < 0) continue;
            const ZgItem& mi = items[macro_of[k]];
            const uint32_t in_buf = p->ops[items[k].first].u.qmatmul.input;
            std::vector<ZgRange> mr;
            for (uint32_t j = 0; j < mi.count; j++) { op_ranges(p, p->ops[mi.first + j], tmp); mr.insert(mr.end(), tmp.begin(), tmp.end()); }
            if (pro_of[k].write) {
                item_rng[k].insert(item_rng[k].end(), mr.begin(), mr.end());
            } else {
                std::vector<ZgRange> keep;
                for (const ZgRange& r : item_rng[k]) if (r.write || r.buf != in_buf) keep.push_back(r);
                std::vector<uint32_t> written;
                for (const ZgRange& r : mr) if (r.write) written.push_back(r.buf);
                for (const ZgRange& r : mr)   // the macro's external inputs: read ranges of buffers the macro does not itself produce
                    if (!r.write && std::find(written.begin(), written.end(), r.buf) == written.end()) keep.push_back(r);
                item_rng[k] = keep;
            }
        }
    }
    struct Access { ZgRange r; int level; };
    std::vector<std::vector<Access>> acc(p->buffers.size() + 2);   // + the virtual GEMM-scratch and communicator buffers
    std::vector<int> level(ni, 0);
    for (size_t k = 0; k < ni; k++) {
        if (absorbed_by[k] >= 0 || pair_absorbed[k]) { level[k] = -1; continue; }
        const std::vector<ZgRange>& rng = item_rng[k];
        int lvl = 0;
        for (const ZgRange& r : rng)
            for (const Access& a : acc[r.buf])
                if (a.level + 1 > lvl && range_conflict(r, a.r)) lvl = a.level + 1;
        level[k] = lvl;
        for (const ZgRange& r : rng) {
            // a write covering the whole buffer orders everything after it: older accesses need not be kept
            if (r.write && !r.dyn && !r.stride && r.buf < p->buffer_elems.size() && r.lo == 0 && r.hi >= p->buffer_elems[r.buf]) acc[r.buf].clear();
            acc[r.buf].push_back({r, lvl});
        }
    }
    std::vector<uint32_t> order;
    for (size_t k = 0; k < ni; k++) if (absorbed_by[k] < 0 && !pair_absorbed[k]) order.push_back((uint32_t)k);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return level[a] < level[b]; });
    const size_t no = order.size();
    p->units.clear();
    p->entry_of_op.assign(n, 0);
    std::vector<ZgBatchEntry> entries;
    std::vector<ZgChainOp> chain_ops;
    // Chains: walking the levels upwards, small items join the open chain (first item of a new level = block barrier).
    // A level that also holds big ops (matvecs, attention, dense matmul, NCCL collectives) closes the chain BEFORE
    // them: they may consume the chain's earlier levels, and nothing in the chain depends on them (same level = no
    // conflict; later levels start a new chain).  The unit order stays a topological order of the dependency DAG.
    auto chain_work = [&](const ZgItem& it) -> size_t {
        static const uint32_t norm_chain_max = [] { const char* e = getenv("ZG_CUDA_NORM_CHAIN_MAX"); return e ? (uint32_t)atoi(e) : 256u; }();   // measured: the cluster kernel beats the chain CTA from SmolLM-135M (576) up: +4 % / +2 % / +7 % tok/s on 135M / 1.7B / Llama-3-8B
        if (it.kind == ITEM_NORM) return norm_of[&it - items.data()].cols <= norm_chain_max ? 1 : 0;   // longer rows: their own cluster kernel (ops.cu k_norm_macro_cluster)
        if (it.kind == ITEM_EWMUL) return ewmul_of[&it - items.data()].n <= 1024 ? 1 : 0;   // transcendental chains: one CTA only when tiny
        if (it.kind != ITEM_OP) return 0;
        const ZgOp& op = p->ops[it.first];
        if (op.tag == ZG_OP_ALLREDUCE) return 0;   // its own multi-CTA kernel (peer memory) or NCCL
        const size_t wk = zg_chain_work(op);
        return wk <= chain_max ? wk : 0;
    };
    std::vector<uint32_t> chain_items;
    std::vector<char> chain_sync, chain_tiny;
    auto close_chain = [&]() {
        if (chain_items.empty()) return true;
        const ZgItem& only = items[chain_items[0]];
        if (chain_items.size() == 1 && only.kind == ITEM_OP && !zg_op_is_batched(p->ops[only.first].tag)) {
            ZgCudaProgram::Unit u; u.ops.push_back(only.first); p->units.push_back(u);   // a lone op: its own (wider) kernel is as good
        } else {
            ZgCudaProgram::Unit chain; chain.chain = true;
            chain.first_entry = (uint32_t)chain_ops.size();
            chain.n_entries = (uint32_t)chain_items.size();
            for (size_t k = 0; k < chain_items.size(); k++) {
                ZgChainOp c;
                const ZgItem& it = items[chain_items[k]];
                const bool sync = chain_sync[k] != 0;
                bool ok;
                if (it.kind == ITEM_NORM) ok = zg_fill_chain_norm(norm_of[chain_items[k]], sync, &c);
                else if (it.kind == ITEM_EWMUL) ok = zg_fill_chain_ewmul(ewmul_of[chain_items[k]], sync, &c);
                else ok = zg_fill_chain_op(p->ops[it.first], p->buffers.data(), it.first, p->d_steps + p->step_off[it.first], sync, &c);
                if (!ok) return false;
                chain_ops.push_back(c);
                for (uint32_t j = 0; j < it.count; j++) chain.ops.push_back(it.first + j);
            }
            // runs of >= 2 tiny ops inside one level execute warp-parallel (group = run length on the run's first op)
            for (size_t k = 0; k < chain_items.size();) {
                size_t e = k;
                if (chain_tiny[k]) { e = k + 1; while (e < chain_items.size() && chain_tiny[e] && !chain_sync[e]) e++; }
                if (e - k >= 2) { chain_ops[chain.first_entry + k].group = (uint32_t)(e - k); k = e; }
                else k++;
            }
            p->units.push_back(chain);
        }
        chain_items.clear(); chain_sync.clear(); chain_tiny.clear();
        return true;
    };
    auto is_tiny = [&](const ZgItem& it) {
        if (it.kind != ITEM_OP) return false;
        const ZgOp& op = p->ops[it.first];
        return op.tag != ZG_OP_RMSNORM && op.tag != ZG_OP_ALLREDUCE && zg_chain_work(op) <= 512;
    };
    std::vector<int> level_of_op(n, -1);
    for (size_t k = 0; k < ni; k++) for (uint32_t j = 0; j < items[k].count; j++) level_of_op[items[k].first + j] = level[k];
    auto level_of_unit_first = [&](const ZgCudaProgram::Unit& u) { return level_of_op[u.ops[0]]; };
    size_t pos = 0;
    while (pos < no) {
        size_t end = pos;
        while (end < no && level[order[end]] == level[order[pos]]) end++;
        bool first_small = true, has_big = false;
        // per-head ops (rope, cache stores): one warp each inside a chain is only a win for a handful of heads; a wide
        // level (32 heads x 4 ops) is better served by the batched multi-CTA kernels
        size_t n_tiny = 0;
        for (size_t k = pos; k < end; k++) n_tiny += (chain_max && chain_work(items[order[k]]) && is_tiny(items[order[k]])) ? 1 : 0;
        const bool tiny_in_chain = n_tiny <= 8;
        auto in_chain = [&](const ZgItem& it) { return chain_max && chain_work(it) != 0 && (tiny_in_chain || !is_tiny(it)); };
        for (int pass = 0; pass < 2; pass++) {   // CTA-wide ops first, then the level's tiny ops as one warp-parallel run
            for (size_t k = pos; k < end; k++) {
                const ZgItem& it = items[order[k]];
                if (!in_chain(it)) { has_big = true; continue; }
                const bool tiny = is_tiny(it);
                if (tiny != (pass == 1)) continue;
                if (chain_items.size() >= kZgChainMaxOps) { if (!close_chain()) return false; }   // kernel boundary = barrier
                chain_sync.push_back(first_small && !chain_items.empty());
                chain_tiny.push_back(tiny);
                chain_items.push_back(order[k]);
                first_small = false;
            }
        }
        if (has_big && !close_chain()) return false;
        std::vector<std::pair<uint64_t, size_t>> open;   // batch signature -> unit index (within this level)
        std::vector<std::pair<const ZgCudaQWeight*, size_t>> open_mv;   // representative weight -> open matvec batch
        for (size_t k = pos; k < end; k++) {
            const ZgItem& it = items[order[k]];
            if (in_chain(it)) continue;   // chained above
            const ZgOp& op = p->ops[it.first];
            if (arn_lead[order[k]]) {
                ZgCudaProgram::Unit u; u.ar_norm = true; u.nm = norm_of[order[k] + 1];
                for (size_t k2 = order[k]; k2 <= (size_t)order[k] + 1; k2++)
                    for (uint32_t j = 0; j < items[k2].count; j++) u.ops.push_back(items[k2].first + j);
                u.ranges = item_rng[order[k]];
                p->units.push_back(u); continue;
            }
            if (pair_lead[order[k]]) {
                ZgCudaProgram::Unit u; u.gemv_pair = true; u.epi = epi_of[order[k]];
                u.entry_ops.push_back(it.first); u.entry_ops.push_back(items[order[k] + 1].first);
                for (size_t k2 = order[k]; k2 <= (size_t)order[k] + 2; k2++)
                    for (uint32_t j = 0; j < items[k2].count; j++) u.ops.push_back(items[k2].first + j);
                u.ranges = item_rng[order[k]];
                p->units.push_back(u); continue;
            }
            if (it.kind == ITEM_OP && op.tag == ZG_OP_QMATMUL && op.u.qmatmul.M >= 1 && op.u.qmatmul.M <= 8 &&
                (p->ctx->gemv_batch > 1 || pro_of[order[k]].kind != 0) && p->qweights[op.u.qmatmul.weight_idx]->fmt != ZG_QFMT_GENERIC) {
                // independent matvecs of one shape / format / row count (and prologue kind) share a launch (q|k|v, gate|up)
                const ZgCudaQWeight* w = p->qweights[op.u.qmatmul.weight_idx];
                const uint32_t pk = pro_of[order[k]].kind;
                size_t ui = (size_t)-1;
                for (auto& o : open_mv) {
                    const ZgCudaQWeight* r = o.first;
                    const ZgCudaProgram::Unit& u = p->units[o.second];
                    if (r->fmt == w->fmt && r->K == w->K && r->N == w->N && p->ops[u.entry_ops[0]].u.qmatmul.M == op.u.qmatmul.M &&
                        u.pros[0].kind == pk && u.entry_ops.size() < (size_t)std::max(p->ctx->gemv_batch, 1)) { ui = o.second; break; }
                }
                if (ui == (size_t)-1) {
                    ZgCudaProgram::Unit u; u.gemv_batch = true;
                    p->units.push_back(u);
                    ui = p->units.size() - 1;
                    open_mv.push_back({w, ui});
                }
                ZgCudaProgram::Unit& u = p->units[ui];
                u.entry_ops.push_back(it.first);
                u.ops.push_back(it.first);
                u.pros.push_back(pro_of[order[k]]);
                if (pk && pro_of[order[k]].write) {   // the absorbed ops belong to this launch
                    const ZgItem& mi = items[macro_of[order[k]]];
                    for (uint32_t j = 0; j < mi.count; j++) u.ops.push_back(mi.first + j);
                }
                u.ranges.insert(u.ranges.end(), item_rng[order[k]].begin(), item_rng[order[k]].end());
                continue;
            }
            if (it.kind == ITEM_KVQ) {   // same-level cache stores / cache-backed attentions of one shape share a launch
                const uint32_t role = (uint32_t)p->kvq_role[it.first];
                const uint32_t sq = role == 2 ? op.u.attention.seq_q : 0;
                size_t ui = (size_t)-1;
                for (size_t q = p->units.size(); q-- > 0;) {
                    const ZgCudaProgram::Unit& u = p->units[q];
                    if (u.kvq == role && u.kvq_seq_q == sq && !u.ops.empty() && level[order[k]] == level_of_unit_first(u)) { ui = q; break; }
                    if (!u.kvq) break;
                }
                if (ui == (size_t)-1) { ZgCudaProgram::Unit u; u.kvq = role; u.kvq_seq_q = sq; p->units.push_back(u); ui = p->units.size() - 1; }
                p->units[ui].ops.push_back(it.first);
                continue;
            }
            if (it.kind == ITEM_ATTN_LAYER) {
                ZgCudaProgram::Unit u; u.attn_blk = attn_of_item[order[k]];
                for (uint32_t j = 0; j < it.count; j++) u.ops.push_back(it.first + j);
                p->units.push_back(u); continue;
            }
            if (it.kind == ITEM_DECODE) {
                ZgCudaProgram::Unit u; u.decode = true;
                for (uint32_t j = 0; j < it.count; j++) u.ops.push_back(it.first + j);
                p->units.push_back(u); continue;
            }
            if (it.kind == ITEM_NORM) {    // long rows: one wide CTA per row instead of the 256-thread chain
                ZgCudaProgram::Unit u;
                for (uint32_t j = 0; j < it.count; j++) u.ops.push_back(it.first + j);
                u.norm = true; u.nm = norm_of[order[k]];
                p->units.push_back(u); continue;
            }
            if (it.kind == ITEM_EWMUL) {   // fused_elementwise + mul as one multi-CTA launch
                ZgCudaProgram::Unit u; u.ops.push_back(it.first); u.ops.push_back(it.first + 1);
                u.ewmul = true; u.em = ewmul_of[order[k]];
                p->units.push_back(u); continue;
            }
            if (!zg_op_is_batched(op.tag)) { ZgCudaProgram::Unit u; u.ops.push_back(it.first); p->units.push_back(u); continue; }
            const uint64_t sig = zg_batch_signature(op);
            size_t ui = (size_t)-1;
            for (auto& o : open) if (o.first == sig) { ui = o.second; break; }
            if (ui == (size_t)-1) {
                ZgCudaProgram::Unit u; u.batched = true;
                p->units.push_back(u);
                ui = p->units.size() - 1;
                open.push_back({sig, ui});
            }
            p->units[ui].entry_ops.push_back(it.first);
            for (uint32_t j = 0; j < it.count; j++) p->units[ui].ops.push_back(it.first + j);
            if (it.kind == ITEM_ATTN_STORE) p->units[ui].store_of[it.first] = it.first + 1;
        }
        pos = end;
    }
    if (!close_chain()) return false;
    for (auto& u : p->units) {
        if (!u.batched) continue;
        u.first_entry = (uint32_t)entries.size();
        u.n_entries = (uint32_t)u.entry_ops.size();
        for (uint32_t i : u.entry_ops) {
            ZgBatchEntry e;
            if (!zg_fill_batch_entry(p->ops[i], p->buffers.data(), i, &e)) return false;
            auto st = u.store_of.find(i);
            if (st != u.store_of.end()) {   // the absorbed slice_assign: second destination of the attention output
                const auto& sa = p->ops[st->second].u.slice_assign;
                e.dst2 = p->buffers[sa.dst]; e.d2_off = sa.dst_offset; e.d2_rs = sa.dst_row_stride; e.d2_cs = sa.dst_col_stride;
            }
            p->entry_of_op[i] = (uint32_t)entries.size();
            entries.push_back(e);
        }
    }
    // cache-backed ops: parameter tables in unit order
    cudaFree(p->d_kvq_store); p->d_kvq_store = nullptr;
    cudaFree(p->d_kvq_attn); p->d_kvq_attn = nullptr;
    p->kvq_entry.assign(n, 0);
    if (p->kvq) {
        std::vector<ZgKvqStore> st_tab; std::vector<ZgKvqAttn> at_tab;
        for (auto& u : p->units) {
            if (!u.kvq) continue;
            u.first_entry = (uint32_t)(u.kvq == 1 ? st_tab.size() : at_tab.size());
            u.n_entries = (uint32_t)u.ops.size();
            for (uint32_t i : u.ops) {
                int8_t* cq; float* cs; uint32_t dh, bs, bpc; size_t ncols;
                if (u.kvq == 1) {
                    const auto& sa = p->ops[i].u.slice_assign;
                    zg_kvq_cache_arrays(p->kv_caches[sa.dst], &cq, &cs, &dh, &bs, &bpc, &ncols);
                    ZgKvqStore e;
                    e.src = p->buffers[sa.src] + sa.src_offset; e.q = cq; e.s = cs; e.src_cs = sa.src_col_stride; e.n_write = sa.cols;
                    e.dyn_idx = i; e.d_head = dh; e.bs = bs; e.bpc = bpc;
                    u.kvq_max_warps = std::max(u.kvq_max_warps, sa.cols * bpc);
                    p->kvq_entry[i] = (uint32_t)st_tab.size();
                    st_tab.push_back(e);
                } else {
                    const auto& a = p->ops[i].u.attention;
                    ZgKvqAttn e;
                    memset(&e, 0, sizeof(e));
                    e.dst = p->buffers[a.dst] + a.dst_off; e.dst_cs = a.dst_cs; e.q = p->buffers[a.q] + a.q_off; e.q_cs = a.q_cs;
                    e.d_head = a.d_head; e.seq_kv = a.seq_kv;
                    zg_kvq_cache_arrays(p->kv_caches[a.k], &cq, &cs, &dh, &bs, &bpc, &ncols);
                    e.k_q = cq; e.k_s = cs; e.k_col_start = a.k_off / a.d_head; e.bs = bs; e.nb = bpc;
                    zg_kvq_cache_arrays(p->kv_caches[a.v], &cq, &cs, &dh, &bs, &bpc, &ncols);
                    e.v_q = cq; e.v_s = cs; e.v_col_start = a.v_off / a.d_head;
                    e.mask = a.has_mask ? p->buffers[a.mask] + a.mask_off : nullptr; e.mask_rs = a.mask_rs; e.mask_cs = a.seq_q > 1 ? a.mask_cs : 0;
                    e.scale = a.scale; e.int8_query = p->kvq_int8; e.dyn_idx = i;
                    p->kvq_entry[i] = (uint32_t)at_tab.size();
                    at_tab.push_back(e);
                }
            }
        }
        if (!st_tab.empty()) {
            ZG_CUDA_OK(cudaMalloc(&p->d_kvq_store, st_tab.size() * sizeof(ZgKvqStore)));
            ZG_CUDA_OK(cudaMemcpy(p->d_kvq_store, st_tab.data(), st_tab.size() * sizeof(ZgKvqStore), cudaMemcpyHostToDevice));
        }
        if (!at_tab.empty()) {
            ZG_CUDA_OK(cudaMalloc(&p->d_kvq_attn, at_tab.size() * sizeof(ZgKvqAttn)));
            ZG_CUDA_OK(cudaMemcpy(p->d_kvq_attn, at_tab.data(), at_tab.size() * sizeof(ZgKvqAttn), cudaMemcpyHostToDevice));
        }
        // split-KV: several CTAs per query column when the launch would otherwise leave most SMs idle (decode: one column per head)
        cudaFree(p->d_kvq_part); p->d_kvq_part = nullptr;
        cudaFree(p->d_kvq_cnt); p->d_kvq_cnt = nullptr;
        size_t part_total = 0, cnt_total = 0;
        for (auto& u : p->units) {
            if (u.kvq != 2) continue;
            const uint32_t rows = u.n_entries * u.kvq_seq_q;
            uint32_t sp = rows ? (uint32_t)p->ctx->sm_count * 2 / rows : 1;
            sp = std::max(1u, std::min(sp, 16u));
            uint32_t dmax = 0;
            for (uint32_t i : u.ops) dmax = std::max(dmax, p->ops[i].u.attention.d_head);
            u.kvq_splits = sp; u.kvq_part_off = part_total; u.kvq_cnt_off = cnt_total;
            part_total += (size_t)rows * sp * (2 + dmax);
            cnt_total += rows;
        }
        if (part_total) {
            ZG_CUDA_OK(cudaMalloc(&p->d_kvq_part, part_total * sizeof(float)));
            ZG_CUDA_OK(cudaMalloc(&p->d_kvq_cnt, cnt_total * sizeof(uint32_t)));
            ZG_CUDA_OK(cudaMemset(p->d_kvq_cnt, 0, cnt_total * sizeof(uint32_t)));
        }
    }
    // fused attention blocks: descriptors + split-KV scratch
    cudaFree(p->d_attn_blocks); p->d_attn_blocks = nullptr;
    cudaFree(p->d_attn_blk_part); p->d_attn_blk_part = nullptr;
    cudaFree(p->d_attn_blk_cnt); p->d_attn_blk_cnt = nullptr;
    if (!p->attn_blocks.empty()) {
        size_t pt = 0, ct = 0;
        for (auto& u : p->units) {
            if (u.attn_blk < 0) continue;
            const ZgAttnBlock& B = p->attn_blocks[u.attn_blk];
            static const uint32_t mult = [] { const char* e = getenv("ZG_CUDA_ATTN_SPLIT_MULT"); return e ? (uint32_t)atoi(e) : 1u; }();
            uint32_t sp = (uint32_t)p->ctx->sm_count * std::max(mult, 1u) / std::max(B.n_heads, 1u);
            sp = std::max(1u, std::min(sp, 8u));
            const uint32_t ni32 = B.d_head <= 64 ? 64 : (B.d_head <= 128 ? 128 : 256);
            u.ab_splits = sp; u.ab_part_off = pt; u.ab_cnt_off = ct;
            pt += (size_t)B.n_heads * sp * (ni32 + 2);
            ct += B.n_heads;
        }
        ZG_CUDA_OK(cudaMalloc(&p->d_attn_blocks, p->attn_blocks.size() * sizeof(ZgAttnBlock)));
        ZG_CUDA_OK(cudaMemcpy(p->d_attn_blocks, p->attn_blocks.data(), p->attn_blocks.size() * sizeof(ZgAttnBlock), cudaMemcpyHostToDevice));
        ZG_CUDA_OK(cudaMalloc(&p->d_attn_blk_part, std::max<size_t>(pt, 1) * sizeof(float)));
        ZG_CUDA_OK(cudaMalloc(&p->d_attn_blk_cnt, std::max<size_t>(ct, 1) * sizeof(uint32_t)));
        ZG_CUDA_OK(cudaMemset(p->d_attn_blk_cnt, 0, std::max<size_t>(ct, 1) * sizeof(uint32_t)));
    }
    // split-KV scratch of the decode attention units
    size_t part_total = 0, cnt_total = 0;
    for (auto& u : p->units) {
        if (!u.batched || u.entry_ops.empty() || p->ops[u.entry_ops[0]].tag != ZG_OP_ATTENTION || !p->ctx->attn_split) continue;
        const ZgOp& a0 = p->ops[u.entry_ops[0]];
        size_t kmin = (size_t)-1;
        for (uint32_t i : u.entry_ops) kmin = std::min(kmin, p->buffer_elems[p->ops[i].u.attention.k] - std::min<size_t>(p->buffer_elems[p->ops[i].u.attention.k], p->ops[i].u.attention.k_off) + a0.u.attention.k_off);
        u.attn_splits = zg_attention_splits(a0, kmin, u.n_entries, p->ctx->sm_count);
        if (u.attn_splits <= 1) { u.attn_splits = 1; continue; }
        u.attn_part_off = part_total; u.attn_cnt_off = cnt_total;
        part_total += zg_attention_part_elems(a0, u.n_entries, u.attn_splits);
        cnt_total += (size_t)u.n_entries * a0.u.attention.seq_q;
    }
    cudaFree(p->d_attn_part); p->d_attn_part = nullptr;
    cudaFree(p->d_attn_cnt); p->d_attn_cnt = nullptr;
    if (part_total) {
        ZG_CUDA_OK(cudaMalloc(&p->d_attn_part, part_total * sizeof(float)));
        ZG_CUDA_OK(cudaMalloc(&p->d_attn_cnt, cnt_total * sizeof(uint32_t)));
        ZG_CUDA_OK(cudaMemset(p->d_attn_cnt, 0, cnt_total * sizeof(uint32_t)));
    }
    // single-op launches (profiling / eager per-op mode) of batched kinds need their own plain entries
    p->single_entry.assign(n, 0);
    for (size_t i = 0; i < n; i++) {
        if (!zg_op_is_batched(p->ops[i].tag)) continue;
        ZgBatchEntry e;
        if (!zg_fill_batch_entry(p->ops[i], p->buffers.data(), (uint32_t)i, &e)) return false;
        p->single_entry[i] = (uint32_t)entries.size();
        entries.push_back(e);
    }
    cudaFree(p->d_chain); p->d_chain = nullptr;
    if (!chain_ops.empty()) {
        ZG_CUDA_OK(cudaMalloc(&p->d_chain, chain_ops.size() * sizeof(ZgChainOp)));
        ZG_CUDA_OK(cudaMemcpy(p->d_chain, chain_ops.data(), chain_ops.size() * sizeof(ZgChainOp), cudaMemcpyHostToDevice));
    }
    cudaFree(p->d_batch); p->d_batch = nullptr;
    if (!entries.empty()) {
        ZG_CUDA_OK(cudaMalloc(&p->d_batch, entries.size() * sizeof(ZgBatchEntry)));
        ZG_CUDA_OK(cudaMemcpy(p->d_batch, entries.data(), entries.size() * sizeof(ZgBatchEntry), cudaMemcpyHostToDevice));
    }
    return true;
}

static bool launch_one(ZgCudaProgram* p, size_t i, cudaStream_t st) {
    ZgCudaCtx* ctx = p->ctx;
    const ZgOp& op = p->ops[i];
    if (p->kvq && i < p->kvq_role.size() && p->kvq_role[i] == 1) {
        const auto& sa = op.u.slice_assign;
        return zg_kvq_launch_stores(p->d_kvq_store + p->kvq_entry[i], 1, sa.cols * (uint32_t)(sa.rows / p->kvq_bs), p->d_dyn, st);
    }
    if (p->kvq && i < p->kvq_role.size() && p->kvq_role[i] == 2) return zg_kvq_launch_attention(p->d_kvq_attn + p->kvq_entry[i], 1, op.u.attention.seq_q, p->d_dyn, nullptr, nullptr, 1, st);
    if (op.tag == ZG_OP_QMATMUL) {
        const auto& q = op.u.qmatmul;
        ZgGemvWs view = p->ws;
        if (view.partials) { view.partials += p->ws_part_off[i]; view.partials_elems -= p->ws_part_off[i]; }
        if (view.counters) { view.counters += p->ws_cnt_off[i]; view.counters_n -= p->ws_cnt_off[i]; }
        return zg_qmatmul_launch(ctx, p->qweights[q.weight_idx], p->buffers[q.input] + q.input_offset,
                                 p->buffers[q.dst] + q.dst_offset, q.M, q.input_row_stride, q.dst_row_stride, &view, st);
    }
    if (op.tag == ZG_OP_ALLREDUCE && zg_peer_allreduce_ok(ctx, op.u.allreduce.n) && (((size_t)(p->buffers[op.u.allreduce.buf] + op.u.allreduce.offset)) & 15) == 0)
        return zg_launch_peer_allreduce(p->buffers[op.u.allreduce.buf] + op.u.allreduce.offset, op.u.allreduce.n, ctx->peer, st);
    if (op.tag == ZG_OP_ALLREDUCE) return zg_comm_allreduce(ctx, p->buffers[op.u.allreduce.buf] + op.u.allreduce.offset, op.u.allreduce.n, st);
    if (op.tag == ZG_OP_ALLGATHER)
        return zg_comm_allgather(ctx, p->buffers[op.u.allgather.src] + op.u.allgather.src_offset,
                                 p->buffers[op.u.allgather.dst] + op.u.allgather.dst_offset, op.u.allgather.n, st);
    if (op.tag == ZG_OP_MATMUL && !p->dense.empty()) {
        auto d = p->dense.find((uint32_t)i);
        if (d != p->dense.end()) {
            const auto& m = op.u.matmul; const ZgMatMulGeometry& g = m.geom;
            return zg_dense_head_launch(ctx, d->second, p->buffers[m.a] + g.a_offset, p->buffers[m.b], g.b_offset, g.b_col_stride,
                                        p->buffers[m.dst] + g.dst_offset, st);
        }
    }
    if (zg_op_is_batched(op.tag)) return zg_launch_batch(op, p->d_batch + p->single_entry[i], 1, p->d_dyn, st);
    return zg_launch_op(ctx, op, p->buffers.data(), p->d_dyn, (uint32_t)i, p->d_steps + p->step_off[i], st);
}

static bool launch_unit(ZgCudaProgram* p, const ZgCudaProgram::Unit& u, cudaStream_t st) {
    if (u.decode) { p->dec.plan.dyn = p->d_dyn; return zg_decode_launch(p->ctx, p->dec, st); }
    if (u.ar_norm) {
        const auto& ar = p->ops[u.ops[0]].u.allreduce;
        return zg_launch_peer_allreduce_norm(p->buffers[ar.buf], ar.n, p->ctx->peer, u.nm, st);
    }
    if (u.gemv_pair) {
        const uint32_t ia = u.entry_ops[0], ib = u.entry_ops[1];
        const auto& qa = p->ops[ia].u.qmatmul; const auto& qb = p->ops[ib].u.qmatmul;
        ZgGemvWs va = p->ws, vb = p->ws;
        if (va.partials) { va.partials += p->ws_part_off[ia]; va.partials_elems -= p->ws_part_off[ia]; vb.partials += p->ws_part_off[ib]; vb.partials_elems -= p->ws_part_off[ib]; }
        if (va.counters) { va.counters += p->ws_cnt_off[ia]; va.counters_n -= p->ws_cnt_off[ia]; vb.counters += p->ws_cnt_off[ib]; vb.counters_n -= p->ws_cnt_off[ib]; }
        return zg_qgemv_launch_pair(p->ctx, p->qweights[qa.weight_idx], p->qweights[qb.weight_idx], p->buffers[qa.input] + qa.input_offset,
                                    p->buffers[qa.dst], p->buffers[qb.dst], &va, &vb, u.epi, st);
    }
    if (u.attn_blk >= 0) {
        const ZgAttnBlock& B = p->attn_blocks[u.attn_blk];
        return zg_launch_attention_layer(p->d_attn_blocks + u.attn_blk, B.n_heads, B.d_head, u.ab_splits, p->d_dyn, p->d_attn_blk_part + u.ab_part_off,
                                         p->d_attn_blk_cnt + u.ab_cnt_off, st);
    }
    if (u.kvq == 1) return zg_kvq_launch_stores(p->d_kvq_store + u.first_entry, u.n_entries, u.kvq_max_warps, p->d_dyn, st);
    if (u.kvq == 2) return zg_kvq_launch_attention(p->d_kvq_attn + u.first_entry, u.n_entries, u.kvq_seq_q, p->d_dyn,
                                                   p->d_kvq_part ? p->d_kvq_part + u.kvq_part_off : nullptr, p->d_kvq_cnt ? p->d_kvq_cnt + u.kvq_cnt_off : nullptr, u.kvq_splits, st);
    if (u.chain) return zg_launch_chain(p->d_chain + u.first_entry, u.n_entries, p->d_dyn, st);
    if (u.ewmul) return zg_launch_ewmul(u.em, st);
    if (u.norm) return zg_launch_norm_macro(u.nm, st);
    if (u.gemv_batch && (u.entry_ops.size() > 1 || u.pros[0].kind != 0)) {
        const ZgCudaQWeight* w[kZgGemvBatch]; const float* xin[kZgGemvBatch]; float* xout[kZgGemvBatch];
        uint32_t irs[kZgGemvBatch], ors[kZgGemvBatch]; ZgGemvWs view[kZgGemvBatch];
        const uint32_t cnt = (uint32_t)u.entry_ops.size();
        for (uint32_t k = 0; k < cnt; k++) {
            const uint32_t i = u.entry_ops[k];
            const auto& q = p->ops[i].u.qmatmul;
            w[k] = p->qweights[q.weight_idx];
            xin[k] = p->buffers[q.input] + q.input_offset; xout[k] = p->buffers[q.dst] + q.dst_offset;
            irs[k] = q.input_row_stride; ors[k] = q.dst_row_stride;
            view[k] = p->ws;
            if (view[k].partials) { view[k].partials += p->ws_part_off[i]; view[k].partials_elems -= p->ws_part_off[i]; }
            if (view[k].counters) { view[k].counters += p->ws_cnt_off[i]; view[k].counters_n -= p->ws_cnt_off[i]; }
        }
        return zg_qgemv_launch_batch(p->ctx, cnt, w, xin, xout, p->ops[u.entry_ops[0]].u.qmatmul.M, irs, ors, view, st, u.pros.data());
    }
    if (!u.batched) return launch_one(p, u.ops[0], st);
    if (u.attn_splits > 1)
        return zg_launch_batch(p->ops[u.entry_ops[0]], p->d_batch + u.first_entry, u.n_entries, p->d_dyn, st,
                               p->d_attn_part + u.attn_part_off, p->d_attn_cnt + u.attn_cnt_off, u.attn_splits);
    return zg_launch_batch(p->ops[u.entry_ops[0]], p->d_batch + u.first_entry, u.n_entries, p->d_dyn, st);
}

static bool launch_all(ZgCudaProgram* p, cudaStream_t st, bool profile) {
    size_t n = p->ops.size();
    if (profile && p->prof_events.size() < n + 1) {
        while (p->prof_events.size() < n + 1) { cudaEvent_t e; cudaEventCreate(&e); p->prof_events.push_back(e); }
    }
    if (!profile) {   // schedule order: dependency levels, per-head ops batched
        for (const auto& u : p->units)
            if (!launch_unit(p, u, st)) return false;
        return true;
    }
    cudaEventRecord(p->prof_events[0], st);
    for (size_t i = 0; i < n; i++) {   // program order, one launch per op: per-tag device times
        if (!launch_one(p, i, st)) return false;
        cudaEventRecord(p->prof_events[i + 1], st);
    }
    return true;
}

// Inside a stream capture: ops whose buffer ranges do not conflict go to different capture streams, so the
// instantiated graph runs them as concurrent branches (q/k/v, gate/up, independent matvecs).  Results are
// identical to program order: every read-after-write, write-after-read and write-after-write pair stays ordered.
static bool launch_all_branched(ZgCudaProgram* p, cudaStream_t origin) {
    ZgCudaCtx* ctx = p->ctx;
    const size_t n = p->units.size();
    const int ns = 1 + (int)ctx->branch.size();
    std::vector<cudaStream_t> strs(1, origin);
    strs.insert(strs.end(), ctx->branch.begin(), ctx->branch.end());
    while (p->dep_events.size() < n + (size_t)ns + 1) {
        cudaEvent_t e;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { zg_set_error("cudaEventCreate failed"); return false; }
        p->dep_events.push_back(e);
    }
    std::vector<std::vector<ZgRange>> rng(n);
    std::vector<ZgRange> tmp;
    std::vector<int> unit_stream(n, 0);
    std::vector<long> last_on(ns, -1), dep(ns);
    std::vector<char> joined(ns, 0);
    joined[0] = 1;
    cudaEvent_t fork_ev = p->dep_events[n];
    ZG_CUDA_OK(cudaEventRecord(fork_ev, origin));
    int rr = 0;
    bool ok = true;
    for (size_t i = 0; i < n && ok; i++) {
        if (!p->units[i].ranges.empty()) rng[i] = p->units[i].ranges;
        else
            for (uint32_t oi : p->units[i].ops) {   // a unit reads / writes the union of its ops' ranges
                op_ranges(p, p->ops[oi], tmp);
                rng[i].insert(rng[i].end(), tmp.begin(), tmp.end());
            }
        std::fill(dep.begin(), dep.end(), -1L);
        int found = 0;
        for (long k = (long)i - 1; k >= 0 && found < ns; k--) {   // latest conflicting unit of every stream
            const int sk = unit_stream[k];
            if (dep[sk] >= 0) continue;
            if (ranges_conflict(rng[i], rng[k])) { dep[sk] = k; found++; }
        }
        int s = -1;
        for (int sk = 0; sk < ns; sk++)
            if (dep[sk] >= 0 && last_on[sk] == dep[sk]) { s = sk; break; }   // continue the producer's stream
        if (s < 0) { s = rr; rr = (rr + 1) % ns; }
        if (!joined[s]) { ZG_CUDA_OK(cudaStreamWaitEvent(strs[s], fork_ev, 0)); joined[s] = 1; }
        for (int sk = 0; sk < ns; sk++)
            if (sk != s && dep[sk] >= 0) ZG_CUDA_OK(cudaStreamWaitEvent(strs[s], p->dep_events[dep[sk]], 0));
        ok = launch_unit(p, p->units[i], strs[s]);
        ZG_CUDA_OK(cudaEventRecord(p->dep_events[i], strs[s]));
        unit_stream[i] = s; last_on[s] = (long)i;
    }
    for (int sk = 1; sk < ns; sk++) {   // join every branch back into the origin stream (also on failure: the capture must end cleanly)
        if (!joined[sk]) continue;
        cudaEventRecord(p->dep_events[n + 1 + sk - 1], strs[sk]);
        cudaStreamWaitEvent(origin, p->dep_events[n + 1 + sk - 1], 0);
    }
    return ok;
}

static bool run_ops(ZgCudaProgram* p) {
    ZgCudaCtx* ctx = p->ctx;
    cudaStream_t st = ctx->stream;
    size_t n = p->ops.size();
    if (n == 0) return true;
    if (p->dyn_dirty) {
        ZG_CUDA_OK(cudaMemcpyAsync(p->d_dyn, p->h_dyn, n * 4, cudaMemcpyHostToDevice, st));
        p->dyn_dirty = false;
    }
    if (ctx->profiling) {
        if (!launch_all(p, st, true)) return false;
        ZG_CUDA_OK(cudaStreamSynchronize(st));
        for (size_t i = 0; i < n; i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, p->prof_events[i], p->prof_events[i + 1]);
            if (p->ops[i].tag < ZG_OP_COUNT) p->profile.time_ns[p->ops[i].tag] += (uint64_t)(ms * 1e6);
        }
        p->profile.backend_op_count += n;
        p->profile.backend_dispatch_count += n;
        p->profile.call_count += 1;
        return true;
    }
    if (!ctx->graph_mode) return launch_all(p, st, false);
    if (!p->graph_valid) {
        if (p->exec) { cudaGraphExecDestroy(p->exec); p->exec = nullptr; }
        if (p->graph) { cudaGraphDestroy(p->graph); p->graph = nullptr; }
        ZG_CUDA_OK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const uint64_t before = g_zg_launches.load(), before_s = g_zg_stream_launches.load();
        bool ok = ctx->branch.empty() ? launch_all(p, st, false) : launch_all_branched(p, st);
        p->graph_kernels = g_zg_launches.load() - before;
        p->graph_streamed = g_zg_stream_launches.load() - before_s;
        g_zg_launches.store(before); // captured, not launched: counted per replay below
        cudaError_t e = cudaStreamEndCapture(st, &p->graph);
        if (!ok) { if (p->graph) { cudaGraphDestroy(p->graph); p->graph = nullptr; } return false; }
        if (e != cudaSuccess) { zg_set_error("graph capture failed: %s", cudaGetErrorString(e)); return false; }
        ZG_CUDA_OK(cudaGraphInstantiate(&p->exec, p->graph, 0));
        p->graph_valid = true;
    }
    ZG_CUDA_OK(cudaGraphLaunch(p->exec, st));
    g_zg_launches.fetch_add(p->graph_kernels, std::memory_order_relaxed);
    return true;
}

__global__ void k_scatter_inputs(const ZgCudaProgram::InSeg* __restrict__ tab, const uint8_t* __restrict__ stage) {
    const ZgCudaProgram::InSeg s = tab[blockIdx.x];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(stage + s.src_off);
    uint32_t* dst = reinterpret_cast<uint32_t*>(s.dst);
    for (uint32_t i = threadIdx.x; i < s.words; i += blockDim.x) dst[i] = src[i];
}

// Many small inputs: pack -> one pinned H2D copy -> scatter kernel.  Returns false when the inputs do not qualify (the
// caller then copies them one by one).
static bool upload_inputs_staged(ZgCudaProgram* p, const ZgIO* in, size_t n_in, cudaStream_t st) {
    if (n_in < 3) return false;
    size_t total = 0;
    for (size_t i = 0; i < n_in; i++) {
        if ((in[i].offset | in[i].size) & 3u) return false;
        total += (in[i].size + 15u) & ~15u;
    }
    if (total == 0 || total > (4u << 20) || n_in > 512) return false;
    for (size_t i = 0; i < n_in; i++)   // overlapping destinations must keep "later input wins": leave them to ordered copies
        for (size_t j = 0; j < i; j++)
            if (in[i].buf_idx == in[j].buf_idx && in[i].offset < in[j].offset + in[j].size && in[j].offset < in[i].offset + in[i].size) return false;
    if (total > p->in_stage_bytes) {
        if (p->h_in_stage) cudaFreeHost(p->h_in_stage);
        cudaFree(p->d_in_stage);
        p->h_in_stage = nullptr; p->d_in_stage = nullptr; p->in_stage_bytes = 0;
        if (cudaMallocHost(&p->h_in_stage, total) != cudaSuccess || cudaMalloc(&p->d_in_stage, total) != cudaSuccess) { cudaGetLastError(); return false; }
        p->in_stage_bytes = total;
    }
    std::vector<ZgCudaProgram::InSeg> tab(n_in);
    size_t off = 0;
    for (size_t i = 0; i < n_in; i++) {
        memcpy(p->h_in_stage + off, in[i].host_ptr, in[i].size);
        tab[i] = {reinterpret_cast<float*>((uint8_t*)p->buffers[in[i].buf_idx] + in[i].offset), (uint32_t)off, in[i].size / 4};
        off += (in[i].size + 15u) & ~15u;
    }
    bool same = tab.size() == p->in_tab_host.size();
    for (size_t i = 0; same && i < tab.size(); i++)
        same = tab[i].dst == p->in_tab_host[i].dst && tab[i].src_off == p->in_tab_host[i].src_off && tab[i].words == p->in_tab_host[i].words;
    if (!same) {
        if (n_in > p->in_tab_cap) {
            cudaFree(p->d_in_tab); p->d_in_tab = nullptr; p->in_tab_cap = 0;
            if (cudaMalloc(&p->d_in_tab, n_in * sizeof(ZgCudaProgram::InSeg)) != cudaSuccess) { cudaGetLastError(); return false; }
            p->in_tab_cap = n_in;
        }
        if (cudaMemcpyAsync(p->d_in_tab, tab.data(), n_in * sizeof(ZgCudaProgram::InSeg), cudaMemcpyHostToDevice, st) != cudaSuccess) return false;
        cudaStreamSynchronize(st);   // `tab` is a local
        p->in_tab_host = tab;
    }
    if (cudaMemcpyAsync(p->d_in_stage, p->h_in_stage, total, cudaMemcpyHostToDevice, st) != cudaSuccess) return false;
    k_scatter_inputs<<<(unsigned)n_in, 128, 0, st>>>(p->d_in_tab, p->d_in_stage);
    ZG_COUNT_LAUNCH();
    return true;
}

extern "C" void zg_cuda_execute(ZgCudaCtx* ctx, ZgCudaProgram* p, const ZgIO* in, size_t n_in, const ZgIO* out, size_t n_out) {
    if (!ctx || !p) return;
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    for (size_t i = 0; i < n_in; i++) {
        const ZgIO& io = in[i];
        if (io.buf_idx >= p->buffers.size() || (size_t)io.offset + io.size > p->buffer_elems[io.buf_idx] * sizeof(float)) {
            zg_set_error("execute: input %zu out of range", i); return; // reference asserts (reference.zig:116)
        }
    }
    if (!upload_inputs_staged(p, in, n_in, st))
        for (size_t i = 0; i < n_in; i++)
            cudaMemcpyAsync((uint8_t*)p->buffers[in[i].buf_idx] + in[i].offset, in[i].host_ptr, in[i].size, cudaMemcpyHostToDevice, st);
    for (auto& d : p->dense) {   // an input that overwrites a promoted operand refreshes its bf16 copy
        const auto& m = p->ops[d.first].u.matmul;
        for (size_t i = 0; i < n_in; i++)
            if (in[i].buf_idx == m.b) { zg_dense_head_refresh(ctx, d.second, p->buffers[m.b], m.geom.b_offset, m.geom.b_col_stride, st); break; }
    }
    if (!run_ops(p)) { cudaStreamSynchronize(st); return; }
    for (size_t i = 0; i < n_out; i++) {
        const ZgIO& io = out[i];
        if (io.buf_idx >= p->buffers.size() || (size_t)io.offset + io.size > p->buffer_elems[io.buf_idx] * sizeof(float)) {
            zg_set_error("execute: output %zu out of range", i); continue;
        }
        cudaMemcpyAsync(io.host_ptr, (uint8_t*)p->buffers[io.buf_idx] + io.offset, io.size, cudaMemcpyDeviceToHost, st);
    }
    if (p->dec.valid) cudaMemcpyAsync(p->dec.h_err, p->dec.d_sync + 64, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    zg_peer_check_enqueue(ctx, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) zg_set_error("execute: %s", cudaGetErrorString(e));
    zg_peer_check_result(ctx);
    if (p->dec.valid && *p->dec.h_err) {   // a bounded wait inside the fused decode kernel gave up: results are invalid
        static const char* what[] = {"", "grid barrier", "weight ring (TMA)", "NVLink peer all-reduce: a peer rank never arrived"};
        zg_set_error("execute: fused decode kernel timed out in its %s wait; the step's outputs are invalid", what[*p->dec.h_err & 3]);
        cudaMemsetAsync(p->dec.d_sync, 0, 128 * sizeof(uint32_t), st);
        cudaStreamSynchronize(st);
        *p->dec.h_err = 0;
    }
}

extern "C" void zg_cuda_execute_device(ZgCudaCtx* ctx, ZgCudaProgram* p) {
    if (!ctx || !p) return;
    cudaSetDevice(ctx->device);
    run_ops(p);
}

extern "C" void zg_cuda_free(ZgCudaCtx* ctx, ZgCudaProgram* p) { (void)ctx; free_program(p); }

// Debug timeline: enable (allocates the device buffer) / disable, and copy out {kind, t_entry, t_after_wait, t_exit} records.
static unsigned long long* g_trace_buf = nullptr;
extern "C" int zg_cuda_trace(ZgCudaCtx* ctx, int enable) {
    if (!ctx) return -1;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (enable && !g_trace_buf) { if (cudaMalloc(&g_trace_buf, (1 + 3 * 16000) * 8) != cudaSuccess) return -1; }
    if (g_trace_buf) cudaMemset(g_trace_buf, 0, (1 + 3 * 16000) * 8);
    zg_trace_set_ops(enable ? g_trace_buf : nullptr);
    zg_trace_set_gemv(enable ? g_trace_buf : nullptr);
    zg_trace_set_gemv_stream(enable ? g_trace_buf : nullptr);
    zg_trace_set_decode(enable ? g_trace_buf : nullptr);
    return 0;
}
extern "C" size_t zg_cuda_trace_read(ZgCudaCtx* ctx, unsigned long long* host, size_t max_records) {
    if (!ctx || !g_trace_buf || !host) return 0;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    unsigned long long n = 0;
    cudaMemcpy(&n, g_trace_buf, 8, cudaMemcpyDeviceToHost);
    if (n > 16000) n = 16000;
    if (n > max_records) n = max_records;
    cudaMemcpy(host, g_trace_buf + 1, n * 3 * 8, cudaMemcpyDeviceToHost);
    return (size_t)n;
}

extern "C" const ZgProfile* zg_cuda_profile(ZgCudaCtx* ctx, ZgCudaProgram* p) {
    if (!ctx || !p || !ctx->profiling) return nullptr;
    return &p->profile;
}

// LlamaInferenceSession.quantizeKV (src/llama_inference.zig:648-679, plan side :277-328) for a compiled program.
extern "C" int zg_cuda_program_quantize_kv(ZgCudaCtx* ctx, ZgCudaProgram* p, size_t block_size, int int8_query) {
    if (!ctx || !p || block_size == 0 || (block_size & 3)) { zg_set_error("program_quantize_kv: bad arguments (block_size must be a positive multiple of 4)"); return -1; }
    if (p->kvq) return 0;   // like the reference: a second call keeps the caches
    cudaSetDevice(ctx->device);
    const size_t n = p->ops.size();
    std::vector<char> role(n, 0);
    std::map<uint32_t, uint32_t> kv_dh;   // cache buffer -> d_head
    for (size_t i = 0; i < n; i++) {
        if (p->ops[i].tag != ZG_OP_ATTENTION) continue;
        const auto& a = p->ops[i].u.attention;
        const uint32_t dh = a.d_head;
        if (a.k_rs != 1 || a.v_rs != 1 || a.q_rs != 1 || a.dst_rs != 1 || a.k_cs != dh || a.v_cs != dh || dh == 0 || dh > 512 || (dh & 3) || dh % block_size ||
            a.k_off % dh || a.v_off % dh || (int8_query && dh / block_size > 32)) {
            zg_set_error("program_quantize_kv: op %zu: attention layout not expressible over a column-major Q8 cache (src/quant.zig:941-947)", i); return -1;
        }
        for (uint32_t b : {a.k, a.v}) {
            if (kv_dh.count(b) && kv_dh[b] != dh) { zg_set_error("program_quantize_kv: buffer %u is read with two head sizes", b); return -1; }
            kv_dh[b] = dh;
        }
        role[i] = 2;
    }
    if (kv_dh.empty()) { zg_set_error("program_quantize_kv: the program has no attention op"); return -1; }
    for (size_t i = 0; i < n; i++) {
        const ZgOp& op = p->ops[i];
        if (op.tag == ZG_OP_SLICE_ASSIGN && kv_dh.count(op.u.slice_assign.dst)) {
            const auto& sa = op.u.slice_assign;
            const uint32_t dh = kv_dh[sa.dst];
            if (sa.rows != dh || sa.dst_row_stride != 1 || sa.dst_col_stride != dh || sa.src_row_stride != 1 || sa.dst_base_offset % dh || sa.dst_offset % dh ||
                (sa.patch_stride && sa.patch_stride % dh) || kv_dh.count(sa.src)) {
                zg_set_error("program_quantize_kv: op %zu: cache write is not a run of whole d_head columns (src/llama_inference.zig:336-348)", i); return -1;
            }
            role[i] = 1;
            continue;
        }
        if (role[i]) {   // attention: q / mask / dst must not be cache buffers
            const auto& a = op.u.attention;
            if (kv_dh.count(a.q) || kv_dh.count(a.dst) || (a.has_mask && kv_dh.count(a.mask))) { zg_set_error("program_quantize_kv: op %zu: a cache buffer is used as q / mask / dst", i); return -1; }
            continue;
        }
        std::vector<ZgRange> rr;
        op_ranges(p, op, rr);
        for (const ZgRange& r : rr)
            if (kv_dh.count(r.buf)) { zg_set_error("program_quantize_kv: op %zu (tag %u) touches a KV-cache buffer directly", i, op.tag); return -1; }
    }
    cudaStreamSynchronize(ctx->stream);
    for (auto& kv : kv_dh) {
        ZgCudaKVCache* c = zg_cuda_kvcache_create(ctx, kv.second, p->buffer_elems[kv.first] / kv.second, block_size);
        if (!c) { for (auto& q : p->kv_caches) zg_cuda_kvcache_free(ctx, q.second); p->kv_caches.clear(); return -1; }
        p->kv_caches[kv.first] = c;
    }
    p->kvq = true; p->kvq_bs = block_size; p->kvq_int8 = int8_query ? 1 : 0; p->kvq_role = role;
    p->graph_valid = false;
    if (!build_schedule(p)) { p->kvq = false; build_schedule(p); return -1; }
    return 0;
}

// Dense B operands of single-row matmuls (the tied LM head) stored as bf16 next to the f32 original; see dense_head.cu.
// Returns the number of matmul ops promoted (0: none qualified), -1 on error.
extern "C" int zg_cuda_program_promote_dense(ZgCudaCtx* ctx, ZgCudaProgram* p, int format) {
    if (!ctx || !p || (format != ZG_DENSE_BF16 && format != ZG_DENSE_F16)) { zg_set_error("program_promote_dense: bad arguments (format must be ZG_DENSE_BF16 or ZG_DENSE_F16)"); return -1; }
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    int n_promoted = 0;
    std::vector<ZgRange> rr;
    for (size_t i = 0; i < p->ops.size(); i++) {
        const ZgOp& op = p->ops[i];
        if (op.tag != ZG_OP_MATMUL || p->dense.count((uint32_t)i)) continue;
        const auto& m = op.u.matmul; const ZgMatMulGeometry& g = m.geom;
        // one activation row (staged in <= 32 KB of shared memory) against a k-contiguous operand with 16-byte aligned rows; small heads stay exact
        if (g.M != 1 || g.b_row_stride != 1 || g.a_col_stride != 1 || (g.K & 7) || g.K < 64 || g.K > 8192 || g.N < 1024 || (g.b_col_stride & 3) || (g.b_offset & 3) ||
            (g.a_offset & 3) || (g.dst_offset & 3) || g.b_col_stride < g.K) continue;
        bool written = false;   // the operand must be a constant of the program: no op writes it
        for (size_t k = 0; k < p->ops.size() && !written; k++) {
            op_ranges(p, p->ops[k], rr);
            for (const ZgRange& r : rr) if (r.write && r.buf == m.b) { written = true; break; }
        }
        if (written) continue;
        ZgDenseHead* h = zg_dense_head_create(ctx, format, p->buffers[m.b], g.b_offset, g.b_col_stride, (uint32_t)g.N, (uint32_t)g.K);
        if (!h) return -1;
        p->dense[(uint32_t)i] = h;
        n_promoted++;
    }
    if (n_promoted) { cudaStreamSynchronize(ctx->stream); p->graph_valid = false; }
    return n_promoted;
}

extern "C" uint64_t zg_cuda_program_stats(const ZgCudaProgram* p, int what) {
    if (!p) return 0;
    switch (what) {
        case 0: return p->graph_kernels;
        case 1: return p->dec.valid ? p->dec_count : 0;
        case 2: return p->dec.valid ? p->dec.plan.n_layers : 0;
        case 3: return p->graph_streamed;
        case 4: { uint64_t b = 0; for (auto& d : p->dense) b += (uint64_t)d.second->N * d.second->K * 2 - (uint64_t)d.second->N * 6; return b; }   // HBM bytes per execution saved by the bf16 operands
        default: return 0;
    }
}

extern "C" void* zg_cuda_program_buffer(ZgCudaProgram* p, uint32_t idx) {
    if (!p || idx >= p->buffers.size()) return nullptr;
    return p->buffers[idx];
}

// ── direct quantized matmul calls ────────────────────────────────────────────
extern "C" int zg_cuda_qmatmul_device(ZgCudaCtx* ctx, const ZgCudaQWeight* w, const float* d_input, float* d_dst,
                                      uint32_t M, uint32_t input_row_stride, uint32_t dst_row_stride) {
    if (!ctx || !w) { zg_set_error("qmatmul_device: null argument"); return -1; }
    cudaSetDevice(ctx->device);
    size_t pe = 0, nc = 0;
    zg_qgemv_ws_need(ctx, w, M, &pe, &nc);
    if (!zg_gemv_ws_reserve(&ctx->ws, pe, nc, ctx->stream, zg_qgemm_scratch_elems(w, M))) return -1;
    return zg_qmatmul_launch(ctx, w, d_input, d_dst, M, input_row_stride, dst_row_stride, &ctx->ws, ctx->stream) ? 0 : -1;
}

extern "C" int zg_cuda_qmatmul_host(ZgCudaCtx* ctx, const ZgCudaQWeight* w, const float* h_input, float* h_dst, uint32_t M) {
    if (!ctx || !w) { zg_set_error("qmatmul_host: null argument"); return -1; }
    cudaSetDevice(ctx->device);
    float *d_in = nullptr, *d_out = nullptr;
    size_t in_b = (size_t)M * w->K * 4, out_b = (size_t)M * w->N * 4;
    if (cudaMalloc(&d_in, in_b ? in_b : 4) != cudaSuccess || cudaMalloc(&d_out, out_b ? out_b : 4) != cudaSuccess) {
        zg_set_error("qmatmul_host: cudaMalloc failed"); cudaFree(d_in); cudaFree(d_out); return -1;
    }
    cudaMemcpyAsync(d_in, h_input, in_b, cudaMemcpyHostToDevice, ctx->stream);
    int rc = zg_cuda_qmatmul_device(ctx, w, d_in, d_out, M, 0, 0);
    cudaMemcpyAsync(h_dst, d_out, out_b, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_in); cudaFree(d_out);
    if (e != cudaSuccess) { zg_set_error("qmatmul_host: %s", cudaGetErrorString(e)); return -1; }
    return rc;
}

// QuantizedWeight.matmulBias (src/quant.zig:581-589): matmul, then dst[m, n] += bias[n]
__global__ void k_add_bias(float* __restrict__ dst, const float* __restrict__ bias, uint32_t M, uint32_t N) {
    const size_t total = (size_t)M * N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) dst[i] += bias[i % N];
}
extern "C" int zg_cuda_qmatmul_bias_host(ZgCudaCtx* ctx, const ZgCudaQWeight* w, const float* h_input, const float* h_bias, float* h_dst, uint32_t M) {
    if (!ctx || !w || !h_bias) { zg_set_error("qmatmul_bias_host: null argument"); return -1; }
    cudaSetDevice(ctx->device);
    float *d_in = nullptr, *d_out = nullptr, *d_b = nullptr;
    const size_t in_b = (size_t)M * w->K * 4, out_b = (size_t)M * w->N * 4, b_b = w->N * 4;
    if (cudaMalloc(&d_in, in_b ? in_b : 4) != cudaSuccess || cudaMalloc(&d_out, out_b ? out_b : 4) != cudaSuccess || cudaMalloc(&d_b, b_b ? b_b : 4) != cudaSuccess) {
        zg_set_error("qmatmul_bias_host: cudaMalloc failed"); cudaFree(d_in); cudaFree(d_out); cudaFree(d_b); return -1;
    }
    cudaMemcpyAsync(d_in, h_input, in_b, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_b, h_bias, b_b, cudaMemcpyHostToDevice, ctx->stream);
    int rc = zg_cuda_qmatmul_device(ctx, w, d_in, d_out, M, 0, 0);
    if (rc == 0 && M && w->N) {
        const size_t total = (size_t)M * w->N;
        k_add_bias<<<(unsigned)std::min<size_t>((total + 255) / 256, 4096), 256, 0, ctx->stream>>>(d_out, d_b, M, (uint32_t)w->N);
        ZG_COUNT_LAUNCH();
    }
    cudaMemcpyAsync(h_dst, d_out, out_b, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_b);
    if (e != cudaSuccess) { zg_set_error("qmatmul_bias_host: %s", cudaGetErrorString(e)); return -1; }
    return rc;
}

// ── raw device memory helpers for tests / bench ──────────────────────────────
extern "C" void* zg_cuda_malloc(ZgCudaCtx* ctx, size_t bytes) {
    if (!ctx) return nullptr;
    cudaSetDevice(ctx->device);
    void* p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { zg_set_error("cudaMalloc(%zu) failed", bytes); return nullptr; }
    return p;
}
extern "C" void zg_cuda_free_device(ZgCudaCtx* ctx, void* p) { if (ctx) cudaSetDevice(ctx->device); cudaFree(p); }
extern "C" int zg_cuda_memcpy_h2d(ZgCudaCtx* ctx, void* d, const void* h, size_t bytes) {
    cudaError_t e = cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { zg_set_error("memcpy_h2d: %s", cudaGetErrorString(e)); return -1; }
    return 0;
}
extern "C" int zg_cuda_memcpy_d2h(ZgCudaCtx* ctx, void* h, const void* d, size_t bytes) {
    cudaError_t e = cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { zg_set_error("memcpy_d2h: %s", cudaGetErrorString(e)); return -1; }
    return 0;
}
extern "C" int zg_cuda_memset(ZgCudaCtx* ctx, void* d, int value, size_t bytes) {
    cudaError_t e = cudaMemsetAsync(d, value, bytes, ctx->stream);
    if (e != cudaSuccess) { zg_set_error("memset: %s", cudaGetErrorString(e)); return -1; }
    return 0;
}
