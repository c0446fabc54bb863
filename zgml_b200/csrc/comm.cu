// NCCL collectives for row-sharded linear layers: one process per GPU, one communicator per
// context, collectives enqueued on the program's stream (captured into its CUDA graph like any kernel).
// libnccl is resolved with dlopen at first use, so single-GPU users need no NCCL at all.
//
// Small all-reduces (the d_model partial sums after the o / down projections: 32 KB at Llama-3-70B, 160 per token)
// are latency-bound, so they do not go through NCCL: every rank exports one cudaMalloc region of slots with
// cudaIpc, maps its peers' regions, and the all-reduce becomes peer stores + a flag in a 16-CTA one-shot
// all-reduce kernel (ops.cu k_allreduce_peer, data + epoch in every 16-byte store) — one NVLink store latency instead of a ring.  NCCL carries the handle
// exchange, the all-gathers, and any all-reduce larger than a slot.  ZG_CUDA_PEER=0 keeps everything on NCCL.
#include "zg_internal.cuh"

#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {

struct NcclId { char internal[128]; };
typedef int (*GetUniqueIdFn)(NcclId*);
typedef int (*CommInitRankFn)(void**, int, NcclId, int);
typedef int (*CommDestroyFn)(void*);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*AllGatherFn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*GetErrorStringFn)(int);

struct Nccl {
    void* lib = nullptr;
    GetUniqueIdFn get_id = nullptr;
    CommInitRankFn init_rank = nullptr;
    CommDestroyFn destroy = nullptr;
    AllReduceFn all_reduce = nullptr;
    AllGatherFn all_gather = nullptr;
    GetErrorStringFn err = nullptr;
} g_nccl;

constexpr int kNcclFloat32 = 7, kNcclSum = 0, kNcclChar = 0;
constexpr uint32_t kPeerSlotFloats = 65536;   // 256 KB per (set, rank): batch-8 partial sums of d_model 8192

bool load_nccl() {
    if (g_nccl.lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) { zg_set_error("NCCL not found (dlopen libnccl.so.2): %s", dlerror()); return false; }
    g_nccl.get_id = (GetUniqueIdFn)dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.init_rank = (CommInitRankFn)dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.destroy = (CommDestroyFn)dlsym(g_nccl.lib, "ncclCommDestroy");
    g_nccl.all_reduce = (AllReduceFn)dlsym(g_nccl.lib, "ncclAllReduce");
    g_nccl.all_gather = (AllGatherFn)dlsym(g_nccl.lib, "ncclAllGather");
    g_nccl.err = (GetErrorStringFn)dlsym(g_nccl.lib, "ncclGetErrorString");
    if (!g_nccl.get_id || !g_nccl.init_rank || !g_nccl.destroy || !g_nccl.all_reduce || !g_nccl.all_gather) {
        zg_set_error("libnccl lacks an expected symbol");
        g_nccl.lib = nullptr;
        return false;
    }
    return true;
}

bool check(int rc, const char* what) {
    if (rc == 0) return true;
    zg_set_error("%s failed: %s", what, g_nccl.err ? g_nccl.err(rc) : "nccl error");
    return false;
}

} // namespace

static bool peer_setup(ZgCudaCtx* ctx);

extern "C" int zg_cuda_comm_unique_id(void* id128) {
    if (!id128 || !load_nccl()) return -1;
    NcclId id;
    if (!check(g_nccl.get_id(&id), "ncclGetUniqueId")) return -1;
    memcpy(id128, &id, sizeof(id));
    return 0;
}

extern "C" int zg_cuda_comm_init(ZgCudaCtx* ctx, const void* id128, int rank, int world) {
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) { zg_set_error("comm_init: bad arguments"); return -1; }
    if (!load_nccl()) return -1;
    cudaSetDevice(ctx->device);
    NcclId id;
    memcpy(&id, id128, sizeof(id));
    void* comm = nullptr;
    if (!check(g_nccl.init_rank(&comm, world, id, rank), "ncclCommInitRank")) return -1;
    ctx->nccl_comm = comm; ctx->rank = rank; ctx->world = world;
    if (world > 1) peer_setup(ctx);   // optional: falls back to NCCL all-reduces when IPC mapping is unavailable
    return 0;
}

// Export this rank's slot region, gather everybody's IPC handle through NCCL, map the peers.  COLLECTIVE: every rank runs
// the same two NCCL calls whatever happens locally (ZG_CUDA_PEER=0 on some ranks only, a failed allocation, a failed
// export or mapping): local failures only clear `ok`, which travels with the handle, and the decision is taken from the
// min-reduced value — so no rank can be left waiting in a collective its peers skipped.
static bool peer_setup(ZgCudaCtx* ctx) {
    const int world = ctx->world, rank = ctx->rank;
    if (world > kZgMaxRanks) return false;   // same on every rank
    bool ok = true;
    if (const char* e = getenv("ZG_CUDA_PEER")) if (e[0] == '0') ok = false;
    const size_t slot_bytes = (size_t)kZgPeerSets * world * kPeerSlotFloats * 2 * sizeof(float);   // every float travels with its epoch
    const size_t seq_bytes = ((2 + kZgPeerCtas) * sizeof(uint32_t) + 15) & ~(size_t)15;
    const size_t cell_bytes = (size_t)kZgPeerSets * kZgPeerCtas * sizeof(unsigned long long);
    const size_t total = slot_bytes + seq_bytes + cell_bytes + 64;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (ok && cudaMalloc(&ctx->peer_mem, total) != cudaSuccess) { cudaGetLastError(); ctx->peer_mem = nullptr; ok = false; }
    if (ok) cudaMemset(ctx->peer_mem, 0, total);
    if (ok && cudaIpcGetMemHandle(&mine, ctx->peer_mem) != cudaSuccess) { cudaGetLastError(); ok = false; }
    // record = 64-byte handle + 64 bytes whose first byte is this rank's ok flag
    constexpr size_t kRec = 128;
    std::vector<char> all(kRec * (size_t)world, 0);
    char rec[kRec];
    memset(rec, 0, sizeof(rec));
    memcpy(rec, &mine, 64);
    rec[64] = ok ? 1 : 0;
    char* d_recs = nullptr;
    bool xfer = cudaMalloc(&d_recs, kRec * (size_t)world) == cudaSuccess;
    if (!xfer) { cudaGetLastError(); d_recs = nullptr; }
    // the collectives run even when the local staging failed (with a scratch buffer the size of one record)
    char* d_fallback = nullptr;
    if (!xfer && cudaMalloc(&d_fallback, kRec * (size_t)world) != cudaSuccess) { cudaGetLastError(); d_fallback = nullptr; }
    char* d_buf = xfer ? d_recs : d_fallback;
    bool gathered = false;
    if (d_buf) {
        if (!xfer) rec[64] = 0;
        cudaMemcpy(d_buf + kRec * rank, rec, kRec, cudaMemcpyHostToDevice);
        gathered = check(g_nccl.all_gather(d_buf + kRec * rank, d_buf, kRec, kNcclChar, ctx->nccl_comm, ctx->stream), "ncclAllGather(ipc handles)") &&
                   cudaStreamSynchronize(ctx->stream) == cudaSuccess &&
                   cudaMemcpy(all.data(), d_buf, kRec * (size_t)world, cudaMemcpyDeviceToHost) == cudaSuccess;
    }
    int mapped_ok = (ok && gathered) ? 1 : 0;
    for (int r = 0; r < world && mapped_ok; r++) if (!all[kRec * r + 64]) mapped_ok = 0;   // some rank opted out or failed: nobody maps
    for (int r = 0; r < world && mapped_ok; r++) {
        if (r == rank) { ctx->peer_mapped[r] = nullptr; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, all.data() + kRec * r, 64);
        if (cudaIpcOpenMemHandle(&ctx->peer_mapped[r], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ctx->peer_mapped[r] = nullptr; mapped_ok = 0; }
    }
    // every rank must take the same path: agree through a min-reduction (reuses the staging buffer; a rank without one
    // cannot take part in NCCL at all and has already failed above together with everybody else's all-gather)
    float h = (float)mapped_ok;
    if (d_buf) {
        float* d_f = reinterpret_cast<float*>(d_buf);
        cudaMemcpy(d_f, &h, sizeof(float), cudaMemcpyHostToDevice);
        if (!check(g_nccl.all_reduce(d_f, d_f, 1, kNcclFloat32, /*ncclMin*/ 3, ctx->nccl_comm, ctx->stream), "ncclAllReduce(peer agreement)")) h = 0.f;
        else {
            cudaStreamSynchronize(ctx->stream);
            cudaMemcpy(&h, d_f, sizeof(float), cudaMemcpyDeviceToHost);
        }
    } else h = 0.f;
    cudaFree(d_recs); cudaFree(d_fallback);
    if (h < 0.5f) {
        for (int r = 0; r < world; r++) if (ctx->peer_mapped[r]) { cudaIpcCloseMemHandle(ctx->peer_mapped[r]); ctx->peer_mapped[r] = nullptr; }
        cudaFree(ctx->peer_mem); ctx->peer_mem = nullptr;
        return false;
    }
    ZgPeerComm& pc = ctx->peer;
    pc.rank = rank; pc.world = world; pc.max_n = kPeerSlotFloats;
    for (int r = 0; r < world; r++) {
        char* base = (char*)(r == rank ? ctx->peer_mem : ctx->peer_mapped[r]);
        pc.slots[r] = (float*)base;
    }
    pc.seq = (uint32_t*)((char*)ctx->peer_mem + slot_bytes);
    pc.cells = (unsigned long long*)((char*)ctx->peer_mem + slot_bytes + seq_bytes);
    return true;
}

// The peer all-reduce kernels raise pc.seq[1] when a peer never arrived (their result is NaN-poisoned).  The word is
// copied to pinned host memory behind the step's work (enqueue) and turned into an error string after the
// synchronisation (result).  Sticky until reported.
void zg_peer_check_enqueue(ZgCudaCtx* ctx, cudaStream_t st) {
    if (!ctx || ctx->world <= 1 || !ctx->peer.max_n || !ctx->peer.seq) return;
    if (!ctx->h_peer_err && cudaMallocHost(&ctx->h_peer_err, sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); ctx->h_peer_err = nullptr; return; }
    cudaMemcpyAsync(ctx->h_peer_err, ctx->peer.seq + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
}
bool zg_peer_check_result(ZgCudaCtx* ctx) {
    if (!ctx || !ctx->h_peer_err || !*ctx->h_peer_err) return true;
    zg_set_error("NVLink peer all-reduce %u timed out on rank %d: a peer rank never delivered its data; this rank's activations are NaN-poisoned",
                 *ctx->h_peer_err, ctx->rank);
    *ctx->h_peer_err = 0;
    if (ctx->peer.seq) cudaMemset(ctx->peer.seq + 1, 0, sizeof(uint32_t));
    return false;
}

bool zg_peer_allreduce_ok(const ZgCudaCtx* ctx, size_t n) {
    return ctx->world > 1 && ctx->peer.max_n != 0 && n != 0 && (n & 3) == 0 && n <= ctx->peer.max_n;
}

extern "C" int zg_cuda_comm_mode(const ZgCudaCtx* ctx) {   // 0 none, 1 NCCL only, 2 NVLink peer-memory all-reduce + NCCL
    if (!ctx || !ctx->nccl_comm) return 0;
    return ctx->peer.max_n ? 2 : 1;
}

extern "C" void zg_cuda_comm_destroy(ZgCudaCtx* ctx) {
    if (!ctx || !ctx->nccl_comm) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->peer_mem) {
        float one = 0.f; float* d_f = nullptr;   // barrier: nobody unmaps while a peer may still store into it
        if (cudaMalloc(&d_f, sizeof(float)) == cudaSuccess) {
            cudaMemcpy(d_f, &one, sizeof(float), cudaMemcpyHostToDevice);
            g_nccl.all_reduce(d_f, d_f, 1, kNcclFloat32, kNcclSum, ctx->nccl_comm, ctx->stream);
            cudaStreamSynchronize(ctx->stream);
            cudaFree(d_f);
        }
        for (int r = 0; r < kZgMaxRanks; r++) if (ctx->peer_mapped[r]) { cudaIpcCloseMemHandle(ctx->peer_mapped[r]); ctx->peer_mapped[r] = nullptr; }
        cudaFree(ctx->peer_mem); ctx->peer_mem = nullptr;
        ctx->peer = ZgPeerComm();
    }
    if (ctx->h_peer_err) { cudaFreeHost(ctx->h_peer_err); ctx->h_peer_err = nullptr; }
    g_nccl.destroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr; ctx->world = 1; ctx->rank = 0;
}

bool zg_comm_allreduce(ZgCudaCtx* ctx, float* buf, size_t n, cudaStream_t st) {
    if (ctx->world == 1) return true;   // a single rank's sum is itself
    if (!ctx->nccl_comm) { zg_set_error("allreduce op without zg_cuda_comm_init"); return false; }
    ZG_COUNT_LAUNCH();
    return check(g_nccl.all_reduce(buf, buf, n, kNcclFloat32, kNcclSum, ctx->nccl_comm, st), "ncclAllReduce");
}

bool zg_comm_allgather(ZgCudaCtx* ctx, const float* src, float* dst, size_t n_per_rank, cudaStream_t st) {
    if (ctx->world == 1) {
        if (dst != src) return cudaMemcpyAsync(dst, src, n_per_rank * sizeof(float), cudaMemcpyDeviceToDevice, st) == cudaSuccess;
        return true;
    }
    if (!ctx->nccl_comm) { zg_set_error("allgather op without zg_cuda_comm_init"); return false; }
    ZG_COUNT_LAUNCH();
    return check(g_nccl.all_gather(src, dst, n_per_rank, kNcclFloat32, ctx->nccl_comm, st), "ncclAllGather");
}
