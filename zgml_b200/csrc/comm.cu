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

// Export this rank's slot region, gather everybody's IPC handle through NCCL, map the peers.
static bool peer_setup(ZgCudaCtx* ctx) {
    const int world = ctx->world, rank = ctx->rank;
    if (world > kZgMaxRanks) return false;
    if (const char* e = getenv("ZG_CUDA_PEER")) if (e[0] == '0') return false;
    const size_t slot_bytes = (size_t)kZgPeerSets * world * kPeerSlotFloats * 2 * sizeof(float);   // every float travels with its epoch
    const size_t flag_bytes = 0;   // readiness travels inside the data cells (epoch words)
    const size_t total = slot_bytes + flag_bytes + (2 + kZgPeerCtas) * sizeof(uint32_t) + 64;
    if (cudaMalloc(&ctx->peer_mem, total) != cudaSuccess) { cudaGetLastError(); return false; }
    cudaMemset(ctx->peer_mem, 0, total);
    cudaIpcMemHandle_t mine;
    if (cudaIpcGetMemHandle(&mine, ctx->peer_mem) != cudaSuccess) { cudaGetLastError(); cudaFree(ctx->peer_mem); ctx->peer_mem = nullptr; return false; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    char* d_handles = nullptr;
    std::vector<cudaIpcMemHandle_t> all(world);
    bool ok = cudaMalloc(&d_handles, 64 * (size_t)world) == cudaSuccess;
    ok = ok && cudaMemcpy(d_handles + 64 * rank, &mine, 64, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && check(g_nccl.all_gather(d_handles + 64 * rank, d_handles, 64, kNcclChar, ctx->nccl_comm, ctx->stream), "ncclAllGather(ipc handles)");
    ok = ok && cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    ok = ok && cudaMemcpy(all.data(), d_handles, 64 * (size_t)world, cudaMemcpyDeviceToHost) == cudaSuccess;
    cudaFree(d_handles);
    int mapped_ok = ok ? 1 : 0;
    for (int r = 0; r < world && mapped_ok; r++) {
        if (r == rank) { ctx->peer_mapped[r] = nullptr; continue; }
        if (cudaIpcOpenMemHandle(&ctx->peer_mapped[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ctx->peer_mapped[r] = nullptr; mapped_ok = 0; }
    }
    // every rank must take the same path: agree through a min-reduction
    float h = (float)mapped_ok;
    float* d_f = nullptr;
    if (cudaMalloc(&d_f, sizeof(float)) == cudaSuccess) {
        cudaMemcpy(d_f, &h, sizeof(float), cudaMemcpyHostToDevice);
        g_nccl.all_reduce(d_f, d_f, 1, kNcclFloat32, /*ncclMin*/ 3, ctx->nccl_comm, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(&h, d_f, sizeof(float), cudaMemcpyDeviceToHost);
        cudaFree(d_f);
    } else h = 0.f;
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
    pc.seq = (uint32_t*)((char*)ctx->peer_mem + slot_bytes + flag_bytes);
    return true;
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
