// NCCL collectives for row-sharded linear layers: one process per GPU, one communicator per
// context, collectives enqueued on the program's stream (captured into its CUDA graph like any kernel).
// libnccl is resolved with dlopen at first use, so single-GPU users need no NCCL at all.
#include "zg_internal.cuh"

#include <dlfcn.h>
#include <string.h>

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

constexpr int kNcclFloat32 = 7, kNcclSum = 0;

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
    return 0;
}

extern "C" void zg_cuda_comm_destroy(ZgCudaCtx* ctx) {
    if (!ctx || !ctx->nccl_comm) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
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
