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
// record covers ZG_TN = 64 output columns (2 quant blocks) x ZG_KC = 64 k-rows:
//
//   q area   int8 formats: 256 units of 16 B; unit u = i*64 + rq*4 + cg holds
//            row 4*rq+i (i in 0..3, rq in 0..15), columns cg*16..cg*16+15, each
//            byte = q + 128 (biased to u8 so a PRMT builds 2^23+u directly).
//            int4 format : 128 units of 16 B; unit u = i*32 + rq*2 + nb holds
//            row 4*rq+i, quant block nb; byte j = (q[j]+8) | (q[j+16]+8) << 4.
//   s area   [rq 16][nb 2][i 4] scales (f16 or f32), scale of row 4*rq+i, block nb.
//
// Records are laid out [n_tile][k_chunk], so the rows one CTA reduces over for one
// column tile are ONE contiguous span -> a single cp.async.bulk per pipeline stage.
// Rows >= K and columns >= N are zero-padded (q = 0, scale = 0).
#define ZG_TN 64
#define ZG_KC 64

struct ZgCudaQWeight {
    int fmt = 0;                 // ZG_QFMT_*
    size_t K = 0, N = 0, bs = 0; // logical [K, N], block size
    // fast formats
    uint32_t n_tiles = 0, n_kc = 0, rec_bytes = 0, q_bytes = 0;
    uint8_t* recs = nullptr;
    // generic format: flat copies
    int8_t* g_data = nullptr;
    float* g_scales = nullptr;
    size_t device_bytes = 0;
};

static inline uint32_t zg_rec_q_bytes(int fmt) { return fmt == ZG_QFMT_I4_F16 ? 2048u : 4096u; }
static inline uint32_t zg_rec_s_bytes(int fmt) { return fmt == ZG_QFMT_I8_F32 ? 512u : 256u; }

// split-K scratch: partial sums [split][M][Np] + one arrival counter per column tile
struct ZgGemvWs {
    float* partials = nullptr;
    size_t partials_elems = 0;
    uint32_t* counters = nullptr;
    size_t counters_n = 0;
};

struct ZgCudaCtx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool owns_stream = true;
    bool graph_mode = true;
    bool profiling = false;
    ZgGemvWs ws; // split-K workspace for the direct zg_cuda_qmatmul_* calls
};

// Grow-only (re)allocation; counters are zero-filled.  Never call between a graph
// capture and its replays: programs own a workspace sized once at compile time.
bool zg_gemv_ws_reserve(ZgGemvWs* ws, size_t partial_elems, size_t counters, cudaStream_t st);
void zg_gemv_ws_free(ZgGemvWs* ws);

// qweight.cu
ZgCudaQWeight* zg_qweight_from_device_flat(ZgCudaCtx* ctx, const int8_t* d_data, const float* d_scales,
                                           size_t K, size_t N, size_t bs, int fmt_hint);
// qgemv.cu
// Plan the split-K geometry for a weight (records per CTA etc.) and launch.
struct ZgGemvPlan {
    uint32_t n_splits = 1, rec_per_cta = 0, smem_bytes = 0, grid = 0, m_block = 1;
};
ZgGemvPlan zg_qgemv_plan(const ZgCudaCtx* ctx, const ZgCudaQWeight* w, uint32_t M);
void zg_qgemv_ws_need(const ZgCudaCtx* ctx, const ZgCudaQWeight* w, uint32_t M, size_t* partial_elems,
                      size_t* counters);
bool zg_qmatmul_launch(ZgCudaCtx* ctx, const ZgCudaQWeight* w, const float* d_in, float* d_out,
                       uint32_t M, uint32_t in_rs, uint32_t out_rs, const ZgGemvWs* ws, cudaStream_t st);
bool zg_qgemv_init(ZgCudaCtx* ctx);

// ops.cu : one launcher per DeviceOp tag (buffers = device pointer table)
struct ZgDevStep { uint32_t op, is_swapped; const float* sec; };
bool zg_launch_op(ZgCudaCtx* ctx, const ZgOp& op, float* const* bufs, const uint32_t* d_dyn,
                  uint32_t op_index, const ZgDevStep* d_steps, cudaStream_t st);
