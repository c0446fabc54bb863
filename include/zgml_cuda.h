/*
 * zgml_cuda.h — C-ABI of the B200 (sm_100a) CUDA backend for zgml's
 * DeviceProgram interface.
 *
 * This is the drop-in boundary.  A thin `src/backend/cuda.zig` (see
 * zig/cuda.zig and INTEGRATION.md) flattens zgml's Zig slices / tagged unions
 * into the PODs below and forwards the six `Backend.VTable` slots
 * (reference src/backend.zig:339-352) to the `zg_cuda_*` entry points, exactly
 * as src/backend/metal.zig wraps src/backend/metal_shim.h.
 *
 * Plain pointers and sizes only; no C++ / torch types.  All offsets and strides
 * are in f32 ELEMENTS unless a field says bytes (ZgIO is bytes, like ProgramIO).
 */
#ifndef ZGML_CUDA_H
#define ZGML_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ── zgml `Op` ordinals (reference src/op.zig:11-61, declaration order) ───── */
enum {
    ZG_EW_NONE = 0, ZG_EW_VIEW = 1, ZG_EW_RESHAPE = 2, ZG_EW_TRANSPOSE = 3,
    ZG_EW_PERMUTE = 4, ZG_EW_AS_STRIDED = 5, ZG_EW_BROADCAST_TO = 6,
    ZG_EW_ADD = 7, ZG_EW_MUL = 8,
    ZG_EW_NEG = 9, ZG_EW_ABS = 10, ZG_EW_SGN = 11, ZG_EW_STEP = 12, ZG_EW_RELU = 13,
    ZG_EW_SQRT = 14, ZG_EW_RECIP = 15, ZG_EW_EXP = 16, ZG_EW_LOG = 17, ZG_EW_GELU = 18,
    ZG_EW_SUM = 19, ZG_EW_MAX = 20, ZG_EW_REPEAT = 21
};

/* ── DeviceOp tags (reference src/backend.zig:179-249, union field order) ─── */
typedef enum ZgOpTag {
    ZG_OP_ELEMENTWISE = 0,
    ZG_OP_MATMUL = 1,
    ZG_OP_QMATMUL = 2,
    ZG_OP_SOFTMAX = 3,
    ZG_OP_LAYERNORM = 4,
    ZG_OP_RMSNORM = 5,
    ZG_OP_REDUCE = 6,
    ZG_OP_REPEAT = 7,
    ZG_OP_SLICE_ASSIGN = 8,
    ZG_OP_ROPE = 9,
    ZG_OP_ATTENTION = 10,
    ZG_OP_FUSED_ELEMENTWISE = 11,
    ZG_OP_COUNT = 12,
    /* Extensions (no zgml counterpart: the reference has no distributed layer).  Emitted only by
     * the sharded session of this repo; need zg_cuda_comm_init().  One NCCL collective each, in program order. */
    ZG_OP_ALLREDUCE = 12, /* buf[offset .. offset + n) = sum over ranks, in place */
    ZG_OP_ALLGATHER = 13  /* dst[dst_offset + r * n ..) = rank r's src[src_offset .. + n), every rank */
} ZgOpTag;

/* MatMulGeometry — reference src/backend.zig:146-158 (all usize). */
typedef struct ZgMatMulGeometry {
    size_t M, N, K;
    size_t a_row_stride, a_col_stride;
    size_t b_row_stride, b_col_stride;
    size_t a_offset, b_offset, dst_offset, dst_row_stride;
} ZgMatMulGeometry;

/* FusedEwStep — reference src/backend.zig:170-175. */
typedef struct ZgFusedEwStep {
    uint32_t op;               /* ZG_EW_* */
    uint32_t is_swapped;       /* chain value sits in src1 position */
    uint32_t secondary_buf;    /* u16 in zgml */
    uint32_t secondary_offset;
} ZgFusedEwStep;

/* One DeviceOp.  Field names and meaning copied 1:1 from src/backend.zig:180-248;
 * u16 buffer indices are widened to u32. */
typedef struct ZgOp {
    uint32_t tag; /* ZgOpTag */
    uint32_t _pad;
    union {
        struct { uint32_t op, dst, src0, src1, n, dst_offset, src0_offset, src1_offset; } elementwise;
        struct { uint32_t dst, a, b, _pad; ZgMatMulGeometry geom; } matmul;
        struct {
            uint32_t dst, input, weight_idx, M, N, K;
            uint32_t input_offset, input_row_stride; /* stride 0 => K */
            uint32_t dst_offset, dst_row_stride;     /* stride 0 => N */
        } qmatmul;
        struct { uint32_t dst, src, rows, cols, src_offset, dst_offset; } softmax;
        struct { uint32_t dst, src, rows, cols; float eps; uint32_t src_offset, dst_offset; } layernorm;
        struct { uint32_t dst, src, rows, cols; float eps; uint32_t src_offset, dst_offset; } rmsnorm;
        struct { uint32_t op, dst, src, n_out, reduce_size, src_offset, dst_offset; } reduce;
        struct {
            uint32_t dst, src, n;
            uint32_t src_ne[4], dst_ne[4], src_strides[4], dst_strides[4];
            uint32_t src_offset, dst_offset;
        } repeat;
        struct {
            uint32_t dst, src, rows, cols;
            uint32_t dst_base_offset, dst_offset, dst_row_stride, dst_col_stride;
            uint32_t src_offset, src_row_stride, src_col_stride, patch_stride;
        } slice_assign;
        struct {
            uint32_t dst, src, cos_sin, half_d, seq_len;
            uint32_t src_off, cs_off, dst_off, src_rs, src_cs, cs_cs;
        } rope;
        struct {
            uint32_t dst, q, k, v, mask, has_mask;
            uint32_t d_head, seq_q, seq_kv;
            float scale;
            uint32_t q_off, k_off, v_off, mask_off, dst_off;
            uint32_t q_rs, q_cs, k_rs, k_cs, v_rs, v_cs, mask_rs, mask_cs, dst_rs, dst_cs;
        } attention;
        struct {
            const ZgFusedEwStep* steps;
            size_t n_steps;
            uint32_t n, dst, src, dst_offset, src_offset;
        } fused_elementwise;
        struct { uint32_t buf, offset, n; } allreduce;
        struct { uint32_t dst, src, n, dst_offset, src_offset; } allgather;
    } u;
} ZgOp;

/* ProgramIO — reference src/backend.zig:252-257.  offset/size in BYTES. */
typedef struct ZgIO {
    uint32_t buf_idx;
    uint32_t offset;
    void* host_ptr;
    uint32_t size;
    uint32_t _pad;
} ZgIO;

/* QuantizedWeightUpload — reference src/backend.zig:260-266.
 * data: i8 [rows*cols] row-major [K=rows, N=cols]; scales: f32, one per
 * `block_size` consecutive FLAT elements (scale index = (k*N+n)/block_size,
 * reference src/quant.zig:525, src/backend/reference.zig:547). */
typedef struct ZgQWeight {
    const int8_t* data;
    size_t n_data;
    const float* scales;
    size_t n_scales;
    size_t rows, cols, block_size;
} ZgQWeight;

/* Extension (direct GGUF->device upload): a descriptor with block_size == ZG_QWEIGHT_RESIDENT
 * refers to a weight already packed in HBM — `data` is the ZgCudaQWeight* that zg_cuda_qweight_upload /
 * zg_cuda_qweight_upload_gguf returned, rows/cols must match it, the other fields are ignored.  The program
 * borrows it (the caller frees it after zg_cuda_free).  Lets a loader stream a model into HBM one tensor at a
 * time instead of holding every expanded i8 copy in host memory until compile (70B: 69.5 GB). */
#define ZG_QWEIGHT_RESIDENT ((size_t)-1)

/* DeviceProgram — reference src/backend.zig:270-275. buffer_sizes in f32 elements. */
typedef struct ZgProgram {
    const ZgOp* ops;
    size_t n_ops;
    size_t n_buffers;
    const size_t* buffer_sizes;
    const ZgIO* initial_uploads;
    size_t n_uploads;
    const ZgQWeight* qweights;
    size_t n_qweights;
} ZgProgram;

/* Subset of RuntimeProfile (reference src/profile.zig:819-842) that a CUDA
 * backend can fill: per-tag device time + counters. */
typedef struct ZgProfile {
    uint64_t time_ns[ZG_OP_COUNT];
    uint64_t backend_op_count;
    uint64_t fallback_op_count;     /* always 0: there is no CPU fallback */
    uint64_t backend_dispatch_count; /* kernels / graph launches issued */
    uint64_t sync_time_ns;
    uint64_t sync_count;
    uint32_t call_count;
    uint32_t _pad;
} ZgProfile;

/* Capabilities to advertise (reference src/backend.zig:14-70). */
typedef struct ZgCapabilities {
    uint32_t compiled_programs, host_visible_program_memory, dense_matmul_f32, dense_matmul_f16;
    uint32_t qmatmul, fused_elementwise, max_fused_elementwise_steps /* 0 = unlimited */;
    uint32_t dynamic_program_refresh, prefill_attention, decode_attention, quantized_kv;
    uint32_t attention_supported, attention_max_seq_kv /* 0 = unlimited */, attention_max_d_head;
} ZgCapabilities;

typedef struct ZgCudaCtx ZgCudaCtx;
typedef struct ZgCudaProgram ZgCudaProgram;

/* ── The six vtable slots + lifecycle ─────────────────────────────────────── */

/* Backend construction (MetalBackend.init analogue, src/backend/metal.zig).
 * NULL when no CUDA device / wrong arch; see zg_cuda_last_error(). */
ZgCudaCtx* zg_cuda_create(int device_ordinal);
void zg_cuda_destroy(ZgCudaCtx* ctx);
void zg_cuda_capabilities(ZgCapabilities* out);

/* VTable.dense_matmul_f32 (src/backend.zig:341): host-pointer GEMM override
 * during graph execution.  Always declines (returns 0) like the fake backend in
 * src/device_inference.zig:750-752: host tensors are not device resident. */
int zg_cuda_dense_matmul_f32(ZgCudaCtx* ctx, float* dst, const float* a, const float* b,
                             const ZgMatMulGeometry* geom);

/* VTable.compile_program (src/backend.zig:343): allocate zero-filled device
 * buffers, apply initial_uploads, copy+repack every qweight into its
 * GPU-resident packed layout, copy the op list.  NULL on any failure. */
ZgCudaProgram* zg_cuda_compile(ZgCudaCtx* ctx, const ZgProgram* program);

/* VTable.refresh_program (src/backend.zig:345): caller's op array after
 * patchSliceAssignOffset / patchAttentionSeqKV (src/device_inference.zig:242-256). */
void zg_cuda_refresh(ZgCudaCtx* ctx, ZgCudaProgram* prog, const ZgOp* ops, size_t n_ops);

/* VTable.execute_program (src/backend.zig:347): upload inputs, run all ops in
 * order, download outputs.  Synchronous: outputs valid on return. */
void zg_cuda_execute(ZgCudaCtx* ctx, ZgCudaProgram* prog, const ZgIO* inputs, size_t n_inputs,
                     const ZgIO* outputs, size_t n_outputs);

/* VTable.free_program (src/backend.zig:349). */
void zg_cuda_free(ZgCudaCtx* ctx, ZgCudaProgram* prog);

/* VTable.get_runtime_profile (src/backend.zig:351). Enabled with
 * zg_cuda_set_profiling(ctx, 1); NULL otherwise. */
const ZgProfile* zg_cuda_profile(ZgCudaCtx* ctx, ZgCudaProgram* prog);
void zg_cuda_set_profiling(ZgCudaCtx* ctx, int enabled);

/* Error convention: NULL / no-op on failure plus this thread-unsafe string. */
const char* zg_cuda_last_error(void);

/* ── Extensions used by tests / bench (not vtable slots) ──────────────────── */

/* Run work on a caller-owned stream (e.g. torch's) instead of the ctx's own. */
void zg_cuda_set_stream(ZgCudaCtx* ctx, void* cuda_stream);
void zg_cuda_sync(ZgCudaCtx* ctx);
/* Kernels launched by this library since process start (claim for gpu_launches). */
uint64_t zg_cuda_launch_count(void);
/* 0: eager launches; 1 (default): capture the op list into a CUDA graph. */
void zg_cuda_set_graph_mode(ZgCudaCtx* ctx, int enabled);
/* Device pointer of program buffer `buf_idx` (for device-resident benches). */
void* zg_cuda_program_buffer(ZgCudaProgram* prog, uint32_t buf_idx);
/* Schedule introspection (tests / benches assert that the fast paths are the ones that run): what == 0: kernels one
 * execution launches (graph nodes); 1: DeviceOps covered by the fused single-token decode kernel (0: general schedule);
 * 2: layers inside that kernel; 3: launches of the streamed single-row matvec kernel (csrc/qgemv_stream.cu) among them;
 * 4: HBM bytes per execution saved by promoted dense operands (zg_cuda_program_promote_dense). */
uint64_t zg_cuda_program_stats(const ZgCudaProgram* prog, int what);
/* Run the program's ops with no host<->device copies (inputs already resident);
 * asynchronous on the ctx stream. */
void zg_cuda_execute_device(ZgCudaCtx* ctx, ZgCudaProgram* prog);

/* Packed, GPU-resident quantized weight (the src/quant.zig QuantizedWeight
 * analogue).  Residency format is chosen losslessly at upload:
 *   ZG_QFMT_I8_F32  36 B / 32 weights  (native fromSlice output)
 *   ZG_QFMT_I8_F16  34 B / 32 weights  (every scale is exactly an f16: Q8_0 origin)
 *   ZG_QFMT_I4_F16  18 B / 32 weights  (additionally all q in [-8,7]: Q4_0 origin)
 *   ZG_QFMT_GENERIC block_size != 32 or N % 32 != 0: flat i8 + f32 scales */
enum { ZG_QFMT_AUTO = 0, ZG_QFMT_I8_F32 = 1, ZG_QFMT_I8_F16 = 2, ZG_QFMT_I4_F16 = 3, ZG_QFMT_GENERIC = 4 };
typedef struct ZgCudaQWeight ZgCudaQWeight;

ZgCudaQWeight* zg_cuda_qweight_upload(ZgCudaCtx* ctx, const ZgQWeight* w, int fmt_hint);
/* Direct GGUF block upload (reference src/models/gguf_loader.zig:99-154 done on
 * device): raw = n_blocks * 34 B (ggml_type 8 = Q8_0) or 18 B (2 = Q4_0);
 * rows = dims[0] = K, cols = dims[1] = N, zgml nibble order. */
ZgCudaQWeight* zg_cuda_qweight_upload_gguf(ZgCudaCtx* ctx, const void* raw, size_t raw_bytes,
                                           uint32_t ggml_type, size_t rows, size_t cols);
/* Extension for benchmarks / multi-GPU tests: random-init GGUF blocks generated IN HBM for the slab
 * [k0, k1) x [n0, n1) of a global [rows_full, cols_full] tensor, then imported like zg_cuda_qweight_upload_gguf.
 * Block bytes are a pure function of (seed, tensor_id, global block index): every world size holds slices of the
 * same model, and zgml_b200/host/llama.py::synth_gguf_blocks reproduces them on the host for the CPU oracle. */
ZgCudaQWeight* zg_cuda_qweight_synth_gguf(ZgCudaCtx* ctx, uint64_t seed, uint64_t tensor_id, uint32_t ggml_type,
                                          size_t rows_full, size_t cols_full, size_t k0, size_t k1, size_t n0, size_t n1);
void zg_cuda_qweight_free(ZgCudaCtx* ctx, ZgCudaQWeight* w);
int zg_cuda_qweight_format(const ZgCudaQWeight* w);
size_t zg_cuda_qweight_device_bytes(const ZgCudaQWeight* w);
/* dequantizeTo (src/quant.zig:594-618) from the packed residency: host_dst gets
 * rows*cols f32, bit-exact with the reference.  Returns 0 on success. */
int zg_cuda_qweight_dequantize(ZgCudaCtx* ctx, const ZgCudaQWeight* w, float* host_dst);
/* QuantizedWeight.matmul (src/quant.zig:475-578) with the DeviceOp.qmatmul
 * stride contract, on DEVICE pointers; async on the ctx stream. */
int zg_cuda_qmatmul_device(ZgCudaCtx* ctx, const ZgCudaQWeight* w, const float* d_input,
                           float* d_dst, uint32_t M, uint32_t input_row_stride,
                           uint32_t dst_row_stride);
/* Same through HOST buffers (copies inside): the e2e leg. Synchronous. */
int zg_cuda_qmatmul_host(ZgCudaCtx* ctx, const ZgCudaQWeight* w, const float* h_input,
                         float* h_dst, uint32_t M);

/* QuantizedWeight.fromSlice / fromTensor — reference src/quant.zig:216-264 — on device: quantizes a host f32 [rows=K, cols=N]
 * row-major matrix per flat block of `block_size` (scale = max_abs / 127, q = trunc(clamp(v * 127 / max_abs))) and packs it.
 * h_data_out [rows*cols] / h_scales_out [ceil(rows*cols / block_size)] (either may be NULL) receive the reference's flat
 * form, bit-identical to the CPU result.  NULL on failure. */
ZgCudaQWeight* zg_cuda_qweight_from_f32(ZgCudaCtx* ctx, const float* h_weights, size_t rows, size_t cols, size_t block_size,
                                        int8_t* h_data_out, float* h_scales_out);
/* QuantizedWeight.matmulBias — reference src/quant.zig:581-589: matmul, then dst[m, n] += bias[n] (host buffers). */
int zg_cuda_qmatmul_bias_host(ZgCudaCtx* ctx, const ZgCudaQWeight* w, const float* h_input, const float* h_bias,
                              float* h_dst, uint32_t M);

/* W8A8 decode path (what `session.quantize()` models run per linear on the reference's aarch64 builds).  All three are
 * integer / IEEE-exact work and bit-identical to the reference.
 * prepareTransposed — reference src/quant.zig:274-317 (twin src/backend/reference.zig:26-70): dequantize [K, N], re-quantize
 * into [N, K] int8 with K-aligned blocks + [N, ceil(K / block_size)] f32 scales; both stay resident with the weight.
 * h_t_data / h_t_scales (either may be NULL) receive copies.  Idempotent.  Returns 0 on success. */
int zg_cuda_qweight_prepare_transposed(ZgCudaCtx* ctx, ZgCudaQWeight* w, int8_t* h_t_data, float* h_t_scales);
/* quantizeInput — reference src/quant.zig:320-341: per K-block scale = max_abs / 127 (1 if zero),
 * q = trunc(clamp(v * (127 / max_abs), +-127)).  Host buffers: h_q [K], h_scales [ceil(K / block_size)].
 * Limits as the reference's stack buffers (src/quant.zig:452-453): K <= 16384, <= 512 blocks. */
int zg_cuda_quantize_input_host(ZgCudaCtx* ctx, const float* h_input, size_t K, size_t block_size, int8_t* h_q, float* h_scales);
/* gemv — reference src/quant.zig:443-459 = quantizeInput + gemvRange (:358-440) over all N outputs, M = 1:
 * dst[n] = sum over K-blocks ascending of f32(int32 dot) * (s_x[b] * s_w[n, b]).  Needs prepare_transposed.
 * _device: device pointers, async on the ctx stream; _host: host buffers, synchronous. */
int zg_cuda_gemv_w8a8_device(ZgCudaCtx* ctx, const ZgCudaQWeight* w, const float* d_input, float* d_dst);
int zg_cuda_gemv_w8a8_host(ZgCudaCtx* ctx, const ZgCudaQWeight* w, const float* h_input, float* h_dst);

/* Quantized KV cache + attention over it — reference src/quant.zig:633-1091, callers
 * src/llama_inference.zig:330-377.  Column-major Q8 cache: column c = d_head int8 + d_head / block_size f32 scales; one
 * column per kv position inside a head's slab.  d_head <= 512 and d_head % block_size == 0 as in the reference; this
 * implementation additionally needs d_head % 4 == 0 and block_size % 4 == 0. */
typedef struct ZgCudaKVCache ZgCudaKVCache;
/* QuantizedKVCache.init / deinit / clear — src/quant.zig:658-686 (zero-filled). NULL on failure. */
ZgCudaKVCache* zg_cuda_kvcache_create(ZgCudaCtx* ctx, size_t d_head, size_t n_cols, size_t block_size);
void zg_cuda_kvcache_free(ZgCudaCtx* ctx, ZgCudaKVCache* cache);
int zg_cuda_kvcache_clear(ZgCudaCtx* ctx, ZgCudaKVCache* cache);
/* storeColumn — src/quant.zig:689-701 — for n_write consecutive columns from src [n_write][d_head] (the slice_assign
 * routing of src/llama_inference.zig:336-348): quantizeInput per column, bit-identical data and scales.
 * _device: device pointer, async on the ctx stream; _host: host buffer, synchronous. */
int zg_cuda_kvcache_store_device(ZgCudaCtx* ctx, ZgCudaKVCache* cache, size_t col_start, size_t n_write, const float* d_src);
int zg_cuda_kvcache_store_host(ZgCudaCtx* ctx, ZgCudaKVCache* cache, size_t col_start, size_t n_write, const float* h_src);
/* whole cache to the host: h_q [d_head * n_cols], h_scales [d_head / block_size * n_cols] (either may be NULL) */
int zg_cuda_kvcache_download(ZgCudaCtx* ctx, const ZgCudaKVCache* cache, int8_t* h_q, float* h_scales);
/* attentionQuantized — src/quant.zig:924-1091, same argument meaning: q / dst column-major [d_head, seq_q] with column
 * strides, kv columns [col_start, col_start + seq_kv) of each cache, optional additive mask [seq_kv, seq_q or 1] with
 * (row, column) strides (column stride 0 broadcasts one column; non-finite entries skip the position; a fully masked
 * query column yields zeros).  int8_query != 0: the aarch64 branch (query quantized per block, int8 x int8 scores);
 * 0: the portable branch (f32 query x int8 keys).  Scores are bit-identical to the reference's; the softmax groups
 * positions differently, so outputs agree to float rounding. */
int zg_cuda_attention_quantized_device(ZgCudaCtx* ctx, float* d_dst, size_t dst_col_stride, const float* d_q, size_t q_col_stride,
                                       size_t d_head, size_t seq_q, const ZgCudaKVCache* k_cache, size_t k_col_start,
                                       const ZgCudaKVCache* v_cache, size_t v_col_start, size_t seq_kv, const float* d_mask,
                                       size_t mask_row_stride, size_t mask_col_stride, float scale, int int8_query);
int zg_cuda_attention_quantized_host(ZgCudaCtx* ctx, float* h_dst, size_t dst_col_stride, const float* h_q, size_t q_col_stride,
                                     size_t d_head, size_t seq_q, const ZgCudaKVCache* k_cache, size_t k_col_start,
                                     const ZgCudaKVCache* v_cache, size_t v_col_start, size_t seq_kv, const float* h_mask,
                                     size_t mask_row_stride, size_t mask_col_stride, float scale, int int8_query);

/* LlamaInferenceSession.quantizeKV for a compiled program — src/llama_inference.zig:648-679 (plan side :277-328): from now on
 * every patched slice_assign into a buffer that attention ops read as K or V stores quantized columns (storeColumn) into a
 * Q8 cache standing in for that buffer, and those attention ops run attentionQuantized over the caches (the routing of
 * src/llama_inference.zig:336-377).  The caches start zeroed, like the reference's; call it right after compile.
 * int8_query != 0: the aarch64 branch of attentionQuantized (query quantized per block), 0: the portable f32-query branch.
 * Returns 0 on success; -1 (program unchanged) when the program's cache accesses are not whole d_head columns. */
int zg_cuda_program_quantize_kv(ZgCudaCtx* ctx, ZgCudaProgram* prog, size_t block_size, int int8_query);

/* Format hint for dense matmul operands (precedent: the WGPU backend's f16 promotion of matmul B buffers that
 * have initial uploads, src/backend/wgpu.zig:1068-1106).  Every matmul op of the program with M == 1 and a k-contiguous B
 * operand that no op writes (the tied LM head x @ token_embed^T, src/models/llama.zig:162-165) gets a 16-bit copy of that operand;
 * the op then streams the copy (half the bytes), bounds each output's error by ||x|| * ||w - round16(w)|| and recomputes every
 * column that could still be the maximum from the f32 original, with the exact kernel's arithmetic.  The argmax — and the
 * winning logits bit for bit — equal the unpromoted program's; the other outputs carry the 16-bit weight rounding.
 * An execute() input that overwrites a promoted operand refreshes its copy.  Returns the number of ops promoted, -1 on error. */
#define ZG_DENSE_BF16 1   /* f32's range, 8 significant bits: other logits within ~2e-3 of the output scale */
#define ZG_DENSE_F16 2    /* the WGPU backend's choice, 11 significant bits: other logits within ~2e-4; |w| > 65504 falls back to the exact row */
int zg_cuda_program_promote_dense(ZgCudaCtx* ctx, ZgCudaProgram* prog, int format);

/* Multi-GPU (one process per GPU): a 128-byte NCCL unique id made on rank 0, distributed by the host
 * (torch.distributed / MPI / file), then one communicator per context.  Returns 0 on success. */
int zg_cuda_comm_unique_id(void* id128);
int zg_cuda_comm_init(ZgCudaCtx* ctx, const void* id128, int rank, int world);
void zg_cuda_comm_destroy(ZgCudaCtx* ctx);
/* 0 = no communicator, 1 = NCCL only, 2 = small all-reduces over NVLink peer memory (cudaIpc) + NCCL for the rest. */
int zg_cuda_comm_mode(const ZgCudaCtx* ctx);

/* Debug timeline of a program's kernels (off by default): records of 3 x u64 {kind << 56 | t_entry, t_after_wait, t_exit},
 * %globaltimer ns stamped by block 0 of each kernel; kinds: 1 elementwise, 2 fused_elementwise, 3 rmsnorm, 4 repeat,
 * 5 slice_assign, 6 rope, 7 attention, 8 chain, 9 dense matmul, 10 quantized matvec. */
int zg_cuda_trace(ZgCudaCtx* ctx, int enable);
size_t zg_cuda_trace_read(ZgCudaCtx* ctx, unsigned long long* host, size_t max_records);

void* zg_cuda_malloc(ZgCudaCtx* ctx, size_t bytes);
void zg_cuda_free_device(ZgCudaCtx* ctx, void* p);
int zg_cuda_memcpy_h2d(ZgCudaCtx* ctx, void* d, const void* h, size_t bytes);
int zg_cuda_memcpy_d2h(ZgCudaCtx* ctx, void* h, const void* d, size_t bytes);
int zg_cuda_memset(ZgCudaCtx* ctx, void* d, int value, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* ZGML_CUDA_H */
