// GPU-resident packed quantized weights: upload, lossless format selection,
// repacking into the record layout described in zg_internal.cuh, GGUF block
// import and dequantization back to f32 (bit-exact with the reference).
//
// Reference semantics restated on device:
//   QuantizedWeightUpload           src/backend.zig:260-266
//   scale index (k*N+n)/block_size  src/quant.zig:525, src/backend/reference.zig:547
//   quantizedWeightFromInfo         src/models/gguf_loader.zig:99-154 (zgml nibble order)
//   dequantizeTo                    src/quant.zig:594-618
#include "zg_internal.cuh"
#include <math.h>

namespace {

__global__ void k_detect_format(const int8_t* __restrict__ data, size_t n_data,
                                const float* __restrict__ scales, size_t n_scales,
                                uint32_t* __restrict__ flags) {
    // flags bit0: some scale is not exactly an f16; bit1: some q outside [-8, 7]
    uint32_t f = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_scales; i += stride) {
        float s = scales[i];
        float r = __half2float(__float2half_rn(s));
        if (__float_as_uint(r) != __float_as_uint(s)) f |= 1u;
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_data; i += stride) {
        int q = data[i];
        if (q < -8 || q > 7) f |= 2u;
    }
    f = __reduce_or_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0 && f) atomicOr(flags, f);
}

// int8 formats: one thread per 16-byte unit = one lane's A fragment of one column tile.
__global__ void k_pack_q8(const int8_t* __restrict__ data, uint8_t* __restrict__ recs,
                          uint32_t rec_bytes, uint32_t n_kc, size_t K, size_t N, size_t n_units) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_units) return;
    size_t rec = gid >> 6;
    uint32_t u = (uint32_t)(gid & 63);
    size_t nb = rec / n_kc, kc = rec % n_kc;
    uint32_t ct = u >> 5, L = u & 31, g = L >> 2, t = L & 3;
    uint32_t w[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        size_t n = nb * ZG_TN + ct * 16 + g + 8 * (r & 1);
        uint32_t x = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            size_t k = kc * ZG_KR + 4 * t + b + 16 * (r >> 1);
            uint32_t q = (k < K && n < N) ? (uint32_t)(uint8_t)data[k * N + n] : 0u;
            x |= q << (8 * b);
        }
        w[r] = x;
    }
    *reinterpret_cast<uint4*>(recs + rec * rec_bytes + (size_t)u * 16) = make_uint4(w[0], w[1], w[2], w[3]);
}

// int4 format: one thread per 16-byte unit = one lane's packed fragments of both column tiles.
// Nibbles are the biased GGUF form u = q + 8 (padding: q = 0 -> u = 8).
__global__ void k_pack_q4(const int8_t* __restrict__ data, uint8_t* __restrict__ recs,
                          uint32_t rec_bytes, uint32_t n_kc, size_t K, size_t N, size_t n_units) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_units) return;
    size_t rec = gid >> 5;
    uint32_t L = (uint32_t)(gid & 31), g = L >> 2, t = L & 3;
    size_t nb = rec / n_kc, kc = rec % n_kc;
    uint32_t w[4];
#pragma unroll
    for (int wi = 0; wi < 4; wi++) {
        uint32_t ct = wi >> 1, half = wi & 1;
        size_t n_lo = nb * ZG_TN + ct * 16 + g, n_hi = n_lo + 8;
        uint32_t x = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            size_t k = kc * ZG_KR + 4 * t + b + 16 * half;
            uint32_t lo = (k < K && n_lo < N) ? (uint32_t)(data[k * N + n_lo] + 8) : 8u;
            uint32_t hi = (k < K && n_hi < N) ? (uint32_t)(data[k * N + n_hi] + 8) : 8u;
            x |= ((lo & 0xFu) | ((hi & 0xFu) << 4)) << (8 * b);
        }
        w[wi] = x;
    }
    *reinterpret_cast<uint4*>(recs + rec * rec_bytes + (size_t)L * 16) = make_uint4(w[0], w[1], w[2], w[3]);
}

// scale slot idx = 8*t + i  <->  record row kk (zg_internal.cuh)
__host__ __device__ inline uint32_t zg_scale_slot_row(uint32_t idx) {
    uint32_t t = idx >> 3, i = idx & 7;
    return i < 4 ? 4 * t + i : 16 + 4 * t + (i - 4);
}
__host__ __device__ inline uint32_t zg_scale_row_slot(uint32_t kk) {
    uint32_t half = kk >> 4, t = (kk & 15) >> 2, b = kk & 3;
    return 8 * t + 4 * half + b;
}

template <typename ST>
__global__ void k_pack_scales(const float* __restrict__ scales, uint8_t* __restrict__ recs,
                              uint32_t rec_bytes, uint32_t q_bytes, uint32_t n_kc, size_t K, size_t N,
                              size_t n_total) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_total) return;
    size_t rec = gid >> 5;
    uint32_t idx = (uint32_t)(gid & 31);
    size_t nb = rec / n_kc, kc = rec % n_kc;
    size_t row = kc * ZG_KR + zg_scale_slot_row(idx);
    float s = 0.0f;
    if (row < K) s = scales[row * (N / 32) + nb];
    ST* dst = reinterpret_cast<ST*>(recs + rec * rec_bytes + q_bytes) + idx;
    if constexpr (sizeof(ST) == 2) *dst = __float2half_rn(s); // exact: format was verified
    else *dst = s;
}

// max scale per 32-column quant block over all k (|s|: zgml scales are positive, but be safe).
__global__ void k_scale_max(const float* __restrict__ scales, float* __restrict__ smax, size_t K, size_t nbN,
                            size_t nb_padded) {
    size_t nb = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (nb >= nb_padded) return;
    float m = 0.0f;
    if (nb < nbN)
        for (size_t k = 0; k < K; k++) m = fmaxf(m, fabsf(scales[k * nbN + nb]));
    // rounded UP to a power of two (so that per-column-group rescaling of the staged activations is exact) and
    // clamped to [2^-60, 2^60] (an all-zero group still gets a finite reciprocal; larger scales: inf -> NaN out)
    float p2 = 8.673617379884035e-19f; // 2^-60
    if (m > p2) {
        int e;
        float f = frexpf(m, &e);           // m = f * 2^e, f in [0.5, 1)
        p2 = (f == 0.5f) ? ldexpf(1.0f, e - 1) : ldexpf(1.0f, e);
    }
    if (!(m <= 1.152921504606847e18f)) p2 = m; // > 2^60, inf or NaN: keep (poisons the outputs like the reference)
    smax[nb] = p2;
}

// the record after the last one: q = 0 everywhere (int4: biased nibbles 8), scales 0 -> contributes exactly 0;
// the matvec kernel points idle ring slots at it instead of branching
__global__ void k_fill_dummy(uint8_t* __restrict__ rec, uint32_t q_bytes, uint32_t rec_bytes, int is_i4) {
    for (uint32_t i = threadIdx.x; i < rec_bytes; i += blockDim.x) rec[i] = (i < q_bytes && is_i4) ? 0x88 : 0x00;
}

// GGUF raw blocks -> flat i8 + f32 scales (src/models/gguf_loader.zig:117-145).
__global__ void k_gguf_expand(const uint8_t* __restrict__ raw, uint32_t ggml_type, size_t n_blocks,
                              size_t n_elems, int8_t* __restrict__ data, float* __restrict__ scales) {
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    size_t elems = n_elems - b * 32 < 32 ? n_elems - b * 32 : 32;
    if (ggml_type == 8) {
        const uint8_t* blk = raw + b * 34;
        __half_raw hr; hr.x = (unsigned short)(blk[0] | (blk[1] << 8));
        scales[b] = __half2float(__half(hr));
        for (size_t i = 0; i < elems; i++) data[b * 32 + i] = (int8_t)blk[2 + i];
    } else {
        const uint8_t* blk = raw + b * 18;
        __half_raw hr; hr.x = (unsigned short)(blk[0] | (blk[1] << 8));
        scales[b] = __half2float(__half(hr));
        for (size_t i = 0; i < elems; i++) {
            uint8_t byte = blk[2 + i / 2];
            uint8_t nib = (i % 2 == 0) ? (byte & 0x0F) : (byte >> 4);
            data[b * 32 + i] = (int8_t)((int)nib - 8);
        }
    }
}

// dequantizeTo from the packed residency (one thread per weight).
__global__ void k_dequant_packed(const uint8_t* __restrict__ recs, int fmt, uint32_t rec_bytes,
                                 uint32_t q_bytes, uint32_t n_kc, size_t K, size_t N,
                                 float* __restrict__ out) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= K * N) return;
    size_t k = gid / N, n = gid % N;
    size_t nb = n / ZG_TN, kc = k / ZG_KR;
    uint32_t kk = (uint32_t)(k % ZG_KR), c = (uint32_t)(n % ZG_TN);
    uint32_t ct = c >> 4, cn = c & 15, g = cn & 7, r0 = cn >> 3;
    uint32_t half = kk >> 4, t = (kk & 15) >> 2, b = kk & 3, L = 4 * g + t;
    const uint8_t* rec = recs + (nb * n_kc + kc) * (size_t)rec_bytes;
    int q;
    if (fmt == ZG_QFMT_I4_F16) {
        uint32_t wi = ct * 2 + half;
        uint8_t byte = rec[L * 16 + wi * 4 + b];
        int nib = r0 ? (byte >> 4) : (byte & 0xF);
        q = nib - 8; // biased nibble, src/models/gguf_loader.zig:137-141
    } else {
        uint32_t r = r0 + 2 * half;
        q = (int)(int8_t)rec[ct * 512 + L * 16 + r * 4 + b];
    }
    uint32_t sidx = zg_scale_row_slot(kk);
    float s;
    if (fmt == ZG_QFMT_I8_F32) s = reinterpret_cast<const float*>(rec + q_bytes)[sidx];
    else s = __half2float(reinterpret_cast<const __half*>(rec + q_bytes)[sidx]);
    out[gid] = (float)q * s; // f32(q) * scale, src/quant.zig:612-615
}

__global__ void k_dequant_flat(const int8_t* __restrict__ data, const float* __restrict__ scales,
                               size_t n, size_t bs, float* __restrict__ out) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n) return;
    out[gid] = (float)data[gid] * scales[gid / bs];
}

inline unsigned grid_for(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

} // namespace

ZgCudaQWeight* zg_qweight_from_device_flat(ZgCudaCtx* ctx, const int8_t* d_data, const float* d_scales,
                                           size_t K, size_t N, size_t bs, int fmt_hint) {
    cudaStream_t st = ctx->stream;
    size_t n_elems = K * N;
    size_t n_blocks = (n_elems + bs - 1) / bs;
    bool fast_ok = (bs == 32) && (N % 32 == 0) && K > 0 && N > 0;
    int fmt = ZG_QFMT_GENERIC;
    if (fast_ok && fmt_hint != ZG_QFMT_GENERIC) {
        uint32_t* d_flags = nullptr;
        uint32_t flags = 0;
        if (cudaMalloc(&d_flags, 4) != cudaSuccess) { zg_set_error("cudaMalloc flags failed"); return nullptr; }
        cudaMemsetAsync(d_flags, 0, 4, st);
        unsigned blocks = (unsigned)((n_elems / 16 + 255) / 256);
        if (blocks < 1) blocks = 1;
        if (blocks > 4096) blocks = 4096;
        k_detect_format<<<blocks, 256, 0, st>>>(d_data, n_elems, d_scales, n_blocks, d_flags);
        ZG_COUNT_LAUNCH();
        cudaMemcpyAsync(&flags, d_flags, 4, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        cudaFree(d_flags);
        int best = (flags & 1u) ? ZG_QFMT_I8_F32 : ((flags & 2u) ? ZG_QFMT_I8_F16 : ZG_QFMT_I4_F16);
        if (fmt_hint == ZG_QFMT_AUTO) {
            fmt = best;
        } else {
            // A hint may ask for a wider (still lossless) format, never a lossy one.
            bool ok = (fmt_hint == ZG_QFMT_I8_F32) || (fmt_hint == ZG_QFMT_I8_F16 && !(flags & 1u)) ||
                      (fmt_hint == ZG_QFMT_I4_F16 && flags == 0);
            if (!ok) { zg_set_error("format hint %d would be lossy for this weight", fmt_hint); return nullptr; }
            fmt = fmt_hint;
        }
    }

    ZgCudaQWeight* w = new ZgCudaQWeight();
    w->fmt = fmt; w->K = K; w->N = N; w->bs = bs;
    if (fmt == ZG_QFMT_GENERIC) {
        if (cudaMalloc(&w->g_data, n_elems ? n_elems : 1) != cudaSuccess ||
            cudaMalloc(&w->g_scales, (n_blocks ? n_blocks : 1) * sizeof(float)) != cudaSuccess) {
            zg_set_error("cudaMalloc for generic qweight failed");
            cudaFree(w->g_data); cudaFree(w->g_scales); delete w; return nullptr;
        }
        cudaMemcpyAsync(w->g_data, d_data, n_elems, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(w->g_scales, d_scales, n_blocks * sizeof(float), cudaMemcpyDeviceToDevice, st);
        cudaStreamSynchronize(st);
        w->device_bytes = n_elems + n_blocks * sizeof(float);
        return w;
    }
    w->n_nb = (uint32_t)(N / ZG_TN);
    w->n_kc = (uint32_t)((K + ZG_KR - 1) / ZG_KR);
    w->q_bytes = zg_rec_q_bytes(fmt);
    w->rec_bytes = w->q_bytes + zg_rec_s_bytes(fmt);
    size_t n_rec = (size_t)w->n_nb * w->n_kc;
    size_t bytes = (n_rec + 1) * w->rec_bytes; // + the dummy record
    size_t nb_padded = (size_t)w->n_nb;
    if (cudaMalloc(&w->recs, bytes) != cudaSuccess || cudaMalloc(&w->smax, nb_padded * sizeof(float)) != cudaSuccess) {
        zg_set_error("cudaMalloc(%zu) for packed qweight failed", bytes);
        cudaFree(w->recs); delete w; return nullptr;
    }
    w->device_bytes = bytes + nb_padded * sizeof(float);
    k_scale_max<<<grid_for(nb_padded, 128), 128, 0, st>>>(d_scales, w->smax, K, N / 32, nb_padded);
    ZG_COUNT_LAUNCH();
    if (fmt == ZG_QFMT_I4_F16) {
        size_t n_units = n_rec * 32;
        k_pack_q4<<<grid_for(n_units, 256), 256, 0, st>>>(d_data, w->recs, w->rec_bytes, w->n_kc, K, N, n_units);
    } else {
        size_t n_units = n_rec * 64;
        k_pack_q8<<<grid_for(n_units, 256), 256, 0, st>>>(d_data, w->recs, w->rec_bytes, w->n_kc, K, N, n_units);
    }
    ZG_COUNT_LAUNCH();
    k_fill_dummy<<<1, 128, 0, st>>>(w->recs + n_rec * w->rec_bytes, w->q_bytes, w->rec_bytes, fmt == ZG_QFMT_I4_F16);
    ZG_COUNT_LAUNCH();
    size_t n_s = n_rec * 32;
    if (fmt == ZG_QFMT_I8_F32)
        k_pack_scales<float><<<grid_for(n_s, 256), 256, 0, st>>>(d_scales, w->recs, w->rec_bytes, w->q_bytes, w->n_kc, K, N, n_s);
    else
        k_pack_scales<__half><<<grid_for(n_s, 256), 256, 0, st>>>(d_scales, w->recs, w->rec_bytes, w->q_bytes, w->n_kc, K, N, n_s);
    ZG_COUNT_LAUNCH();
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        zg_set_error("packing qweight failed: %s", cudaGetErrorString(e));
        cudaFree(w->recs); cudaFree(w->smax); delete w; return nullptr;
    }
    return w;
}

extern "C" ZgCudaQWeight* zg_cuda_qweight_upload(ZgCudaCtx* ctx, const ZgQWeight* qw, int fmt_hint) {
    if (!ctx || !qw || qw->block_size == 0) { zg_set_error("qweight_upload: bad arguments"); return nullptr; }
    size_t K = qw->rows, N = qw->cols, bs = qw->block_size;
    size_t n_elems = K * N, n_blocks = (n_elems + bs - 1) / bs;
    if (qw->n_data < n_elems || qw->n_scales < n_blocks) { // src/backend.zig:289-291
        zg_set_error("qweight_upload: data/scales shorter than rows*cols requires");
        return nullptr;
    }
    cudaSetDevice(ctx->device);
    int8_t* d_data = nullptr; float* d_scales = nullptr;
    if (cudaMalloc(&d_data, n_elems ? n_elems : 1) != cudaSuccess ||
        cudaMalloc(&d_scales, (n_blocks ? n_blocks : 1) * sizeof(float)) != cudaSuccess) {
        zg_set_error("qweight_upload: staging cudaMalloc failed");
        cudaFree(d_data); cudaFree(d_scales); return nullptr;
    }
    cudaMemcpyAsync(d_data, qw->data, n_elems, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(d_scales, qw->scales, n_blocks * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
    ZgCudaQWeight* w = zg_qweight_from_device_flat(ctx, d_data, d_scales, K, N, bs, fmt_hint);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_data); cudaFree(d_scales);
    return w;
}

// QuantizedWeight.fromSlice (src/quant.zig:216-256) on device: one warp per flat block of `bs` weights:
// max_abs -> scale = max_abs / 127 (1 when the block is all zero), q = trunc(clamp(v * (127 / max_abs), +-127)).
// Separate IEEE operations, no contraction: data and scales are bit-identical to the reference's.
__global__ void k_from_slice(const float* __restrict__ w, size_t n_elems, uint32_t bs, int8_t* __restrict__ data, float* __restrict__ scales) {
    const size_t b = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const size_t start = b * bs;
    if (start >= n_elems) return;
    const size_t end = start + bs < n_elems ? start + bs : n_elems;
    float mx = 0.0f;
    for (size_t j = start + lane; j < end; j += 32) { const float a = fabsf(w[j]); if (a > mx) mx = a; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float scale = mx > 0.0f ? __fdiv_rn(mx, 127.0f) : 1.0f;
    const float inv = mx > 0.0f ? __fdiv_rn(127.0f, mx) : 0.0f;
    if (lane == 0) scales[b] = scale;
    for (size_t j = start + lane; j < end; j += 32) {
        float q = __fmul_rn(w[j], inv);
        q = q < -127.0f ? -127.0f : (q > 127.0f ? 127.0f : q);
        data[j] = (int8_t)(int)q;   // float -> int conversion truncates toward zero, like @intFromFloat
    }
}

extern "C" ZgCudaQWeight* zg_cuda_qweight_from_f32(ZgCudaCtx* ctx, const float* h_weights, size_t rows, size_t cols, size_t block_size,
                                                   int8_t* h_data_out, float* h_scales_out) {
    if (!ctx || !h_weights || block_size == 0 || block_size > 0xFFFFFFFFull) { zg_set_error("qweight_from_f32: bad arguments"); return nullptr; }
    const size_t n_elems = rows * cols, n_blocks = (n_elems + block_size - 1) / block_size;
    cudaSetDevice(ctx->device);
    float* d_w = nullptr; int8_t* d_data = nullptr; float* d_scales = nullptr;
    if (cudaMalloc(&d_w, (n_elems ? n_elems : 1) * sizeof(float)) != cudaSuccess || cudaMalloc(&d_data, n_elems ? n_elems : 1) != cudaSuccess ||
        cudaMalloc(&d_scales, (n_blocks ? n_blocks : 1) * sizeof(float)) != cudaSuccess) {
        zg_set_error("qweight_from_f32: staging cudaMalloc failed");
        cudaFree(d_w); cudaFree(d_data); cudaFree(d_scales); return nullptr;
    }
    cudaMemcpyAsync(d_w, h_weights, n_elems * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
    if (n_blocks) {
        k_from_slice<<<grid_for(n_blocks * 32, 256), 256, 0, ctx->stream>>>(d_w, n_elems, (uint32_t)block_size, d_data, d_scales);
        ZG_COUNT_LAUNCH();
    }
    if (h_data_out) cudaMemcpyAsync(h_data_out, d_data, n_elems, cudaMemcpyDeviceToHost, ctx->stream);
    if (h_scales_out) cudaMemcpyAsync(h_scales_out, d_scales, n_blocks * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    ZgCudaQWeight* w = zg_qweight_from_device_flat(ctx, d_data, d_scales, rows, cols, block_size, ZG_QFMT_AUTO);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_w); cudaFree(d_data); cudaFree(d_scales);
    return w;
}

extern "C" ZgCudaQWeight* zg_cuda_qweight_upload_gguf(ZgCudaCtx* ctx, const void* raw, size_t raw_bytes,
                                                      uint32_t ggml_type, size_t rows, size_t cols) {
    if (!ctx || !raw) { zg_set_error("qweight_upload_gguf: bad arguments"); return nullptr; }
    if (ggml_type != 8 && ggml_type != 2) { // isDirectQuantizedMatmulType, gguf_loader.zig:95-97
        zg_set_error("qweight_upload_gguf: unsupported ggml type %u (only Q8_0=8, Q4_0=2)", ggml_type);
        return nullptr;
    }
    size_t n_elems = rows * cols;
    size_t n_blocks = (n_elems + 31) / 32;
    size_t need = n_blocks * (ggml_type == 8 ? 34 : 18);
    if (raw_bytes < need) { zg_set_error("qweight_upload_gguf: raw buffer too small"); return nullptr; }
    cudaSetDevice(ctx->device);
    uint8_t* d_raw = nullptr; int8_t* d_data = nullptr; float* d_scales = nullptr;
    if (cudaMalloc(&d_raw, need ? need : 1) != cudaSuccess || cudaMalloc(&d_data, n_elems ? n_elems : 1) != cudaSuccess ||
        cudaMalloc(&d_scales, (n_blocks ? n_blocks : 1) * sizeof(float)) != cudaSuccess) {
        zg_set_error("qweight_upload_gguf: staging cudaMalloc failed");
        cudaFree(d_raw); cudaFree(d_data); cudaFree(d_scales); return nullptr;
    }
    cudaMemcpyAsync(d_raw, raw, need, cudaMemcpyHostToDevice, ctx->stream);
    if (n_blocks) {
        k_gguf_expand<<<grid_for(n_blocks, 128), 128, 0, ctx->stream>>>(d_raw, ggml_type, n_blocks, n_elems, d_data, d_scales);
        ZG_COUNT_LAUNCH();
    }
    ZgCudaQWeight* w = zg_qweight_from_device_flat(ctx, d_data, d_scales, rows, cols, 32, ZG_QFMT_AUTO);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_raw); cudaFree(d_data); cudaFree(d_scales);
    return w;
}

// ── synthetic random-init GGUF blocks, generated on the device ────────────────────────────────────────────────────
// Benchmarks and multi-GPU tests need tens of GB of random-init weights (Llama-3-70B-shape Q4_0: 39 GB of blocks) that are
// THE SAME MODEL at every world size.  Every byte of a tensor's block b is a pure function of (seed, tensor_id, global
// block index), so a rank generates exactly its slab [k0, k1) x [n0, n1) of the global [K, N] tensor — in HBM, without a
// host copy — and any other world size (or the host twin zgml_b200/host/llama.py::synth_gguf_blocks, which feeds the CPU
// oracle) sees the same weights.  Word i (8 bytes) of block b = splitmix64-finalizer(key + 8 * b + i); Q8_0 quants are
// clamped to >= -127 and the f16 scale keeps its random low mantissa bits under a fixed exponent e (as
// host/llama.py::synthetic_gguf_blocks: dequantized magnitude ~ sqrt(6 / K), kaimingUniform, src/nn.zig:91-105).
__host__ __device__ inline uint64_t zg_mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void k_synth_gguf(uint8_t* __restrict__ raw, uint64_t key, uint32_t ggml_type, uint32_t exp_bits, size_t n_full, size_t k0, size_t n0,
                             size_t slab_blocks_per_row, size_t n_blocks) {
    const uint32_t bb = ggml_type == 8 ? 34u : 18u, words = ggml_type == 8 ? 5u : 3u;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_blocks * words) return;
    const size_t lb = idx / words;
    const uint32_t wi = (uint32_t)(idx % words);
    const size_t k = k0 + lb / slab_blocks_per_row, nblk = n0 / 32 + lb % slab_blocks_per_row;
    const uint64_t gb = (uint64_t)k * (n_full / 32) + nblk;
    uint64_t w = zg_mix64(key + 8ull * gb + wi);
    uint8_t* dst = raw + lb * bb + wi * 8;
    for (uint32_t j = 0; j < 8 && wi * 8 + j < bb; j++) {
        uint8_t byte = (uint8_t)(w >> (8 * j));
        const uint32_t pos = wi * 8 + j;
        if (pos == 1) byte = (uint8_t)((byte & 3u) | (exp_bits << 2));
        else if (pos >= 2 && ggml_type == 8 && byte == 0x80) byte = 0x81;   // int8 -128 -> -127
        dst[j] = byte;
    }
}

extern "C" ZgCudaQWeight* zg_cuda_qweight_synth_gguf(ZgCudaCtx* ctx, uint64_t seed, uint64_t tensor_id, uint32_t ggml_type, size_t rows_full,
                                                     size_t cols_full, size_t k0, size_t k1, size_t n0, size_t n1) {
    if (!ctx || (ggml_type != 8 && ggml_type != 2) || k1 <= k0 || n1 <= n0 || k1 > rows_full || n1 > cols_full || (cols_full % 32) || (n0 % 32) || (n1 % 32)) {
        zg_set_error("qweight_synth_gguf: bad arguments (slab [%zu,%zu) x [%zu,%zu) of [%zu,%zu], column bounds must be multiples of 32)", k0, k1, n0, n1, rows_full, cols_full);
        return nullptr;
    }
    cudaSetDevice(ctx->device);
    const size_t rows = k1 - k0, cols = n1 - n0, n_elems = rows * cols, n_blocks = n_elems / 32, bb = ggml_type == 8 ? 34 : 18;
    const double qmax = ggml_type == 8 ? 127.0 : 7.0;
    int e = (int)floor(log2(sqrt(6.0 / (double)rows_full) / qmax)) - 1 + 15;   // biased f16 exponent, as synthetic_gguf_blocks
    if (e < 1) e = 1;
    if (e > 30) e = 30;
    uint8_t* d_raw = nullptr; int8_t* d_data = nullptr; float* d_scales = nullptr;
    if (cudaMalloc(&d_raw, n_blocks * bb) != cudaSuccess || cudaMalloc(&d_data, n_elems) != cudaSuccess || cudaMalloc(&d_scales, n_blocks * sizeof(float)) != cudaSuccess) {
        zg_set_error("qweight_synth_gguf: staging cudaMalloc failed");
        cudaFree(d_raw); cudaFree(d_data); cudaFree(d_scales); return nullptr;
    }
    const uint64_t key = zg_mix64(seed * 0x9E3779B97F4A7C15ull + tensor_id);
    const size_t threads = n_blocks * (ggml_type == 8 ? 5 : 3);
    k_synth_gguf<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(d_raw, key, ggml_type, (uint32_t)e, cols_full, k0, n0, cols / 32, n_blocks);
    ZG_COUNT_LAUNCH();
    k_gguf_expand<<<grid_for(n_blocks, 128), 128, 0, ctx->stream>>>(d_raw, ggml_type, n_blocks, n_elems, d_data, d_scales);
    ZG_COUNT_LAUNCH();
    ZgCudaQWeight* w = zg_qweight_from_device_flat(ctx, d_data, d_scales, rows, cols, 32, ZG_QFMT_AUTO);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_raw); cudaFree(d_data); cudaFree(d_scales);
    return w;
}

extern "C" void zg_cuda_qweight_free(ZgCudaCtx* ctx, ZgCudaQWeight* w) {
    if (!w) return;
    if (ctx) cudaSetDevice(ctx->device);
    cudaFree(w->recs); cudaFree(w->smax); cudaFree(w->g_data); cudaFree(w->g_scales);
    cudaFree(w->t_data); cudaFree(w->t_scales); cudaFree(w->x_q); cudaFree(w->x_s);
    delete w;
}

extern "C" int zg_cuda_qweight_format(const ZgCudaQWeight* w) { return w ? w->fmt : -1; }
extern "C" size_t zg_cuda_qweight_device_bytes(const ZgCudaQWeight* w) { return w ? w->device_bytes : 0; }

bool zg_qweight_dequant_to_device(ZgCudaCtx* ctx, const ZgCudaQWeight* w, float* d_out) {
    const size_t n = w->K * w->N;
    if (n == 0) return true;
    if (w->fmt == ZG_QFMT_GENERIC)
        k_dequant_flat<<<grid_for(n, 256), 256, 0, ctx->stream>>>(w->g_data, w->g_scales, n, w->bs, d_out);
    else
        k_dequant_packed<<<grid_for(n, 256), 256, 0, ctx->stream>>>(w->recs, w->fmt, w->rec_bytes, w->q_bytes, w->n_kc, w->K, w->N, d_out);
    ZG_COUNT_LAUNCH();
    return cudaGetLastError() == cudaSuccess;
}

extern "C" int zg_cuda_qweight_dequantize(ZgCudaCtx* ctx, const ZgCudaQWeight* w, float* host_dst) {
    if (!ctx || !w || !host_dst) { zg_set_error("qweight_dequantize: bad arguments"); return -1; }
    cudaSetDevice(ctx->device);
    size_t n = w->K * w->N;
    if (n == 0) return 0;
    float* d_out = nullptr;
    if (cudaMalloc(&d_out, n * sizeof(float)) != cudaSuccess) { zg_set_error("qweight_dequantize: cudaMalloc failed"); return -1; }
    zg_qweight_dequant_to_device(ctx, w, d_out);
    cudaMemcpyAsync(host_dst, d_out, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_out);
    if (e != cudaSuccess) { zg_set_error("qweight_dequantize: %s", cudaGetErrorString(e)); return -1; }
    return 0;
}
