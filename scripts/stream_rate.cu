// Micro-benchmark: HBM read bandwidth of a GEMV-like access pattern on B200 as a function of bytes in flight.
// Each warp streams its own contiguous run of `run_bytes`; per round it has U x 512 B (one LDG.128 per lane)
// outstanding.  Buffers rotate over 1 GB so nothing is served by L2.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE> __device__ __forceinline__ uint4 ld(const uint4* p) {
    uint4 r;
    if (MODE == 0) asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
template <int U, int MODE>
__global__ void k(const uint4* base, size_t run_units, uint32_t* out) {
    const size_t gw = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint4* p = base + gw * run_units + (threadIdx.x & 31);
    uint4 r[U];
    uint32_t acc = 0;
    const size_t rounds = run_units / (32 * U);
#pragma unroll
    for (int u = 0; u < U; u++) r[u] = ld<MODE>(p + u * 32);
    for (size_t i = 1; i <= rounds; i++) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            acc ^= r[u].x ^ r[u].y ^ r[u].z ^ r[u].w;
            if (i < rounds) r[u] = ld<MODE>(p + (i * U + u) * 32);
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}
template <int U, int MODE>
void run(const uint4* buf, size_t buf_bytes, int ctas_per_sm, int warps, size_t run_bytes, uint32_t* out) {
    int grid = 148 * ctas_per_sm;
    size_t per_launch = (size_t)grid * warps * run_bytes;
    int copies = (int)(buf_bytes / per_launch);
    if (copies < 1) { printf("buffer too small\n"); return; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int c = 0; c < copies && c < 3; c++) k<U, MODE><<<grid, warps * 32>>>(buf + (size_t)c * per_launch / 16, run_bytes / 16, out);
    cudaEventRecord(e0);
    int n = 0;
    for (int rep = 0; rep < 4; rep++)
        for (int c = 0; c < copies; c++, n++) k<U, MODE><<<grid, warps * 32>>>(buf + (size_t)c * per_launch / 16, run_bytes / 16, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double us = ms * 1e3 / n;
    printf("mode %d U %2d ctas/SM %d warps %2d run %6zu B: %5.1f MB/launch %7.2f us/launch %7.1f GB/s  in-flight/SM %5.1f KB  %s\n", MODE, U, ctas_per_sm, warps,
           run_bytes, per_launch / 1e6, us, per_launch / us / 1e3, ctas_per_sm * warps * U * 0.5, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    size_t buf_bytes = (size_t)1 << 30;
    uint4* buf; cudaMalloc(&buf, buf_bytes); cudaMemset(buf, 1, buf_bytes);
    uint32_t* out; cudaMalloc(&out, 4);
    // ~62 MB per launch (like 4096x14336 q8_0): 148*3*8 warps x 16 KB = 58 MB
    run<2, 0>(buf, buf_bytes, 3, 8, 16384, out);
    run<4, 0>(buf, buf_bytes, 3, 8, 16384, out);
    run<8, 0>(buf, buf_bytes, 3, 8, 16384, out);
    run<16, 0>(buf, buf_bytes, 3, 8, 16384, out);
    run<4, 1>(buf, buf_bytes, 3, 8, 16384, out);
    run<8, 1>(buf, buf_bytes, 3, 8, 16384, out);
    run<4, 2>(buf, buf_bytes, 3, 8, 16384, out);
    run<8, 2>(buf, buf_bytes, 3, 8, 16384, out);
    // fewer warps, deeper
    run<8, 0>(buf, buf_bytes, 1, 8, 49152, out);
    run<16, 0>(buf, buf_bytes, 1, 8, 49152, out);
    run<8, 0>(buf, buf_bytes, 2, 8, 24576, out);
    run<16, 0>(buf, buf_bytes, 2, 8, 24576, out);
    run<4, 0>(buf, buf_bytes, 2, 16, 12288, out);
    run<8, 0>(buf, buf_bytes, 2, 16, 12288, out);
    // small launches (like 4096x4096 q8_0 = 18 MB, q4_0 = 9.5 MB)
    run<4, 0>(buf, buf_bytes, 3, 8, 5120, out);
    run<8, 0>(buf, buf_bytes, 3, 8, 5120, out);
    run<4, 0>(buf, buf_bytes, 3, 8, 2560, out);
    run<8, 0>(buf, buf_bytes, 2, 16, 2048, out);
    // big launches
    run<8, 0>(buf, buf_bytes, 3, 8, 65536, out);
    run<8, 1>(buf, buf_bytes, 3, 8, 65536, out);
    return 0;
}
