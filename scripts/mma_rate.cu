// Micro-benchmark: issue rate of legacy mma.sync variants on sm_100a (one CTA per SM, W warps, ILP independent chains).
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND, int ILP>
__global__ void k(int iters, int* out) {
    int d[ILP][4];
    float f[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) { d[i][j] = 0; f[i][j] = 0.f; }
    unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(d[i][0]), "+r"(d[i][1]), "+r"(d[i][2]), "+r"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(f[i][0]), "+f"(f[i][1]), "+f"(f[i][2]), "+f"(f[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 2)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(f[i][0]), "+f"(f[i][1]), "+f"(f[i][2]), "+f"(f[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 3)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(f[i][0]), "+f"(f[i][1]), "+f"(f[i][2]), "+f"(f[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 4)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.f32.e4m3.e4m3.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(f[i][0]), "+f"(f[i][1]), "+f"(f[i][2]), "+f"(f[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    int s = 0; float fs = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) { s += d[i][j]; fs += f[i][j]; }
    if (s == 123456789 || fs == 1.2345f) out[0] = s;
}
template <int KIND>
void run(const char* name, int macs_per_mma, int warps) {
    int* out; cudaMalloc(&out, 4);
    const int iters = 2000, ILP = 4;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<KIND, ILP><<<148, warps * 32>>>(10, out);
    cudaEventRecord(e0);
    k<KIND, ILP><<<148, warps * 32>>>(iters, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double n = 148.0 * warps * iters * ILP;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cyc = ms * 1e-3 * clk * 1e3;
    printf("%-10s warps/SM %2d: %.2f ms, %.1f cycles per mma per SM-subpartition (4/SM), %.1f TMAC/s  err=%s\n", name, warps, ms,
           cyc / (iters * ILP * warps / 4.0), n * macs_per_mma / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main() {
    for (int w : {4, 8, 16}) {
        run<0>("imma s8u8", 16 * 8 * 32, w);
        run<1>("hmma f16", 16 * 8 * 16, w);
        run<2>("hmma bf16", 16 * 8 * 16, w);
        run<3>("tf32", 16 * 8 * 8, w);
        run<4>("fp8 e4m3", 16 * 8 * 32, w);
    }
    return 0;
}
