/* The reference's own known-answer program for the path (src/backend/reference.zig:710-761, also
 * src/backend/conformance.zig:81-112), driven through the C-ABI from plain C exactly the way the Zig translator
 * (zig/cuda.zig) would: compile_program -> execute_program -> free_program, then the strided variant.
 * Build: gcc -std=c99 -I include tests/cpp/cabi_golden.c -L zgml_b200/lib -lzgml_cuda -Wl,-rpath,... -lm */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "zgml_cuda.h"

static int close_to(const float* got, const float* want, int n, float tol) {
    for (int i = 0; i < n; i++)
        if (!(fabsf(got[i] - want[i]) <= tol)) { printf("  [%d] got %g want %g\n", i, got[i], want[i]); return 0; }
    return 1;
}

int main(void) {
    ZgCudaCtx* ctx = zg_cuda_create(0);
    if (!ctx) { printf("SKIP: %s\n", zg_cuda_last_error()); return 77; }

    /* data = {2,-1,3, 4,-2,1, -3,5,2}, scales = {.5,.25,1}, bs = 4, K = N = 3 (flat blocks cross row boundaries) */
    const int8_t data[9] = {2, -1, 3, 4, -2, 1, -3, 5, 2};
    const float scales[3] = {0.5f, 0.25f, 1.0f};
    ZgQWeight qw;
    memset(&qw, 0, sizeof(qw));
    qw.data = data; qw.n_data = 9; qw.scales = scales; qw.n_scales = 3; qw.rows = 3; qw.cols = 3; qw.block_size = 4;

    /* 1. dense rows: x = {1,2,3}, {-1,.5,4} -> {2.75,2.25,8.0}, {-3.0,5.25,6.625} */
    {
        float x[6] = {1, 2, 3, -1, 0.5f, 4}, y[6] = {0};
        const float want[6] = {2.75f, 2.25f, 8.0f, -3.0f, 5.25f, 6.625f};
        ZgOp op;
        memset(&op, 0, sizeof(op));
        op.tag = ZG_OP_QMATMUL;
        op.u.qmatmul.dst = 1; op.u.qmatmul.input = 0; op.u.qmatmul.weight_idx = 0;
        op.u.qmatmul.M = 2; op.u.qmatmul.N = 3; op.u.qmatmul.K = 3;
        size_t sizes[2] = {6, 6};
        ZgProgram prog;
        memset(&prog, 0, sizeof(prog));
        prog.ops = &op; prog.n_ops = 1; prog.n_buffers = 2; prog.buffer_sizes = sizes; prog.qweights = &qw; prog.n_qweights = 1;
        ZgCudaProgram* h = zg_cuda_compile(ctx, &prog);
        if (!h) { printf("FAIL compile: %s\n", zg_cuda_last_error()); return 1; }
        ZgIO in = {0, 0, x, sizeof(x), 0}, out = {1, 0, y, sizeof(y), 0};
        zg_cuda_refresh(ctx, h, &op, 1);
        zg_cuda_execute(ctx, h, &in, 1, &out, 1);
        zg_cuda_free(ctx, h);
        if (!close_to(y, want, 6, 1e-6f)) { printf("FAIL dense\n"); return 1; }
    }
    /* 2. input_offset = 1, input_row_stride = 4, dst_offset = 1, dst_row_stride = 4; untouched dst cells stay -7 */
    {
        float x[9] = {9, 1, 2, 3, 9, -1, 0.5f, 4, 9}, y[9];
        const float want[9] = {-7, 2.75f, 2.25f, 8.0f, -7, -3.0f, 5.25f, 6.625f, -7};
        for (int i = 0; i < 9; i++) y[i] = -7.0f;
        ZgOp op;
        memset(&op, 0, sizeof(op));
        op.tag = ZG_OP_QMATMUL;
        op.u.qmatmul.dst = 1; op.u.qmatmul.input = 0; op.u.qmatmul.weight_idx = 0;
        op.u.qmatmul.M = 2; op.u.qmatmul.N = 3; op.u.qmatmul.K = 3;
        op.u.qmatmul.input_offset = 1; op.u.qmatmul.input_row_stride = 4; op.u.qmatmul.dst_offset = 1; op.u.qmatmul.dst_row_stride = 4;
        size_t sizes[2] = {9, 9};
        ZgIO up = {1, 0, y, sizeof(y), 0};
        ZgProgram prog;
        memset(&prog, 0, sizeof(prog));
        prog.ops = &op; prog.n_ops = 1; prog.n_buffers = 2; prog.buffer_sizes = sizes; prog.initial_uploads = &up; prog.n_uploads = 1;
        prog.qweights = &qw; prog.n_qweights = 1;
        ZgCudaProgram* h = zg_cuda_compile(ctx, &prog);
        if (!h) { printf("FAIL compile (strided): %s\n", zg_cuda_last_error()); return 1; }
        float got[9];
        ZgIO in = {0, 0, x, sizeof(x), 0}, out = {1, 0, got, sizeof(got), 0};
        zg_cuda_execute(ctx, h, &in, 1, &out, 1);
        zg_cuda_free(ctx, h);
        if (!close_to(got, want, 9, 1e-6f)) { printf("FAIL strided\n"); return 1; }
    }
    /* 3. a descriptor that contradicts the op (rows != K) must make compile return NULL (src/backend.zig:284-292) */
    {
        ZgOp op;
        memset(&op, 0, sizeof(op));
        op.tag = ZG_OP_QMATMUL;
        op.u.qmatmul.dst = 1; op.u.qmatmul.input = 0; op.u.qmatmul.M = 1; op.u.qmatmul.N = 3; op.u.qmatmul.K = 4;
        size_t sizes[2] = {4, 3};
        ZgProgram prog;
        memset(&prog, 0, sizeof(prog));
        prog.ops = &op; prog.n_ops = 1; prog.n_buffers = 2; prog.buffer_sizes = sizes; prog.qweights = &qw; prog.n_qweights = 1;
        if (zg_cuda_compile(ctx, &prog) != NULL) { printf("FAIL: mismatched qweight accepted\n"); return 1; }
    }
    zg_cuda_destroy(ctx);
    printf("PASS\n");
    return 0;
}
