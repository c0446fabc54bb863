#!/usr/bin/env python
"""One quantized-matvec shape through a compiled DeviceProgram: `--copies` distinct weights, either independent (batched 8 per
launch by the scheduler) or `--chain` (op i reads op i-1's output; K -> N -> K -> ... needs --copies even unless K == N).
Prints device GB/s on the algorithmic bytes.  Used for A/B runs of the matvec kernels and as the ncu target for one shape."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zgml_b200 import CudaBackend, DeviceOp, DeviceProgram, ProgramIO, QuantizedWeight
from zgml_b200.backend import ResidentQuantizedWeight

ap = argparse.ArgumentParser()
ap.add_argument("K", type=int); ap.add_argument("N", type=int)
ap.add_argument("--kind", default="q4_0"); ap.add_argument("--copies", type=int, default=16)
ap.add_argument("--chain", action="store_true"); ap.add_argument("--reps", type=int, default=50)
a = ap.parse_args()
be = CudaBackend(0)
gt = 2 if a.kind == "q4_0" else 8
blk = 18 if a.kind == "q4_0" else 34
ws, ops, sizes = [], [], []
if a.chain:
    sizes = [a.K] + [(a.N if i % 2 == 0 else a.K) for i in range(a.copies)]
    for i in range(a.copies):
        k, n = (a.K, a.N) if i % 2 == 0 else (a.N, a.K)
        ws.append(QuantizedWeight.synth_gguf(be, 7, i, gt, k, n, 0, k, 0, n))
        ops.append(DeviceOp.qmatmul(i + 1, i, i, 1, n, k))
else:
    sizes = [a.K, a.copies * a.N]
    for i in range(a.copies):
        ws.append(QuantizedWeight.synth_gguf(be, 7, i, gt, a.K, a.N, 0, a.K, 0, a.N))
        ops.append(DeviceOp.qmatmul(1, 0, i, 1, a.N, a.K, dst_offset=i * a.N))
x = (np.random.default_rng(0).standard_normal(a.K) * 0.1).astype(np.float32)
prog = DeviceProgram(ops, sizes, [ProgramIO(0, x)], [ResidentQuantizedWeight(w) for w in ws])
h = be.compile_program(prog)
assert h is not None
out = np.zeros(sizes[-1], np.float32)
be.execute_program(h, [], [ProgramIO(len(sizes) - 1, out)])
for _ in range(3):
    be.lib.zg_cuda_execute_device(be.ctx, h.ptr)
be.sync()
t0 = time.perf_counter()
for _ in range(a.reps):
    be.lib.zg_cuda_execute_device(be.ctx, h.ptr)
be.sync()
dt = (time.perf_counter() - t0) / a.reps
nbytes = a.copies * (a.K * a.N // 32) * blk
print(f"{a.kind} {a.K}x{a.N} copies {a.copies} {'chain' if a.chain else 'batch'}: {1e6 * dt / a.copies:.2f} us/matvec, {nbytes / dt / 1e9:.0f} GB/s, kernels {be.program_stats(h)['kernels']}, finite {bool(np.isfinite(out).all())}")
be.free_program(h)
for w in ws:
    w.free()
be.close()
