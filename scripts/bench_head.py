#!/usr/bin/env python
"""Device time of the dense LM-head matvec x[K] @ W[N, K]^T alone: exact f32 kernel against the argmax-safe 16-bit copies
(zg_cuda_program_promote_dense).  SmolLM shapes by default; several operand copies rotate so L2 does not serve the weights."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zgml_b200 import CudaBackend, DeviceOp, DeviceProgram, ProgramIO
be = CudaBackend(0)
for (K, N) in [(2048, 49152), (576, 49152)]:
    r = np.random.default_rng(0)
    W = (r.standard_normal((N, K)) * 0.02).astype(np.float32).ravel()
    x = r.standard_normal(K).astype(np.float32)
    copies = 3
    ops = [DeviceOp.matmul(1 + copies + i, 0, 1 + i, 1, N, K, K, 1, 1, K) for i in range(copies)]   # independent heads: one dependency level
    prog = DeviceProgram(ops, [K] + [N * K] * copies + [N] * copies, [ProgramIO(1 + i, W) for i in range(copies)], [])
    for fmt in ("f32", "f16", "bf16"):
        h = be.compile_program(prog)
        if fmt != "f32":
            assert be.promote_dense_weights(h, fmt) == copies
        out = np.zeros(N, np.float32)
        be.execute_program(h, [ProgramIO(0, x)], [ProgramIO(1 + copies, out)])
        for _ in range(3):
            be.lib.zg_cuda_execute_device(be.ctx, h.ptr)
        be.sync()
        n = 50
        t0 = time.perf_counter()
        for _ in range(n):
            be.lib.zg_cuda_execute_device(be.ctx, h.ptr)
        be.sync()
        us = (time.perf_counter() - t0) / n / copies * 1e6
        nbytes = N * K * (4 if fmt == "f32" else 2)
        print(f"head {N}x{K} {fmt}: {us:.1f} us, {nbytes / us / 1e3:.0f} GB/s, argmax {int(np.argmax(out))}, kernels {be.program_stats(h)['kernels']}")
        be.free_program(h)
be.close()
