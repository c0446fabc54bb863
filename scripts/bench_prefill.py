#!/usr/bin/env python
"""Prefill GEMM throughput on the tcgen05 path: M x K x N quantized matmul, CUDA-event timed, weights rotating
over > L2 worth of copies.  Prints one JSON line per case (TFLOP/s = 2 M N K / t; 3xBF16 issues 3x that in MMAs)."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zgml_b200 import CudaBackend, QuantizedWeight

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=2048)
ap.add_argument("--shapes", default="4096x4096,4096x14336,14336x4096")
ap.add_argument("--kind", default="q8_0")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--check", action="store_true", help="compare 64 sampled rows with a torch fp64 product of the dequantized weight")
args = ap.parse_args()
be = CudaBackend(0)
stream = torch.cuda.Stream()
be.set_stream(stream.cuda_stream)
r = np.random.default_rng(0)
for shp in args.shapes.split(","):
    K, N = (int(v) for v in shp.split("x"))
    if args.kind == "q4_0":
        data = r.integers(-8, 8, K * N, dtype=np.int8)
    else:
        data = r.integers(-127, 128, K * N, dtype=np.int8)
    scales = r.uniform(1e-3, 1e-2, K * N // 32).astype(np.float16).astype(np.float32)
    ws = [QuantizedWeight.upload(be, data, scales, K, N, 32) for _ in range(3)]
    x = torch.randn(args.M, K, device="cuda")
    y = torch.empty(args.M, N, device="cuda")
    with torch.cuda.stream(stream):
        for w in ws:
            w.matmul_device(x.data_ptr(), y.data_ptr(), args.M)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(args.iters):
            ws[i % 3].matmul_device(x.data_ptr(), y.data_ptr(), args.M)
        e1.record(stream)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    err = None
    if args.check:   # floating-point kernel: torch reference of the same product, fp64 accumulate (tolerance: north star's 1e-3)
        ws[(args.iters - 1) % 3].matmul_device(x.data_ptr(), y.data_ptr(), args.M)
        torch.cuda.synchronize()
        wd = (torch.from_numpy(data.astype(np.float32)).view(K * N // 32, 32) * torch.from_numpy(scales)[:, None]).view(K, N).cuda()
        rows = torch.linspace(0, args.M - 1, 64).long().cuda()
        ref = x[rows].double() @ wd.double()
        err = float((y[rows].double() - ref).abs().max() / ref.abs().max())
        del wd, ref
    flops = 2.0 * args.M * N * K
    print(json.dumps({"metric": "prefill_qgemm", "M": args.M, "K": K, "N": N, "kind": args.kind, "ms": round(ms, 4),
                      "tflops": round(flops / ms / 1e9, 1), "tok_per_s_this_linear": round(args.M / ms * 1e3),
                      "mode": "bf16x1" if os.environ.get("ZG_GEMM_X1") == "1" else "3xBF16",
                      "tile": "cta_pair 256x256" if os.environ.get("ZG_GEMM_CTA2") == "1" else "128x256" if os.environ.get("ZG_GEMM_MH", "1") != "2" else "256x128",
                      "max_rel_err": err}))
    for w in ws:
        w.free()
be.close()
