#!/usr/bin/env python
"""W8A8 gemv (quantizeInput + gemvRange, reference src/quant.zig:320-459) on device: GB/s = (N*K int8 + N*K/bs f32 scales
+ 4K + 4N bytes) / CUDA-event time, rotating over more distinct weights than L2 holds.  One JSON line per shape."""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zgml_b200 import CudaBackend, QuantizedWeight

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="4096x4096,4096x14336")
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--rotation-mb", type=int, default=384)
args = ap.parse_args()
be = CudaBackend(0)
stream = torch.cuda.Stream()
be.set_stream(stream.cuda_stream)
r = np.random.default_rng(0)
for shp in args.shapes.split(","):
    K, N = (int(v) for v in shp.split("x"))
    bytes_per = N * K + N * (K // 32) * 4 + 4 * K + 4 * N
    copies = max(2, -(-args.rotation_mb * 1_000_000 // bytes_per))
    data = r.integers(-127, 128, K * N, dtype=np.int8)
    scales = r.uniform(1e-3, 1e-2, K * N // 32).astype(np.float32)
    ws = []
    for _ in range(copies):
        w = QuantizedWeight.upload(be, data, scales, K, N, 32)
        w.prepare_transposed()
        ws.append(w)
    x = torch.randn(K, device="cuda")
    y = torch.empty(N, device="cuda")
    with torch.cuda.stream(stream):
        for w in ws[:3]:
            w.gemv_device(x.data_ptr(), y.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        graph = torch.cuda.CUDAGraph()                      # launch-bound otherwise: two tiny launches per gemv from Python
        with torch.cuda.graph(graph, stream=stream):
            for i in range(args.iters):
                ws[i % copies].gemv_device(x.data_ptr(), y.data_ptr())
        graph.replay()
        torch.cuda.synchronize()
        e0.record(stream)
        graph.replay()
        e1.record(stream)
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / args.iters * 1e3
    print(json.dumps({"metric": "gemv_w8a8", "K": K, "N": N, "block_size": 32, "us": round(us, 2), "bytes": bytes_per,
                      "gbps": round(bytes_per / us / 1e3, 1), "copies": copies,
                      "note": "two launches per gemv (quantizeInput, gemvRange), replayed from one CUDA graph"}))
    for w in ws:
        w.free()
be.close()
