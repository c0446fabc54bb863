#!/usr/bin/env python
"""Quantized KV cache inside programs (zg_cuda_program_quantize_kv): attention GB/s against the algorithmic bytes of
SURVEY.md §8f-2 — 2 * seq_kv * (d_head + 4 * d_head / bs) per head and decode step — next to the f32 cache, and SmolLM-1.7B
Q4_0 decode tok/s with the mode on / off at 512 and 2040 tokens of context."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zgml_b200 import CudaBackend, DeviceOp, DeviceProgram, ProgramIO  # noqa: E402
from zgml_b200.host import llama  # noqa: E402


def attention_only(be, H, dh, S, ctx, quant, layers=8):
    """`layers` independent layers' worth of per-head {K store, V store, attention} over their own caches (> L2 in total at S = 2048)."""
    ops, sizes = [], [H * dh, S]                  # 0: q / new k / new v source, 1: mask column; one output buffer per head
    for L in range(layers):
        kc, vc = len(sizes), len(sizes) + 1
        sizes += [H * S * dh, H * S * dh]
        for h in range(H):
            base = h * S * dh
            ops.append(DeviceOp.slice_assign(kc, 0, dh, 1, base, base + ctx * dh, 1, dh, h * dh, 1, dh, dh))
            ops.append(DeviceOp.slice_assign(vc, 0, dh, 1, base, base + ctx * dh, 1, dh, h * dh, 1, dh, dh))
        for h in range(H):
            base = h * S * dh
            sizes.append(dh)
            ops.append(DeviceOp.attention(len(sizes) - 1, 0, kc, vc, 1, True, dh, 1, ctx + 1, float(1 / np.sqrt(dh)), h * dh, base, base, 0, 0,
                                          1, dh, 1, dh, 1, dh, 1, S, 1, dh))
    prog = DeviceProgram(ops, sizes, [], [])
    h = be.compile_program(prog)
    assert h is not None
    if quant:
        be.quantize_kv(h, 32, False)
    x = np.random.default_rng(0).standard_normal(H * dh).astype(np.float32)
    out = np.zeros(dh, np.float32)
    be.execute_program(h, [ProgramIO(0, x), ProgramIO(1, np.zeros(S, np.float32))], [ProgramIO(len(sizes) - 1, out)])
    be.sync()
    n = 100
    t0 = time.perf_counter()
    for _ in range(n):
        be.lib.zg_cuda_execute_device(be.ctx, h.ptr)
    be.sync()
    dt = (time.perf_counter() - t0) / n
    per_elem = (1 + 4 / 32) if quant else 4
    nbytes = layers * H * 2 * (ctx + 1) * dh * per_elem
    be.free_program(h)
    return {"heads": H, "d_head": dh, "context": ctx, "layers": layers, "cache": "q8 (bs 32)" if quant else "f32", "us_per_layer": round(1e6 * dt / layers, 2),
            "algorithmic_bytes_per_layer": int(nbytes / layers), "gbps": round(nbytes / dt / 1e9, 1)}


def decode(be, ctx, quant):
    cfg = llama.SMOLLM_1_7B
    w, handles = llama.synthetic_resident_shard(be, cfg, "q4_0", seed=0)
    sess = llama.DeviceLlamaSession(be, cfg, w, 1)
    if quant:
        be.quantize_kv(sess.handle, 32, False)
    sess.pos = ctx
    tok = int(np.argmax(sess.step(1)))
    be.sync()
    n = 64
    t0 = time.perf_counter()
    for _ in range(n):
        be.lib.zg_cuda_execute_device(be.ctx, sess.handle.ptr)
    be.sync()
    dt = (time.perf_counter() - t0) / n
    sess.close()
    for h in handles:
        h.free()
    return {"model": "smollm-1.7b q4_0", "context": ctx, "cache": "q8 (bs 32)" if quant else "f32", "device_tok_s": round(1 / dt, 1), "ms_per_token": round(1e3 * dt, 3)}


def main():
    be = CudaBackend(0)
    res = []
    for ctx in (512, 2040):
        for quant in (False, True):
            res.append(attention_only(be, 32, 64, 2048, ctx, quant))
    for ctx in (512, 2040):
        for quant in (False, True):
            res.append(decode(be, ctx, quant))
    be.close()
    for r in res:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
