#!/usr/bin/env python
"""Per-CTA timeline of the streamed matvec kernel (trace kind 12: every CTA records entry, first staging, exit and its SM):
is the tail of a launch a few slow SMs, a second round of CTAs, or late starters?"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zgml_b200.backend as _zb
_zb._LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libzgml_cuda_trace.so")   # built with -DZG_STREAM_TRACE_ALL (see scripts/README.md)
from zgml_b200 import CudaBackend, DeviceOp, DeviceProgram, ProgramIO, QuantizedWeight
from zgml_b200.backend import ResidentQuantizedWeight
K, N, copies = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
chain = len(sys.argv) > 4 and sys.argv[4] == "chain"
be = CudaBackend(0)
ws, ops = [], []
if chain:
    sizes = [K] + [(N if i % 2 == 0 else K) for i in range(copies)]
    for i in range(copies):
        k, n = (K, N) if i % 2 == 0 else (N, K)
        ws.append(QuantizedWeight.synth_gguf(be, 7, i, 2, k, n, 0, k, 0, n)); ops.append(DeviceOp.qmatmul(i + 1, i, i, 1, n, k))
else:
    sizes = [K, copies * N]
    for i in range(copies):
        ws.append(QuantizedWeight.synth_gguf(be, 7, i, 2, K, N, 0, K, 0, N)); ops.append(DeviceOp.qmatmul(1, 0, i, 1, N, K, dst_offset=i * N))
x = (np.random.default_rng(0).standard_normal(K) * 0.1).astype(np.float32)
h = be.compile_program(DeviceProgram(ops, sizes, [ProgramIO(0, x)], [ResidentQuantizedWeight(w) for w in ws]))
out = np.zeros(sizes[-1], np.float32)
be.execute_program(h, [], [ProgramIO(len(sizes) - 1, out)])
for _ in range(3):
    be.lib.zg_cuda_execute_device(be.ctx, h.ptr)
be.sync()
be.lib.zg_cuda_trace(be.ctx, 1)
be.lib.zg_cuda_execute_device(be.ctx, h.ptr)
be.sync()
buf = (C.c_uint64 * (3 * 16000))()
n = be.lib.zg_cuda_trace_read(be.ctx, buf, 16000)
be.lib.zg_cuda_trace(be.ctx, 0)
rec = np.frombuffer(buf, dtype=np.uint64)[:3 * n].reshape(n, 3).copy()
kind = (rec[:, 0] >> np.uint64(56)).astype(int)
M = np.uint64((1 << 56) - 1)
sel = kind == 12
M48 = np.uint64((1 << 48) - 1)
t_in, t_go, t_out = (rec[sel, 0] & M48).astype(np.int64), (rec[sel, 1] & M48).astype(np.int64), (rec[sel, 2] & M48).astype(np.int64)
cta = (rec[sel, 2] >> np.uint64(48)).astype(int)
smid = (rec[sel, 1] >> np.uint64(56)).astype(int)
print(f"{sel.sum()} CTA records")
order = np.argsort(t_in)
t_in, t_go, t_out, smid, cta = t_in[order], t_go[order], t_out[order], smid[order], cta[order]
# split into launches: gaps in entry time
t0 = t_in.min()
bounds = [0] + [i for i in range(1, len(t_in)) if t_go[i] - t_go[i - 1] > 3000] + [len(t_in)]
for a, b in zip(bounds[:-1], bounds[1:]):
    if b - a < 8: continue
    go, ex, en, sm, cb = t_go[a:b], t_out[a:b], t_in[a:b], smid[a:b], cta[a:b]
    g0 = go.min()
    dur = (ex - go) / 1e3
    print(f"launch: {b - a} CTAs on {len(set(sm))} SMs | entry spread {(en.max() - en.min()) / 1e3:.1f} us | wait-return spread {(go.max() - g0) / 1e3:.1f} | "
          f"CTA work us min {dur.min():.1f} p50 {np.median(dur):.1f} p90 {np.percentile(dur, 90):.1f} max {dur.max():.1f} | kernel {(ex.max() - g0) / 1e3:.1f} us")
    per_sm = {}
    for s_, d in zip(sm, dur): per_sm.setdefault(s_, []).append(d)
    cnt = np.bincount([len(v) for v in per_sm.values()])
    print("   CTAs per SM histogram:", {i: int(c) for i, c in enumerate(cnt) if c}, "| mean CTA work by SM-id octile:",
          [round(float(np.mean([np.mean(per_sm[s_]) for s_ in sorted(per_sm)[i::8]])), 1) for i in range(8)])
    late = np.argsort(-ex)[:6]
    print("   latest finishers (cta, sm, work us):", [(int(cb[i]), int(sm[i]), round(float(dur[i]), 1)) for i in late])
    oc = np.argsort(cb)
    print("   work us by CTA index (every 12th):", [round(float(dur[i]), 1) for i in oc[::12]])
be.free_program(h)
for w in ws: w.free()
be.close()
