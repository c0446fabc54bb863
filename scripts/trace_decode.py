#!/usr/bin/env python
"""In-graph timeline of one decode step (zg_cuda_trace): per-kernel work time, PDL wait time and the gaps between
dependent kernels.  Single GPU; `--emulate-world N` runs rank 0's shard of an N-way split (collectives = identity)."""
import argparse
import collections
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zgml_b200 import CudaBackend  # noqa: E402
from zgml_b200.host import llama  # noqa: E402

KINDS = {1: "elementwise", 2: "fused_ew", 3: "rmsnorm", 4: "repeat", 5: "slice_assign", 6: "rope", 7: "attention", 8: "chain", 9: "matmul", 10: "qgemv", 11: "allreduce",
         32: "F:norm+qkv", 33: "F:attention", 34: "F:merge+o", 35: "F:norm+gate|up", 36: "F:act+down", 37: "F:allreduce", 40: "F:item"}
MODELS = {"smollm-135m": llama.SMOLLM_135M, "smollm-1.7b": llama.SMOLLM_1_7B, "llama3-8b": llama.LLAMA3_8B, "llama3-70b": llama.LLAMA3_70B}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="smollm-1.7b")
    ap.add_argument("--kind", default="q4_0")
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--context", type=int, default=512)
    ap.add_argument("--emulate-world", type=int, default=1)
    ap.add_argument("--show", type=int, default=40)
    ap.add_argument("--batch", type=int, default=1)
    args = ap.parse_args()
    cfg = MODELS[args.model]
    if args.layers:
        cfg = llama.LlamaConfig(**{**cfg.__dict__, "n_layers": args.layers})
    be = CudaBackend(0)
    w, handles = llama.synthetic_resident_shard(be, cfg, args.kind, 0, 0, args.emulate_world)
    T = args.batch
    sess = llama.DeviceLlamaSession(be, cfg, w, T)
    toks = list(range(1, T + 1))
    for i in range(4):
        sess.execute_at(toks, args.context + i * T)
    be.lib.zg_cuda_trace(be.ctx, 1)
    sess.execute_at(toks, args.context + 4 * T)
    buf = (C.c_uint64 * (3 * 16000))()
    n = be.lib.zg_cuda_trace_read(be.ctx, buf, 16000)
    be.lib.zg_cuda_trace(be.ctx, 0)
    rec = np.frombuffer(buf, dtype=np.uint64)[:3 * n].reshape(n, 3).copy()
    kind = (rec[:, 0] >> np.uint64(56)).astype(int)
    t_in = (rec[:, 0] & np.uint64((1 << 56) - 1)).astype(np.int64)
    t_go = (rec[:, 1] & np.uint64((1 << 56) - 1)).astype(np.int64)
    t_out = (rec[:, 2] & np.uint64((1 << 56) - 1)).astype(np.int64)
    live = kind != 0
    kind, t_in, t_go, t_out = kind[live], t_in[live], t_go[live], t_out[live]
    order = np.argsort(t_go)
    kind, t_in, t_go, t_out = kind[order], t_in[order], t_go[order], t_out[order]
    t0 = t_in.min()
    pts = kind >= 64
    if pts.any():   # fused decode kernel: named points of CTA 0 (kind = 64 + 16 * phase + point)
        PH = ["qkv", "o", "gate|up", "down"]
        PT = {0: "begin", 2: "prologue", 8: "attn-begin", 9: "attn-end", 10: "matvecs", 11: "barrier", 12: "allreduce"}
        tk, tt = kind[pts], t_in[pts]
        o2 = np.argsort(tt)
        tk, tt = tk[o2], tt[o2]
        print(f"fused decode kernel: {len(tk)} trace points; showing the middle layer (us since the previous point)")
        begins = [i for i, k in enumerate(tk) if k == 64]
        if begins:
            b = begins[len(begins) // 2]
            e = begins[len(begins) // 2 + 1] if len(begins) // 2 + 1 < len(begins) else len(tk)
            for i in range(b, e):
                k = int(tk[i]) - 64
                print(f"    {PH[k // 16]:8s} {PT.get(k % 16, k % 16):10s} +{(tt[i] - tt[i - 1]) / 1e3 if i else 0:6.2f}")
            print(f"    layer total {(tt[e - 1] - tt[b]) / 1e3:.2f} us")
        keep = ~pts
        kind, t_in, t_go, t_out = kind[keep], t_in[keep], t_go[keep], t_out[keep]
        n = len(kind)
    print(f"{n} kernels, step span {(t_out.max() - t0) / 1e3:.1f} us")
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for k, a, b, c in zip(kind, t_in, t_go, t_out):
        agg[k][0] += 1; agg[k][1] += c - b; agg[k][2] += b - a
    for k, (cnt, work, wait) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"  {KINDS.get(k, k):14s} n={cnt:4d} work(after wait -> exit of block 0) total {work / 1e3:8.1f} us avg {work / cnt / 1e3:5.2f} us | resident before wait avg {wait / cnt / 1e3:5.2f} us")
    mid = n // 2
    print("  sequence sample (us from step start): kind, entry, after_wait, exit, work")
    for i in range(mid, min(mid + args.show, n)):
        print(f"    {KINDS.get(kind[i], kind[i]):14s} {(t_in[i] - t0) / 1e3:9.2f} {(t_go[i] - t0) / 1e3:9.2f} {(t_out[i] - t0) / 1e3:9.2f} {(t_out[i] - t_go[i]) / 1e3:7.2f}")
    sess.close()
    for h in handles:
        h.free()
    be.close()


if __name__ == "__main__":
    main()
