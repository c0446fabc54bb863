#!/usr/bin/env python
"""Row-sharded LLaMA decode tok/s (BASELINE.json config 5: Llama-3-70B-shape Q4_0, weights sharded over the GPUs
of one box; all-reduce of the d_model partial sums over NVLink peer memory, all-gather of the logits over NCCL,
both inside the decode CUDA graph).  One process per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      scripts/bench_sharded.py --model llama3-70b --kind q4_0 --tokens 32 --batch 1,8

Every rank streams its own random-init shard into HBM (host memory: one tensor at a time), decodes greedily from
`--context`, and rank 0 prints one JSON line per batch size: tok/s through step() (host patching + copies inside,
e2e) and the device-only graph replay time, max over ranks.  `--batch T` runs T-token programs (zgml's token_len)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from zgml_b200 import CudaBackend  # noqa: E402
from zgml_b200.host import llama  # noqa: E402

MODELS = {"smollm-135m": llama.SMOLLM_135M, "smollm-1.7b": llama.SMOLLM_1_7B, "llama3-8b": llama.LLAMA3_8B, "llama3-70b": llama.LLAMA3_70B}


def run_sharded_decode(be, cfg, kind, rank, world, dist, tokens=32, batches=(1,), context=0, model_name="", trace=False):
    """Load this rank's shard once, then time one session per batch size.  Returns the result dicts (every rank)."""
    import torch
    t0 = time.perf_counter()
    w, handles = llama.synthetic_resident_shard(be, cfg, kind, seed=0, rank=rank, world=world)
    t_load = time.perf_counter() - t0
    dev_bytes = sum(h.device_bytes for h in handles)
    out = []
    for T in batches:
        sess = llama.DeviceLlamaSession(be, cfg, w, T)
        pos = context
        toks = [(i + 1) % cfg.vocab_size for i in range(T)]
        lg = sess.execute_at(toks, pos)  # warm-up: captures the graph
        pos += T
        if world > 1:
            dist.barrier()
        launches0 = be.launch_count()
        t0 = time.perf_counter()
        for _ in range(tokens):
            lg = sess.execute_at(toks, pos)
            toks = toks[1:] + [int(np.argmax(lg))]
            pos += T
        dt = time.perf_counter() - t0
        launches = be.launch_count() - launches0
        be.sync()
        if world > 1:
            dist.barrier()
        t1 = time.perf_counter()
        for _ in range(tokens):
            be.lib.zg_cuda_execute_device(be.ctx, sess.handle.ptr)
        be.sync()
        dt_dev = time.perf_counter() - t1
        if world > 1:
            t = torch.tensor([dt, dt_dev], dtype=torch.float64, device="cuda" if dist.get_backend() == "nccl" else "cpu")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, dt_dev = t.tolist()
        out.append({"metric": "llama_decode_tok_s_sharded", "model": model_name, "kind": kind, "n_layers": cfg.n_layers, "n_gpus": world,
                    "batch": T, "value": round(T * tokens / dt, 1), "unit": "tok/s", "ms_per_step": round(1e3 * dt / tokens, 3),
                    "device_ms_per_step": round(1e3 * dt_dev / tokens, 3), "device_tok_s": round(T * tokens / dt_dev, 1),
                    "steps": tokens, "context": context, "ops_per_step": sess.n_ops, "kernels_per_step": launches // tokens,
                    "weight_bytes_per_gpu": dev_bytes, "hbm_gbps_per_gpu_on_weights": round(dev_bytes / (dt_dev / tokens) / 1e9, 1),
                    "hbm_floor_ms_per_step": round(dev_bytes / 6540.8e6, 3),
                    "load_s": round(t_load, 1), "comm": be.comm_mode(), "collectives_per_step": 2 * cfg.n_layers + 1 if world > 1 else 0,
                    "data": "synthetic random-init GGUF blocks, streamed per shard", "last_token": int(np.argmax(lg))})
        if trace and rank == 0:
            import collections
            import ctypes as C
            kinds = {1: "elementwise", 2: "fused_ew", 3: "rmsnorm", 4: "repeat", 5: "slice_assign", 6: "rope", 7: "attention", 8: "chain", 9: "matmul", 10: "qgemv", 11: "allreduce"}
            be.lib.zg_cuda_trace(be.ctx, 1)
        if trace:
            sess.execute_at(toks, pos)
        if trace and rank == 0:
            buf = (C.c_uint64 * (3 * 16000))()
            n = be.lib.zg_cuda_trace_read(be.ctx, buf, 16000)
            be.lib.zg_cuda_trace(be.ctx, 0)
            rec = np.frombuffer(buf, dtype=np.uint64)[:3 * n].reshape(n, 3).astype(np.int64)
            kind = rec[:, 0] >> 56
            t_in, t_go, t_out = rec[:, 0] & ((1 << 56) - 1), rec[:, 1] & ((1 << 56) - 1), rec[:, 2] & ((1 << 56) - 1)
            order = np.argsort(t_go)
            kind, t_in, t_go, t_out = kind[order], t_in[order], t_go[order], t_out[order]
            agg = collections.defaultdict(lambda: [0, 0])
            for k, b, c in zip(kind, t_go, t_out):
                if c > b:
                    agg[int(k)][0] += 1; agg[int(k)][1] += int(c - b)
            print(f"trace batch {T}: {n} records, span {(t_out.max() - t_in.min()) / 1e3:.1f} us", file=sys.stderr)
            for k, (cnt, work) in sorted(agg.items(), key=lambda x: -x[1][1]):
                print(f"  {kinds.get(k, k):12s} n={cnt:5d} work total {work / 1e3:9.1f} us avg {work / cnt / 1e3:6.2f} us", file=sys.stderr)
            mid = n // 2
            for i in range(mid, min(mid + 40, n)):
                print(f"    {kinds.get(int(kind[i]), kind[i]):12s} {(t_in[i] - t_in.min()) / 1e3:9.2f} {(t_go[i] - t_in.min()) / 1e3:9.2f} {(t_out[i] - t_in.min()) / 1e3:9.2f} {(t_out[i] - t_go[i]) / 1e3:7.2f}", file=sys.stderr)
        sess.close()
    for h in handles:
        h.free()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama3-70b", choices=sorted(MODELS))
    ap.add_argument("--kind", default="q4_0", choices=["q8_0", "q4_0"])
    ap.add_argument("--tokens", type=int, default=32)
    ap.add_argument("--batch", default="1")
    ap.add_argument("--context", type=int, default=0)
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--max-seq", type=int, default=0)
    ap.add_argument("--trace", action="store_true", help="rank 0 prints an in-graph kernel timeline of one more step to stderr")
    ap.add_argument("--emulate-world", type=int, default=0,
                    help="single process: run rank 0's shard of an N-way split with the collectives as identity (per-rank compute profile)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    cfg = MODELS[args.model]
    over = {}
    if args.layers:
        over["n_layers"] = args.layers
    if args.max_seq:
        over["max_seq_len"] = args.max_seq
    if over:
        cfg = llama.LlamaConfig(**{**cfg.__dict__, **over})
    be = CudaBackend(local)
    if world > 1:
        be.comm_init_torch()
    if args.emulate_world:
        world_shape = args.emulate_world
        orig = llama.synthetic_resident_shard
        llama.synthetic_resident_shard = lambda be_, cfg_, kind_, seed=0, rank=0, world=1, **kw: orig(be_, cfg_, kind_, seed, 0, world_shape, **kw)
    res = run_sharded_decode(be, cfg, args.kind, rank, world, dist, args.tokens, [int(b) for b in args.batch.split(",")], args.context, args.model, trace=args.trace)
    be.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        for line in res:
            print(json.dumps(line))


if __name__ == "__main__":
    main()
