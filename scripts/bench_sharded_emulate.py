#!/usr/bin/env python
"""Device time per decode step of rank 0's shard of an 8-way sharded Llama-3-70B-shape Q4_0 model on ONE GPU (collectives are
identity at world 1): isolates the per-rank compute chain from the NVLink exchange.  Used for A/B runs of scheduling knobs."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zgml_b200 import CudaBackend
from zgml_b200.host import llama
cfg = llama.LlamaConfig(**{**llama.LLAMA3_70B.__dict__, "n_layers": int(os.environ.get("LAYERS", "16"))})
world = int(os.environ.get("EMULATE_WORLD", "8"))
be = CudaBackend(0)
w, handles = llama.synthetic_resident_shard(be, cfg, "q4_0", 0, 0, world)
sess = llama.DeviceLlamaSession(be, cfg, w, 1)
for i in range(4):
    sess.execute_at([1], 512 + i)
be.sync()
t0 = time.perf_counter()
n = int(os.environ.get("N_REPLAY", "200"))
for _ in range(n):
    be.lib.zg_cuda_execute_device(be.ctx, sess.handle.ptr)
be.sync()
dt = (time.perf_counter() - t0) / n
print(f"emulated rank 0 of {world}: {cfg.n_layers} layers, {1e6 * dt:.1f} us/step, {1e6 * dt / cfg.n_layers:.2f} us/layer (incl. head), ZG_CUDA_DECODE={os.environ.get("ZG_CUDA_DECODE", "0")}, fused layers {be.program_stats(sess.handle)["fused_decode_layers"]}, "
      f"kernels {be.program_stats(sess.handle)['kernels']}")
sess.close()
for h in handles:
    h.free()
be.close()
