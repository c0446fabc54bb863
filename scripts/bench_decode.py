#!/usr/bin/env python
"""Decode tok/s of a synthetic GGUF-direct LLaMA through the Backend interface (BASELINE.json configs 1 and 3).

Mirrors benchmarks/llama_smollm_bench.zig:147-177 (`runDeviceVariant`): 1 warm-up step, then `--tokens`
greedy steps from `--context` already-filled positions, wall clock around step() (host patching + refresh +
execute with host<->device copies = the e2e number); prints one JSON line.  `--cpu-tokens N` also times the
oracle executor (reference semantics on the host cores) on the same program for N tokens."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

from zgml_b200 import CudaBackend  # noqa: E402
from zgml_b200.host import llama  # noqa: E402

MODELS = {"smollm-135m": llama.SMOLLM_135M, "smollm-1.7b": llama.SMOLLM_1_7B, "llama3-8b": llama.LLAMA3_8B, "llama3-70b": llama.LLAMA3_70B}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="smollm-135m", choices=sorted(MODELS))
    ap.add_argument("--kind", default="q8_0", choices=["q8_0", "q4_0"])
    ap.add_argument("--tokens", type=int, default=64)
    ap.add_argument("--context", type=int, default=0, help="decode starts at this position (KV cache below it holds zeros)")
    ap.add_argument("--cpu-tokens", type=int, default=0)
    ap.add_argument("--profile", action="store_true", help="per-DeviceOp-tag device time (one launch per op, program order)")
    ap.add_argument("--head", default="f32", choices=["f32", "f16", "bf16"], help="tied LM head operand format (promote_dense_weights: argmax-safe 16-bit copy)")
    ap.add_argument("--layers", type=int, default=0, help="override n_layers (memory-bounded experiments only)")
    ap.add_argument("--gguf", default="", help="decode a GGUF file (Q8_0 / Q4_0 linears) instead of in-memory synthetic weights: memory-mapped, "
                    "raw blocks go tensor by tensor to HBM (host/gguf.py::load_resident)")
    ap.add_argument("--write-gguf", default="", help="first write the synthetic model of --model / --kind to this path, then decode it from the file")
    args = ap.parse_args()
    cfg = MODELS[args.model]
    if args.layers:
        cfg = llama.LlamaConfig(**{**cfg.__dict__, "n_layers": args.layers})
    be = CudaBackend(0)
    handles = []
    t0 = time.perf_counter()
    if args.gguf or args.write_gguf:
        from zgml_b200.host import gguf
        if args.write_gguf:
            gguf.write_llama_gguf(args.write_gguf, cfg, args.kind, seed=0)
        gf = gguf.GGUFFile.open(args.gguf or args.write_gguf)
        cfg = gguf.config_from_gguf(gf)
        w, handles = gguf.load_resident(be, gf, cfg)
        t = gf.get_tensor_info("blk.0.attn_q.weight").type_
        args.kind = "q8_0" if t == gguf.GGMLType.q8_0 else "q4_0"
    else:
        w = llama.synthetic_weights(cfg, args.kind, seed=0)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    sess = llama.DeviceLlamaSession(be, cfg, w, 1)
    t_compile = time.perf_counter() - t0
    promoted = be.promote_dense_weights(sess.handle, args.head) if args.head != "f32" else 0
    sess.pos = args.context
    tok = 1
    lg = sess.step(tok)  # warm-up (captures the graph)
    tok = int(np.argmax(lg))
    toks = []
    launches0 = be.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.tokens):
        lg = sess.step(tok)
        tok = int(np.argmax(lg))
        toks.append(tok)
    dt = time.perf_counter() - t0
    launches = be.launch_count() - launches0
    # device-only: replay the step's graph with inputs resident (no host patching, no copies)
    be.sync()
    t1 = time.perf_counter()
    for _ in range(args.tokens):
        be.lib.zg_cuda_execute_device(be.ctx, sess.handle.ptr)
    be.sync()
    dt_dev = time.perf_counter() - t1
    qbytes = sum(k * n for k, n in llama.linear_shapes(cfg).values()) * cfg.n_layers // 32 * (34 if args.kind == "q8_0" else 18)
    head_bytes = cfg.vocab_size * cfg.d_model * 4 if cfg.tied_lm_head else cfg.vocab_size * cfg.d_model // 32 * (34 if args.kind == "q8_0" else 18)
    line = {"metric": "llama_decode_tok_s", "model": args.model, "kind": args.kind, "n_layers": cfg.n_layers, "value": round(args.tokens / dt, 1),
            "unit": "tok/s", "ms_per_token": round(1e3 * dt / args.tokens, 3), "device_ms_per_token": round(1e3 * dt_dev / args.tokens, 3),
            "device_tok_s": round(args.tokens / dt_dev, 1), "tokens": args.tokens, "context": args.context,
            "ops_per_token": sess.n_ops, "kernels_per_token": launches // args.tokens, "schedule": be.program_stats(sess.handle),
            "weight_bytes_per_token": qbytes + head_bytes, "hbm_gbps_on_weights": round((qbytes + head_bytes) / (dt / args.tokens) / 1e9, 1),
            "lm_head": args.head, "dense_ops_promoted": promoted, "head_bytes_saved_per_token": be.program_stats(sess.handle)["dense_bytes_saved_per_execution"],
            "gen_s": round(t_gen, 1), "compile_s": round(t_compile, 1), "data": "synthetic random-init GGUF-direct weights"}
    if args.profile:
        be.set_profiling(True)
        for _ in range(4):
            sess.execute_at([tok], sess.pos - 1)
        prof = be.get_runtime_profile(sess.handle)
        names = ["elementwise", "matmul", "qmatmul", "softmax", "layernorm", "rmsnorm", "reduce", "repeat", "slice_assign", "rope", "attention", "fused_elementwise"]
        line["profile_us_per_token_by_tag"] = {n: round(prof.time_ns[i] / 1e3 / prof.call_count, 1) for i, n in enumerate(names) if prof.time_ns[i]}
        be.set_profiling(False)
    if args.cpu_tokens:
        from llama_reference import OracleBackend
        w_cpu = w
        if handles:   # resident descriptors only exist on the device: the CPU arm reads the same file in the reference's host form
            from zgml_b200.host import gguf
            w_cpu = gguf.load_direct_quantized(gguf.GGUFFile.open(args.gguf or args.write_gguf), cfg)
        ref = llama.DeviceLlamaSession(OracleBackend(native=True), cfg, w_cpu, 1)
        ref.pos = args.context
        t = 1
        lg = ref.step(t)
        t = int(np.argmax(lg))
        ctoks = []
        t0 = time.perf_counter()
        for _ in range(args.cpu_tokens):
            lg = ref.step(t)
            t = int(np.argmax(lg))
            ctoks.append(t)
        cdt = time.perf_counter() - t0
        line["cpu_reference_tok_s"] = round(args.cpu_tokens / cdt, 2)
        line["cpu_kind"] = "oracle executor (reference.zig semantics), 1 thread"
        line["greedy_tokens_match_cpu"] = ctoks == toks[:args.cpu_tokens]
        ref.close()
    sess.close()
    for h in handles:
        h.free()
    be.close()
    if args.gguf or args.write_gguf:
        line["data"] = f"GGUF file {args.gguf or args.write_gguf} (memory-mapped, raw blocks -> HBM)"
    print(json.dumps(line))


if __name__ == "__main__":
    main()
