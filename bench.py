#!/usr/bin/env python
"""bench.py — quantized GEMV HBM GB/s on B200 (BASELINE.json metric, configs[1]).

A "step" is one pass of the hot path over the whole microbench set: for every case
(K x N in {4096x4096, 4096x14336}) x (block format in {int8+f32 scale, Q8_0-origin,
Q4_0-origin}), one batch-1 quantized matvec over EACH of R distinct GPU-resident weight
copies (R chosen so a case's rotation set is >= 512 MB, i.e. > 4x the 126 MB L2: every
weight byte is served by HBM).  Each case is one compiled DeviceProgram (R qmatmul ops)
behind the reference's Backend interface, replayed as a CUDA graph.

  value      : algorithmic GB/s, inputs resident in HBM (zg_cuda_execute_device)
  e2e        : same steps through execute_program with HOST (pinned) buffers — activations
               uploaded and outputs downloaded inside the timed region, like
               reference src/device_inference.zig:260-263 does every token
  roofline   : dominant kernel vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline: oracle port of QuantizedWeight.matmul (src/quant.zig:475-578), 1 thread
               (the reference runs it single-threaded, src/inference_utils.zig:192)

`--impl reference` times the oracle port on all host threads instead (no GPU work).
Algorithmic bytes per GEMV = (K*N/32)*B_blk + 4K + 4N, B_blk = 36 / 34 / 18.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPES = [(4096, 4096), (4096, 14336)]          # (K, N): zgml rows=K, cols=N
FORMATS = [("i8_f32", 36), ("q8_0", 34), ("q4_0", 18)]
if os.environ.get("ZG_BENCH_SHAPES"):            # kernel-tuning sweeps only (not the judged workload)
    SHAPES = [tuple(int(v) for v in t.split("x")) for t in os.environ["ZG_BENCH_SHAPES"].split(",")]
if os.environ.get("ZG_BENCH_FORMATS"):
    FORMATS = [f for f in FORMATS if f[0] in os.environ["ZG_BENCH_FORMATS"].split(",")]
ROTATION_BYTES = 512 << 20
METRIC = "quant_gemv_hbm_gbps"
UNIT = "GB/s"


def alg_bytes(K, N, blk):
    return (K * N // 32) * blk + 4 * K + 4 * N


def host_weight(K, N, fmt, seed):
    """Synthetic weight in the reference's host form (i8 data + f32 scales, bs=32), SURVEY §8d config 2."""
    r = np.random.default_rng(seed)
    if fmt == "q4_0":
        data = r.integers(-8, 8, K * N, dtype=np.int8)
    else:
        data = r.integers(-127, 128, K * N, dtype=np.int8)
    scales = r.uniform(1e-3, 1e-2, K * N // 32).astype(np.float32)
    if fmt != "i8_f32":
        scales = scales.astype(np.float16).astype(np.float32)  # GGUF-origin scales are exact f16
    return data, scales


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML; nvidia-smi fallback)."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTE = {"sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _sample(self):
        if self.nv is not None:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            try:
                bits = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                bits = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for name, bit in {**self.BAD, **self.NOTE}.items():
                if bits & bit:
                    self.reasons.add(name)
        else:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,"
                                  "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                                  "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip().split(",")
            if len(out) >= 6:
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], out[2:6]):
                    if v.strip() == "Active":
                        self.reasons.add(name)

    def _loop(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self._stop.wait(0.02 if self.nv is not None else 0.2)

    def start(self):
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        try:
            self._sample()
        except Exception:
            pass
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


_CPU_WS = {}


def cpu_leg(threads, budget_s, passes_cap, native=True):
    """Oracle W8·f32 matmul over one GEMV per case; returns (GB/s, seconds per pass, passes, per-case GB/s)."""
    from oracle import oracle
    ws = _CPU_WS
    if not ws:
        for (K, N) in SHAPES:
            for fmt, blk in FORMATS:
                d, s = host_weight(K, N, fmt, 1)
                ws[(K, N, fmt)] = (oracle.QuantizedWeight(d, s, K, N, 32), blk)
        for (K, N, fmt), (qw, blk) in ws.items():  # warm-up pass (llama_smollm_bench.zig:147 does one too)
            qw.matmul(np.zeros(K, np.float32), 1, threads=threads, native=native)
    x = np.random.default_rng(7).standard_normal(4096).astype(np.float32)
    per_case = {k: 0.0 for k in ws}
    passes, t0 = 0, time.perf_counter()
    while passes < passes_cap and (time.perf_counter() - t0 < budget_s or passes < 1):
        for key, (qw, blk) in ws.items():
            t = time.perf_counter()
            qw.matmul(x, 1, threads=threads, native=native)
            per_case[key] += time.perf_counter() - t
        passes += 1
    total = sum(per_case.values())
    total_bytes = sum(alg_bytes(K, N, blk) for (K, N, fmt), (qw, blk) in ws.items()) * passes
    cases = [{"K": K, "N": N, "format": fmt, "gbps": alg_bytes(K, N, ws[(K, N, fmt)][1]) * passes / per_case[(K, N, fmt)] / 1e9}
             for (K, N, fmt) in ws]
    return total_bytes / total / 1e9, total / passes, passes, cases


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def config_dict(n_gpus):
    return {"workload": "quantized matvec microbench 4096x4096 and 4096x14336, each quant.zig block format "
                        "(int8+f32 scale 36 B, Q8_0-origin 34 B, Q4_0-origin 18 B per 32 weights), batch 1",
            "shapes_KxN": ["4096x4096", "4096x14336"], "formats": [f for f, _ in FORMATS], "batch": 1,
            "l2": "inputs larger than L2: >= 512 MB of distinct weight copies per case, visited round-robin",
            "parallelism": "1 GPU" if n_gpus == 1 else f"{n_gpus} ranks, each with its own weight shards (no data-path collective)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    thr = host_threads()
    # each step = one GEMV per case (a bounded sample of the GPU arm's step, which visits R copies per case)
    from oracle import oracle  # noqa: F401  (build once, outside the timed steps)
    t_all = []
    gb, sec, passes, cases = cpu_leg(thr, 0.0, 1)  # warm-up + build
    for _ in range(max(args.warmup - 1, 0)):
        cpu_leg(thr, 0.0, 1)
    total_bytes = sum(alg_bytes(K, N, blk) for (K, N) in SHAPES for _, blk in FORMATS)
    for _ in range(args.steps):
        gb, sec, passes, cases = cpu_leg(thr, 0.0, 1)
        t_all.append(sec)
    ms = 1e3 * sum(t_all) / len(t_all)
    value = total_bytes / (ms / 1e3) / 1e9
    sample = "one batch-1 matvec per case (6 GEMVs, 3 of each shape) per step; oracle port of src/quant.zig:475-578, N split over threads"
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32 (i8 weights x f32 scales x f32 activations)",
                      "data": "synthetic", "config": config_dict(args.gpus),
                      "cpu_baseline": {"value": value, "unit": UNIT, "cores": thr, "kind": "port", "sample": sample},
                      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


def run_b200(args):
    import torch
    import torch.distributed as dist
    from zgml_b200 import CudaBackend, DeviceOp, DeviceProgram, ProgramIO, QuantizedWeightUpload
    from zgml_b200.backend import _io_array

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 backend has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    be = CudaBackend(local)
    stream = torch.cuda.Stream()
    be.set_stream(stream.cuda_stream)

    rot = args.rotation_mb << 20
    cases = []
    for (K, N) in SHAPES:
        for fmt, blk in FORMATS:
            R = max(2, -(-rot // alg_bytes(K, N, blk)))
            data, scales = host_weight(K, N, fmt, seed=1 + rank)
            qws = [QuantizedWeightUpload(data, scales, K, N, 32) for _ in range(R)]  # R distinct device copies
            ops = [DeviceOp.qmatmul(1, 0, i, 1, N, K, dst_offset=i * N) for i in range(R)]
            prog = DeviceProgram(ops, [K, R * N], [], qws)
            h = be.compile_program(prog)
            if h is None:
                from zgml_b200.backend import last_error
                raise SystemExit(f"compile_program failed: {last_error()}")
            x = torch.randn(K, dtype=torch.float32).pin_memory()
            out = torch.empty(R * N, dtype=torch.float32).pin_memory()
            cases.append({"K": K, "N": N, "format": fmt, "blk": blk, "R": R, "h": h, "prog": prog, "ops": ops,
                          "x": x, "out": out, "xin": _io_array([ProgramIO(0, x.numpy())]), "oout": _io_array([ProgramIO(1, out.numpy())]),
                          "bytes": alg_bytes(K, N, blk) * R})
            be.lib.zg_cuda_execute(be.ctx, h.ptr, cases[-1]["xin"], 1, cases[-1]["oout"], 1)  # uploads x; captures the graph
    step_bytes = sum(c["bytes"] for c in cases)
    lib, ctx = be.lib, be.ctx

    def step_device():
        for c in cases:
            lib.zg_cuda_execute_device(ctx, c["h"].ptr)

    def step_e2e():
        for c in cases:  # VTable.execute_program: upload inputs -> run ops -> download outputs, synchronous
            lib.zg_cuda_execute(ctx, c["h"].ptr, c["xin"], 1, c["oout"], 1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step_device()
        barrier()
        n_ev = len(cases) + 1
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(n_ev)] for _ in range(args.steps)]
        launches0 = be.launch_count()
        sampler.start()
        barrier()
        for s in range(args.steps):
            evs[s][0].record(stream)
            for i, c in enumerate(cases):
                lib.zg_cuda_execute_device(ctx, c["h"].ptr)
                evs[s][i + 1].record(stream)
        barrier()
        launches = be.launch_count() - launches0
        total_ms = evs[0][0].elapsed_time(evs[-1][-1])
        case_ms = [sum(evs[s][i].elapsed_time(evs[s][i + 1]) for s in range(args.steps)) for i in range(len(cases))]

        # e2e: host buffers, copies inside the timed region
        for _ in range(max(args.warmup, 3)):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()

    if world > 1:
        t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = t.tolist()
    ms_per_step = total_ms / args.steps
    value = world * step_bytes / (ms_per_step / 1e3) / 1e9
    e2e_value = world * step_bytes / (e2e_ms / args.steps / 1e3) / 1e9
    peak, peak_kind = peaks()

    case_out = []
    for c, ms in zip(cases, case_ms):
        us = ms * 1e3 / (args.steps * c["R"])
        gb = alg_bytes(c["K"], c["N"], c["blk"]) / (us * 1e-6) / 1e9
        case_out.append({"K": c["K"], "N": c["N"], "format": c["format"], "copies": c["R"], "us_per_gemv": round(us, 3),
                         "gbps": round(gb, 1), "frac_of_measured_hbm": round(gb / peak, 4), "frac_of_8TBps": round(gb / 8000.0, 4),
                         "share_of_step": round(ms / sum(case_ms), 4)})
    dom_i = max(range(len(cases)), key=lambda i: case_ms[i])
    dom, domc = case_out[dom_i], cases[dom_i]
    # one launch carries up to 8 same-level same-shape matvecs (csrc/backend.cu gemv_batch): per-launch figures below
    per_launch = min(8, domc["R"])
    traffic, traffic_note = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            t = json.load(open(tp)).get(f"{domc['K']}x{domc['N']}_{domc['format']}")
            if isinstance(t, dict):
                traffic = t["dram_bytes_per_launch"] if t.get("gemvs_per_launch") == per_launch else int(t["dram_bytes_per_launch"] / max(t.get("gemvs_per_launch", 1), 1) * per_launch)
                traffic_note = f"ncu dram__bytes_read.sum + dram__bytes_write.sum per launch of {t.get('gemvs_per_launch')} matvecs (profiles/traffic.json)"
            elif t is not None:
                traffic = int(t) * per_launch
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": f"qgemv_kernel<{domc['format']}> {domc['K']}x{domc['N']} batch 1, {per_launch} matvecs per launch",
                "achieved": dom["gbps"], "peak": peak, "peak_kind": f"{peak_kind} copy bandwidth (MEASURED_PEAKS.json hbm_gbs)",
                "unit": "GB/s", "frac": round(dom["gbps"] / peak, 4), "frac_all_cases": round(value / world / peak, 4),
                "algorithmic_bytes_per_launch": alg_bytes(domc["K"], domc["N"], domc["blk"]) * per_launch,
                "launch_us": round(dom["us_per_gemv"] * per_launch, 3), "matvecs_per_launch": per_launch,
                "traffic": traffic, "traffic_note": traffic_note}

    line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (i8/i4 weights x f16/f32 block scales x f32 activations, fp32 accumulate)",
            "data": "synthetic", "config": config_dict(world), "clocks": clocks,
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "ms_per_step": round(e2e_ms / args.steps, 4),
                    "h2d_bytes_per_step": sum(4 * c["K"] for c in cases), "d2h_bytes_per_step": sum(4 * c["R"] * c["N"] for c in cases)},
            "gpu_launches": int(launches), "roofline": roofline, "cases": case_out}

    if rank == 0 and world == 1 and not args.no_extras:
        with torch.cuda.stream(stream):
            line["extras"] = run_extras(be, torch)
    if world > 1 and not args.no_extras:
        # BASELINE.json config 5 beside the weak-scaling GEMV number: Llama-3-70B-shape Q4_0 decode with every linear
        # row-sharded over the N GPUs (collectives on the data path), batch 1 and 8.  Never part of `value`.
        try:
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            import bench_sharded
            from zgml_b200.host import llama
            be.comm_init_torch()
            with torch.cuda.stream(stream):
                res = bench_sharded.run_sharded_decode(be, llama.LLAMA3_70B, "q4_0", rank, world, dist, tokens=16, batches=(1, 8),
                                                       context=512, model_name="llama3-70b")
            line["extras"] = {"llama3_70b_q4_0_decode_sharded": res}
        except Exception as e:  # extras never break the contract line
            line["extras"] = {"sharded_decode_error": repr(e)}
    if rank == 0 and world == 1 and not args.no_cpu:
        gb, sec, passes, ccases = cpu_leg(1, args.cpu_seconds, 200)
        line["cpu_baseline"] = {"value": round(gb, 3), "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"{passes} passes of one batch-1 matvec per case (6 GEMVs/pass, {sec * 1e3:.1f} ms/pass); "
                                          "oracle port of QuantizedWeight.matmul, single thread like src/inference_utils.zig:192",
                                "host_threads_available": host_threads(), "cases": [{**c, "gbps": round(c["gbps"], 3)} for c in ccases]}
    for c in cases:
        be.free_program(c["h"])
    be.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


def run_extras(be, torch):
    """Bounded side measurements of the other BASELINE.json configs (never part of `value`): the tcgen05 prefill
    GEMM at Llama-3-8B shapes (config 4) and greedy decode of a SmolLM-135M-shape Q8_0 program (config 1)."""
    out = {}
    try:
        from zgml_b200 import QuantizedWeight
        r = np.random.default_rng(3)
        M, gem = 2048, []
        for (K, N) in [(4096, 4096), (4096, 14336)]:
            data = r.integers(-127, 128, K * N, dtype=np.int8)
            scales = r.uniform(1e-3, 1e-2, K * N // 32).astype(np.float16).astype(np.float32)
            ws = [QuantizedWeight.upload(be, data, scales, K, N, 32) for _ in range(2)]
            x = torch.randn(M, K, device="cuda")
            y = torch.empty(M, N, device="cuda")
            for w in ws:
                w.matmul_device(x.data_ptr(), y.data_ptr(), M)
            be.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            st = torch.cuda.current_stream()
            e0.record(st)
            for i in range(10):
                ws[i % 2].matmul_device(x.data_ptr(), y.data_ptr(), M)
            e1.record(st)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            gem.append({"M": M, "K": K, "N": N, "format": "q8_0", "ms": round(ms, 4), "tflops": round(2.0 * M * N * K / ms / 1e9, 1)})
            for w in ws:
                w.free()
        issued = round(3 * max(g["tflops"] for g in gem), 1)
        bf16_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops", 0) or 0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 0.0
        out["prefill_qgemm_tcgen05"] = {"mode": "3xBF16 (hi/lo split of both operands, fp32 accumulate in TMEM)", "cases": gem,
                                        "bf16_mma_tflops_issued": issued,
                                        "frac_of_measured_bf16_peak": round(issued / bf16_peak, 3) if bf16_peak else None}
    except Exception as e:  # extras never break the contract line
        out["prefill_error"] = repr(e)
    try:
        from zgml_b200.host import llama
        cfg = llama.SMOLLM_135M
        sess = llama.DeviceLlamaSession(be, cfg, llama.synthetic_weights(cfg, "q8_0", seed=0), 1)
        tok = int(np.argmax(sess.step(1)))
        t0 = time.perf_counter()
        n = 64
        for _ in range(n):
            tok = int(np.argmax(sess.step(tok)))
        dt = time.perf_counter() - t0
        out["decode_smollm_135m_q8_0"] = {"tok_s": round(n / dt, 1), "ms_per_token": round(1e3 * dt / n, 3), "ops_per_token": sess.n_ops,
                                          "path": "DeviceProgram through execute_program (host copies inside), greedy"}
        sess.close()
    except Exception as e:
        out["decode_error"] = repr(e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rotation-mb", type=int, default=ROTATION_BYTES >> 20)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
