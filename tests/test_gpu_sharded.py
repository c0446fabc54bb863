"""Row-sharded decode on real GPUs and the resident-weight extension.

* 1 GPU: a program built from weights streamed into HBM first (ZG_QWEIGHT_RESIDENT descriptors) equals the program
  that uploads the same weights through compile_program; collective ops at world 1 are identity / copy.
* >= 2 GPUs (skipped otherwise; `gpurun --gpus 2`): two processes, one per GPU, NCCL all-reduce / all-gather inside
  the program's CUDA graph; logits vs the unsharded oracle within 1e-3 relative, greedy tokens identical."""
import os
import socket

import numpy as np
import pytest

from llama_reference import OracleBackend
from zgml_b200.host.llama import (DeviceLlamaSession, LlamaConfig, LlamaWeights, shard_weights, synthetic_gguf_blocks,
                                  synthetic_resident_shard, synthetic_weights)

pytestmark = pytest.mark.gpu

CFG = LlamaConfig(vocab_size=256, d_model=128, n_layers=2, n_heads=4, n_kv_heads=2, d_ff=192, max_seq_len=32,
                  rope_base=5e5, tied_lm_head=False)
CFG_TIED = LlamaConfig(vocab_size=256, d_model=128, n_layers=2, n_heads=4, n_kv_heads=2, d_ff=192, max_seq_len=32)


def rel(got, want):
    return float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64))) / (np.max(np.abs(want)) + 1e-30))


def greedy(sess, first, n):
    toks, logs, t = [], [], first
    for _ in range(n):
        lg = sess.step(t).copy()
        t = int(np.argmax(lg))
        toks.append(t)
        logs.append(lg)
    return toks, np.stack(logs)


def test_resident_weights_program_equals_uploaded_program(cuda_backend):
    from zgml_b200 import QuantizedWeight
    from zgml_b200.backend import ResidentQuantizedWeight
    from oracle import oracle
    w = synthetic_weights(CFG, "q4_0", seed=4, embed_scale=1.0)
    handles = []

    def resident(up):
        h = QuantizedWeight.upload(cuda_backend, up.data, up.scales, up.rows, up.cols, up.block_size)
        handles.append(h)
        return ResidentQuantizedWeight(h)

    wr = LlamaWeights(CFG, w.token_embed, [{n: resident(q) for n, q in L.items()} for L in w.layers], w.norm1, w.norm2,
                      w.norm_f, resident(w.out_proj))
    a, b = DeviceLlamaSession(cuda_backend, CFG, w), DeviceLlamaSession(cuda_backend, CFG, wr)
    ta, la = greedy(a, 1, 4)
    tb, lb = greedy(b, 1, 4)
    a.close(); b.close()
    for h in handles:
        h.free()
    assert ta == tb and np.array_equal(la, lb)
    # streamed GGUF blocks: same bytes through the oracle's importer give the same greedy decode
    r = np.random.default_rng(5)
    raw = synthetic_gguf_blocks(r, 64, 96, "q8_0")
    h = QuantizedWeight.from_gguf_blocks(cuda_backend, raw, 8, 64, 96)
    assert np.array_equal(h.dequantize_to(), oracle.QuantizedWeight.from_gguf(raw, 8, 64, 96).dequantize_to())
    h.free()


def test_synthetic_resident_shard_world1_decodes(cuda_backend):
    w, handles = synthetic_resident_shard(cuda_backend, CFG_TIED, "q8_0", seed=2, embed_scale=1.0)
    s = DeviceLlamaSession(cuda_backend, CFG_TIED, w)
    toks, logs = greedy(s, 1, 3)
    s.close()
    for h in handles:
        h.free()
    assert np.isfinite(logs).all() and np.ptp(logs) > 0


def test_collective_ops_at_world_1_are_identity(cuda_backend):
    from zgml_b200 import DeviceOp, DeviceProgram, ProgramIO
    x = np.arange(8, dtype=np.float32)
    prog = DeviceProgram([DeviceOp.allreduce(0, 8), DeviceOp.allgather(1, 0, 4, dst_offset=2, src_offset=3)], [8, 8], [], [])
    h = cuda_backend.compile_program(prog)
    assert h is not None
    a, b = np.zeros(8, np.float32), np.zeros(8, np.float32)
    cuda_backend.execute_program(h, [ProgramIO(0, x)], [ProgramIO(0, a), ProgramIO(1, b)])
    cuda_backend.free_program(h)
    assert np.array_equal(a, x) and np.array_equal(b[2:6], x[3:7]) and b[:2].sum() == 0 and b[6:].sum() == 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spawn(fn, world, args_after_port):
    """mp.spawn with a fresh rendezvous port; a port that another process grabbed between the probe and rank 0's listen
    (EADDRINUSE, seen once on a busy 2-GPU box) is retried with a new one."""
    import torch.multiprocessing as mp
    for attempt in range(3):
        try:
            mp.spawn(fn, args=(world, _free_port()) + tuple(args_after_port), nprocs=world, join=True)
            return
        except Exception as e:  # ProcessRaisedException carries the child's traceback as text
            if "EADDRINUSE" not in str(e) or attempt == 2:
                raise


def _rank_main(rank, world, port, cfg, kind, token_len, graph, out):
    import torch.distributed as dist
    from zgml_b200 import CudaBackend
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)   # only carries the NCCL id
    be = CudaBackend(rank)
    try:
        be.comm_init_torch()
        be.set_graph_mode(graph)
        w = shard_weights(synthetic_weights(cfg, kind, seed=11, embed_scale=1.0), rank, world)
        sess = DeviceLlamaSession(be, cfg, w, token_len)
        if token_len == 1:
            toks, logs = greedy(sess, 1, 6)
        else:
            logs = sess.execute_at([(5 * i + 2) % cfg.vocab_size for i in range(token_len)], 0).copy()[None]
        sess.close()
        out[rank] = logs
    finally:
        be.close()
        dist.destroy_process_group()


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("cfg,kind,token_len,graph", [(CFG, "q4_0", 1, True), (CFG_TIED, "q8_0", 1, True), (CFG, "q8_0", 1, False),
                                                      (CFG, "q8_0", 4, True), (CFG, "q4_0", 24, True)],
                         ids=["untied-q4-graph", "tied-q8-graph", "untied-q8-eager", "chunk4", "chunk24-tensor-core"])
def test_two_gpu_sharded_decode_matches_unsharded_oracle(cfg, kind, token_len, graph):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    out = mp.Manager().dict()
    _spawn(_rank_main, world, (cfg, kind, token_len, graph, out))
    w = synthetic_weights(cfg, kind, seed=11, embed_scale=1.0)
    ref = DeviceLlamaSession(OracleBackend(), cfg, w, token_len)
    if token_len == 1:
        _, want = greedy(ref, 1, 6)
    else:
        want = ref.execute_at([(5 * i + 2) % cfg.vocab_size for i in range(token_len)], 0).copy()[None]
    ref.close()
    for r in range(world):
        assert rel(out[r], want) < 1e-3
        assert (np.argmax(out[r], axis=1) == np.argmax(want, axis=1)).all()
    assert np.array_equal(out[0], out[1])


# ── Llama-3-70B WIDTH (d_model 8192, 64 / 8 heads, d_ff 28672), 2 layers, world 2 / 4 / 8: the shapes of BASELINE.json config 5 ──
CFG_70B_WIDTH = LlamaConfig(vocab_size=8192, d_model=8192, n_layers=2, n_heads=64, n_kv_heads=8, d_ff=28672, max_seq_len=64,
                            rope_base=5e5, tied_lm_head=False)


def _rank_main_synth(rank, world, port, cfg, kind, peer, fused, context, n_tok, out):
    import torch.distributed as dist
    if not peer:
        os.environ["ZG_CUDA_PEER"] = "0"          # NCCL all-reduces instead of the NVLink peer-memory kernel
    if fused == "arnorm":
        os.environ["ZG_CUDA_AR_NORM"] = "1"       # all-reduce + the norm block behind it in one launch (ops.cu k_allreduce_norm)
    elif fused:
        os.environ["ZG_CUDA_DECODE"] = "1"        # the persistent fused decode kernel (peer all-reduce phases inside)
    from zgml_b200 import CudaBackend
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    be = CudaBackend(rank)
    try:
        be.comm_init_torch()
        w, handles = synthetic_resident_shard(be, cfg, kind, seed=17, rank=rank, world=world)   # slices of ONE model, whatever the world size
        sess = DeviceLlamaSession(be, cfg, w, 1)
        sess.pos = context
        toks, logs = greedy(sess, 1, n_tok)
        out[rank] = (logs, be.comm_mode(), be.program_stats(sess.handle)["fused_decode_layers"])
        sess.close()
        for h in handles:
            h.free()
    finally:
        be.close()
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("path", ["peer", "nccl", "fused", "arnorm"])
def test_sharded_70b_width_matches_unsharded_oracle(world, path):
    """Every rank holds its slab of the same hashed synthetic model; rank-identical logits, 1e-3 against the unsharded oracle
    executor on the host form of that model, identical greedy tokens.  Self-skips below `world` GPUs."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs (gpurun --gpus {world})")
    import torch.multiprocessing as mp
    from zgml_b200.host.llama import synthetic_model_host
    cfg, kind, n_tok, context = CFG_70B_WIDTH, "q4_0", 3, 20
    out = mp.Manager().dict()
    _spawn(_rank_main_synth, world, (cfg, kind, path != "nccl", "arnorm" if path == "arnorm" else path == "fused", context, n_tok, out))
    ref = DeviceLlamaSession(OracleBackend(native=True), cfg, synthetic_model_host(cfg, kind, seed=17), 1)
    ref.pos = context
    _, want = greedy(ref, 1, n_tok)
    ref.close()
    for r in range(world):
        logs, mode, fused_layers = out[r]
        assert mode == ("nccl" if path == "nccl" else "nvlink-peer+nccl")   # the path under test really ran
        assert fused_layers == (cfg.n_layers if path == "fused" else 0)
        assert rel(logs, want) < 1e-3
        assert (np.argmax(logs, axis=1) == np.argmax(want, axis=1)).all()
        assert np.array_equal(logs, out[0][0])                 # bit-identical on every rank
