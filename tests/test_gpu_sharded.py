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
    mp.spawn(_rank_main, args=(world, _free_port(), cfg, kind, token_len, graph, out), nprocs=world, join=True)
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
