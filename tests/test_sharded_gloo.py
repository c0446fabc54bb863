"""Row-sharded LLaMA decode on the CPU, world_size 2 over gloo: every rank lowers its shard
(zgml_b200/host/llama.py `shard_weights` + `build_program`), the oracle executor runs the ops and gloo the
collectives.  The sharded logits must equal the unsharded oracle run on the same weights (the all-reduce only
regroups the k-sum: 1e-5 relative), greedy tokens identical."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from llama_reference import OracleBackend, ShardedOracleBackend
from oracle import oracle
from zgml_b200 import abi
from zgml_b200.host.llama import (LLAMA3_70B, DeviceLlamaSession, LlamaConfig, build_program, check_shardable,
                                  shard_weights, slice_columns, slice_rows, synthetic_weights)

TIED = LlamaConfig(vocab_size=128, d_model=128, n_layers=2, n_heads=4, n_kv_heads=2, d_ff=192, max_seq_len=16)
UNTIED = LlamaConfig(vocab_size=192, d_model=128, n_layers=2, n_heads=4, n_kv_heads=4, d_ff=256, max_seq_len=16,
                     rope_base=5e5, tied_lm_head=False)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, cfg, kind, token_len, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = shard_weights(synthetic_weights(cfg, kind, seed=11, embed_scale=1.0), rank, world)
        sess = DeviceLlamaSession(ShardedOracleBackend(dist), cfg, w, token_len)
        logs = []
        if token_len == 1:
            t = 1
            for _ in range(4):
                lg = sess.step(t).copy()
                t = int(np.argmax(lg))
                logs.append(lg)
        else:
            logs.append(sess.execute_at([(5 * i + 2) % cfg.vocab_size for i in range(token_len)], 0).copy())
        sess.close()
        out[rank] = np.stack(logs)
    finally:
        dist.destroy_process_group()


def _reference(cfg, kind, token_len):
    w = synthetic_weights(cfg, kind, seed=11, embed_scale=1.0)
    sess = DeviceLlamaSession(OracleBackend(), cfg, w, token_len)
    logs = []
    if token_len == 1:
        t = 1
        for _ in range(4):
            lg = sess.step(t).copy()
            t = int(np.argmax(lg))
            logs.append(lg)
    else:
        logs.append(sess.execute_at([(5 * i + 2) % cfg.vocab_size for i in range(token_len)], 0).copy())
    sess.close()
    return np.stack(logs)


@pytest.mark.parametrize("cfg,kind,token_len", [(TIED, "q8_0", 1), (UNTIED, "q4_0", 1), (UNTIED, "q8_0", 3)],
                         ids=["tied-gqa-q8", "untied-mha-q4", "untied-chunk3"])
def test_two_rank_sharded_decode_equals_unsharded(cfg, kind, token_len):
    world = 2
    oracle.lib()   # build once before forking
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rank_main, args=(world, _free_port(), cfg, kind, token_len, out), nprocs=world, join=True)
    want = _reference(cfg, kind, token_len)
    for r in range(world):
        got = out[r]
        assert got.shape == want.shape
        assert np.max(np.abs(got - want)) <= 1e-5 * np.max(np.abs(want))
        assert (np.argmax(got, axis=1) == np.argmax(want, axis=1)).all()
    assert np.array_equal(out[0], out[1])   # every rank ends with the same logits


def test_slabs_keep_whole_quant_blocks():
    r = np.random.default_rng(0)
    K, N = 64, 96
    w = r.standard_normal((K, N)).astype(np.float32)
    full = oracle.QuantizedWeight.from_slice(w, K, N, 32)
    from zgml_b200.backend import QuantizedWeightUpload
    up = QuantizedWeightUpload(full.data, full.scales, K, N, 32)
    deq = full.dequantize_to()
    c = slice_columns(up, 32, 96)
    assert np.array_equal(oracle.QuantizedWeight(c.data, c.scales, K, 64, 32).dequantize_to(), deq[:, 32:96])
    rr = slice_rows(up, 16, 48)
    assert np.array_equal(oracle.QuantizedWeight(rr.data, rr.scales, 32, N, 32).dequantize_to(), deq[16:48])
    # column slabs concatenate to the whole product; row slabs sum to it
    x = r.standard_normal(K).astype(np.float32)
    want = full.matmul(x, 1)[0]
    a = oracle.QuantizedWeight(*[getattr(slice_columns(up, 0, 32), f) for f in ("data", "scales", "rows", "cols", "block_size")]).matmul(x, 1)[0]
    assert np.array_equal(a, want[:32])
    lo, hi = slice_rows(up, 0, 32), slice_rows(up, 32, 64)
    s = (oracle.QuantizedWeight(lo.data, lo.scales, 32, N, 32).matmul(x[:32], 1)[0]
         + oracle.QuantizedWeight(hi.data, hi.scales, 32, N, 32).matmul(x[32:], 1)[0])
    assert np.allclose(s, want, rtol=0, atol=1e-5 * np.max(np.abs(want)))


def test_llama3_70b_shards_over_2_4_8():
    for g in (2, 4, 8):
        check_shardable(LLAMA3_70B, g)            # 1024 / 128 / 3584 / 16032 at g = 8: all multiples of 32
    with pytest.raises(ValueError):
        check_shardable(LLAMA3_70B, 16)           # 8 KV heads do not divide over 16


def test_sharded_program_has_two_allreduces_per_layer_and_one_allgather():
    w = shard_weights(synthetic_weights(UNTIED, "q8_0", seed=1), 1, 2)
    lp = build_program(UNTIED, w, 1)
    tags = [o.tag for o in lp.program.ops]
    assert tags.count(abi.OP_ALLREDUCE) == 2 * UNTIED.n_layers and tags.count(abi.OP_ALLGATHER) == 1
    assert tags.count(abi.OP_ATTENTION) == UNTIED.n_heads // 2 * UNTIED.n_layers
    assert lp.program.qweights[0].cols == UNTIED.d_model // 2 and lp.program.qweights[3].rows == UNTIED.d_model // 2
