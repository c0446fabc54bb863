"""CPU tests of the GGUF container reader / writer and the LLaMA loader glue (zgml_b200/host/gguf.py).

The container tests are the reference's own (src/gguf.zig:569-911), rebuilt byte for byte with the same synthetic
buffers; the block decode is checked bit-exact against the oracle's restatement of quantizedWeightFromInfo / loadTensor
(src/models/gguf_loader.zig:33-154) and its golden blocks (:484-572); the loader is exercised end to end through a
GGUF file written by our tool and the reference executor."""
import struct

import numpy as np
import pytest

from llama_reference import OracleBackend
from oracle import oracle
from zgml_b200.host import gguf
from zgml_b200.host.gguf import GGMLType, GGUFError, GGUFFile, GGUFWriter, MetaValueType, TensorInfo
from zgml_b200.host.llama import DeviceLlamaSession, LlamaConfig, linear_shapes


def u32(v): return struct.pack("<I", v)
def u64(v): return struct.pack("<Q", v)
def s(x): return u64(len(x)) + x.encode()
def pad(b, a=32): return b + bytes(GGUFFile.align_up(len(b), a) - len(b))


def build_test_buffer():  # buildTestBuffer, src/gguf.zig:505-567
    b = u32(0x46554747) + u32(3) + u64(2) + u64(3)
    b += s("general.architecture") + u32(MetaValueType.string) + s("llama")
    b += s("general.name") + u32(MetaValueType.string) + s("test-model")
    b += s("llama.context_length") + u32(MetaValueType.uint32) + u32(2048)
    b += s("token_embd.weight") + u32(2) + u64(128) + u64(64) + u32(GGMLType.f32) + u64(0)
    b += s("output.weight") + u32(2) + u64(64) + u64(32) + u32(GGMLType.f32) + u64(128 * 64 * 4)
    b = pad(b)
    return b + bytes(i & 0xFF for i in range(128 * 64 * 4 + 64 * 32 * 4))


def test_parse_synthetic_v3_buffer_metadata_and_tensor_infos():  # src/gguf.zig:569-664
    gf = GGUFFile.parse_buffer(build_test_buffer())
    assert gf.version == 3
    assert gf.get_meta_string("general.architecture") == "llama" and gf.get_meta_string("general.name") == "test-model"
    assert gf.get_meta_u32("llama.context_length") == 2048
    assert gf.get_meta("llama.context_length").type == MetaValueType.uint32
    assert gf.get_meta_string("nonexistent.key") is None and gf.get_meta_u32("nonexistent.key") is None and gf.get_meta("nonexistent.key") is None
    e = gf.get_tensor_info("token_embd.weight")
    assert (e.n_dims, e.dims[0], e.dims[1], e.type_, e.offset, e.n_elems(), e.data_size()) == (2, 128, 64, GGMLType.f32, 0, 128 * 64, 128 * 64 * 4)
    o = gf.get_tensor_info("output.weight")
    assert (o.n_dims, o.dims[0], o.dims[1], o.type_, o.offset, o.n_elems(), o.data_size()) == (2, 64, 32, GGMLType.f32, 128 * 64 * 4, 64 * 32, 64 * 32 * 4)
    assert gf.get_tensor_info("nonexistent") is None


def test_data_offset_alignment_and_tensor_data():  # src/gguf.zig:666-694
    buf = build_test_buffer()
    gf = GGUFFile.parse_buffer(buf)
    assert gf.data_offset % 32 == 0 and gf.data_offset > 0
    e, o = gf.get_tensor_info("token_embd.weight"), gf.get_tensor_info("output.weight")
    assert gf.get_tensor_data(e).size == 128 * 64 * 4 and gf.get_tensor_data(o).size == 64 * 32 * 4
    assert bytes(gf.get_tensor_data(e)[:5]) == bytes([0, 1, 2, 3, 4])
    assert gf.get_tensor_data(o)[0] == (128 * 64 * 4) & 0xFF      # starts right after the first tensor
    assert gf.get_tensor_f32(e).size == 128 * 64


def test_literal_magic_byte_order():  # src/gguf.zig:579-597
    gf = GGUFFile.parse_buffer(pad(b"GGUF" + u32(3) + u64(0) + u64(0)))
    assert gf.version == 3 and len(gf.tensors) == 0


def test_ggml_type_block_and_type_sizes():  # src/gguf.zig:696-718
    for t, bs, ts in [(GGMLType.f32, 1, 4), (GGMLType.f16, 1, 2), (GGMLType.f64, 1, 8), (GGMLType.i8, 1, 1), (GGMLType.i32, 1, 4),
                      (GGMLType.q4_0, 32, 18), (GGMLType.q8_0, 32, 34), (GGMLType.q4_k, 256, 144)]:
        assert (t.block_size, t.type_size) == (bs, ts)
    assert all(t.type_size > 0 and t.block_size in (1, 32, 256) for t in GGMLType)


def test_tensor_info_n_elems_and_data_size():  # src/gguf.zig:720-754
    assert (TensorInfo("test", 1, (100, 1, 1, 1), GGMLType.f32, 0).n_elems(), TensorInfo("test", 1, (100, 1, 1, 1), GGMLType.f32, 0).data_size()) == (100, 400)
    t3 = TensorInfo("test3d", 3, (10, 20, 30, 1), GGMLType.f16, 0)
    assert (t3.n_elems(), t3.data_size()) == (6000, 12000)
    tq = TensorInfo("quantized", 1, (256, 1, 1, 1), GGMLType.q4_0, 0)
    assert (tq.n_elems(), tq.data_size()) == (256, 144)            # 8 blocks x 18 bytes


def test_v2_buffer_parsing():  # src/gguf.zig:756-797
    b = u32(0x46554747) + u32(2) + u32(1) + u32(1)
    b += s("general.architecture") + u32(MetaValueType.string) + s("test")
    b += s("w") + u32(1) + u64(4) + u32(GGMLType.f32) + u64(0)
    gf = GGUFFile.parse_buffer(pad(b) + bytes(16))
    assert gf.version == 2 and gf.get_meta_string("general.architecture") == "test"
    w = gf.get_tensor_info("w")
    assert (w.n_elems(), w.data_size()) == (4, 16)


def test_custom_alignment_via_general_alignment():  # src/gguf.zig:799-838
    b = u32(0x46554747) + u32(3) + u64(1) + u64(1)
    b += s("general.alignment") + u32(MetaValueType.uint32) + u32(64)
    b += s("x") + u32(1) + u64(2) + u32(GGMLType.f32) + u64(0)
    header_end = len(b)
    gf = GGUFFile.parse_buffer(pad(b, 64) + bytes(8))
    assert gf.data_offset % 64 == 0 and gf.data_offset >= header_end


def test_invalid_magic_and_unsupported_version_rejected():  # src/gguf.zig:840-863
    with pytest.raises(GGUFError, match="InvalidMagic"):
        GGUFFile.parse_buffer(u32(0xDEADBEEF) + bytes(28))
    with pytest.raises(GGUFError, match="UnsupportedVersion"):
        GGUFFile.parse_buffer(u32(0x46554747) + u32(99) + bytes(24))
    with pytest.raises(GGUFError, match="UnsupportedGGMLType"):
        GGUFFile.parse_buffer(pad(u32(0x46554747) + u32(3) + u64(1) + u64(0) + s("t") + u32(1) + u64(4) + u32(4) + u64(0)))   # type 4 was removed


def test_array_metadata_value_and_align_up():  # src/gguf.zig:865-911
    b = u32(0x46554747) + u32(3) + u64(0) + u64(1)
    b += s("tokenizer.scores") + u32(MetaValueType.array) + u32(MetaValueType.float32) + u64(3) + struct.pack("<fff", 1.5, 2.5, 3.5)
    v = GGUFFile.parse_buffer(pad(b)).get_meta("tokenizer.scores")
    assert v.type == MetaValueType.array and v.value.elem_type == MetaValueType.float32 and v.value.len == 3 and len(v.value.data) == 12
    assert [GGUFFile.align_up(*a) for a in [(0, 32), (1, 32), (31, 32), (32, 32), (33, 32), (33, 64), (65, 64)]] == [0, 32, 32, 32, 64, 64, 128]


def test_writer_round_trip_all_metadata_types():
    w = GGUFWriter(version=3, alignment=64)
    w.add_meta("a.u8", MetaValueType.uint8, 200); w.add_meta("a.i8", MetaValueType.int8, -5)
    w.add_meta("a.u16", MetaValueType.uint16, 60000); w.add_meta("a.i16", MetaValueType.int16, -300)
    w.add_meta("a.i32", MetaValueType.int32, -7); w.add_meta("a.f32", MetaValueType.float32, 0.5)
    w.add_meta("a.bool", MetaValueType.bool_, 1); w.add_meta("a.u64", MetaValueType.uint64, 1 << 40)
    w.add_meta("a.i64", MetaValueType.int64, -(1 << 40)); w.add_meta("a.f64", MetaValueType.float64, 0.25)
    w.add_meta("a.strs", MetaValueType.array, ["x", "yz"], elem_type=MetaValueType.string)
    w.add_tensor("t0", (3,), GGMLType.f32, np.arange(3, dtype="<f4"))
    w.add_tensor("t1", (32, 2), GGMLType.q8_0, np.arange(68, dtype=np.uint8))
    gf = GGUFFile.parse_buffer(w.tobytes())
    assert gf.data_offset % 64 == 0 and gf.get_meta_u32("general.alignment") == 64
    assert [gf.get_meta(k).value for k in ("a.u8", "a.i8", "a.u16", "a.i16", "a.i32", "a.f32", "a.bool", "a.u64", "a.i64", "a.f64")] == \
        [200, -5, 60000, -300, -7, 0.5, True, 1 << 40, -(1 << 40), 0.25]
    assert gf.get_meta_u32("a.i32") is None and gf.get_meta_u32("a.u64") is None      # negative / too large for u32
    assert gf.get_meta("a.strs").value.len == 2
    assert np.array_equal(gf.get_tensor_f32(gf.get_tensor_info("t0")), [0, 1, 2])
    t1 = gf.get_tensor_info("t1")
    assert t1.offset % 64 == 0 and bytes(gf.get_tensor_data(t1)) == bytes(range(68))
    with pytest.raises(GGUFError):
        w.add_tensor("bad", (32, 2), GGMLType.q8_0, np.zeros(10, np.uint8))


@pytest.mark.parametrize("t,ggml", [(GGMLType.q8_0, 8), (GGMLType.q4_0, 2)])
def test_block_decode_matches_reference_rule(t, ggml):  # src/models/gguf_loader.zig:99-154,171-204
    K, N = 64, 96
    r = np.random.default_rng(int(t))
    nb = K * N // 32
    raw = r.integers(0, 256, (nb, t.type_size), dtype=np.uint8)
    raw[:, 0:2] = r.uniform(-2, 2, nb).astype(np.float16).view(np.uint8).reshape(nb, 2)
    raw = raw.ravel()
    info = TensorInfo("w", 2, (K, N, 1, 1), t, 0)
    qw = gguf.quantized_weight_from_info(info, raw)
    o = oracle.QuantizedWeight.from_gguf(raw, ggml, K, N)
    assert (qw.rows, qw.cols, qw.block_size) == (K, N, 32)
    assert np.array_equal(qw.data, o.data) and np.array_equal(qw.scales.view(np.uint32), o.scales.view(np.uint32))
    w = GGUFWriter(); w.add_tensor("w", (K, N), t, raw)
    deq = gguf.load_tensor_f32(GGUFFile.parse_buffer(w.tobytes()), "w")
    want = oracle.dequant_q8_0(raw, K * N) if ggml == 8 else oracle.dequant_q4_0(raw, K * N)
    assert np.array_equal(deq.view(np.uint32), want.view(np.uint32))
    with pytest.raises(GGUFError, match="UnsupportedShape"):
        gguf.quantized_weight_from_info(TensorInfo("w", 1, (K * N, 1, 1, 1), t, 0), raw)
    with pytest.raises(GGUFError, match="UnsupportedType"):
        gguf.quantized_weight_from_info(TensorInfo("w", 2, (K, N, 1, 1), GGMLType.q4_k, 0), raw)


def test_golden_blocks_from_reference_tests():  # src/models/gguf_loader.zig:484-572
    blk = np.zeros(18, np.uint8)
    blk[0:2] = np.array([1.0], np.float16).view(np.uint8)
    blk[2] = 0xF8                                                  # elem 0 = 8 - 8 = 0, elem 1 = 15 - 8 = 7, the rest 0 - 8 = -8
    qw = gguf.quantized_weight_from_info(TensorInfo("w", 2, (16, 2, 1, 1), GGMLType.q4_0, 0), blk)
    assert list(qw.data[:4]) == [0, 7, -8, -8] and qw.scales[0] == 1.0
    blk8 = np.zeros(34, np.uint8)
    blk8[0:2] = np.array([0.5], np.float16).view(np.uint8)
    blk8[2:4] = np.array([10, -5], np.int8).view(np.uint8)
    qw8 = gguf.quantized_weight_from_info(TensorInfo("w", 2, (16, 2, 1, 1), GGMLType.q8_0, 0), blk8)
    assert list(qw8.data[:3]) == [10, -5, 0] and qw8.scales[0] == 0.5 and (qw8.rows, qw8.cols, qw8.block_size) == (16, 2, 32)
    w = GGUFWriter(); w.add_tensor("w", (16, 2), GGMLType.q8_0, blk8)
    assert list(gguf.load_tensor_f32(GGUFFile.parse_buffer(w.tobytes()), "w")[:3]) == [5.0, -2.5, 0.0]
    f16 = np.array([1.5, -0.25], np.float16)
    w = GGUFWriter(); w.add_tensor("h", (2,), GGMLType.f16, f16)
    assert list(gguf.load_tensor_f32(GGUFFile.parse_buffer(w.tobytes()), "h")) == [1.5, -0.25]


@pytest.mark.parametrize("kind,tied", [("q8_0", True), ("q4_0", False)])
def test_llama_gguf_file_round_trip_through_the_reference_executor(tmp_path, kind, tied):
    """write_llama_gguf -> GGUFFile.open (memory-mapped) -> config_from_gguf + load_direct_quantized -> the decode
    program on the reference executor: finite, token-dependent logits; a second load gives identical ones."""
    cfg = LlamaConfig(vocab_size=64, d_model=32, n_layers=2, n_heads=4, n_kv_heads=2, d_ff=64, max_seq_len=16, rope_base=5e5, tied_lm_head=tied)
    path = str(tmp_path / "tiny.gguf")
    gguf.write_llama_gguf(path, cfg, kind, seed=3, embed_scale=1.0)
    gf = GGUFFile.open(path)
    assert gguf.config_from_gguf(gf) == cfg                        # configFromGGUF, src/models/gguf_loader.zig:214-236
    t = GGMLType.q8_0 if kind == "q8_0" else GGMLType.q4_0
    for name, (K, N) in linear_shapes(cfg).items():
        info = gf.get_tensor_info(f"blk.1.{gguf._LINEAR_NAMES[name]}")
        assert (info.dims[0], info.dims[1], info.type_) == (K, N, t) and info.offset % 32 == 0
        assert info.data_size() == K * N // 32 * t.type_size       # TensorInfo.dataSize, src/gguf.zig:174-178
    w = gguf.load_direct_quantized(gf)
    assert (w.out_proj is None) == tied and w.token_embed.shape == (64, 32) and len(w.layers) == 2
    logits = []
    for _ in range(2):
        sess = DeviceLlamaSession(OracleBackend(), cfg, gguf.load_direct_quantized(GGUFFile.open(path)), 1)
        logits.append([sess.step(tok).copy() for tok in (1, 7)])
        sess.close()
    assert np.isfinite(logits[0][1]).all() and np.ptp(logits[0][1]) > 0 and not np.array_equal(logits[0][0], logits[0][1])
    assert np.array_equal(logits[0][1], logits[1][1])
    with pytest.raises(GGUFError, match="TensorNotFound"):
        gguf.load_tensor_f32(gf, "blk.9.attn_norm.weight")


@pytest.mark.parametrize("kind,tied", [("q8_0", True), ("q4_0", False)])
def test_block_level_shards_equal_sharding_the_expanded_weights(tmp_path, kind, tied):
    """`shard_blocks` (raw GGUF bytes, what each rank uploads) decodes to exactly the slabs `llama.shard_weights` cuts
    from the host-expanded model: column slabs for q / k / v / gate / up / head, row slabs for o / down."""
    from zgml_b200.host.llama import shard_weights
    cfg = LlamaConfig(vocab_size=128, d_model=128, n_layers=1, n_heads=4, n_kv_heads=2, d_ff=192, max_seq_len=16, tied_lm_head=tied)
    path = str(tmp_path / "m.gguf")
    gguf.write_llama_gguf(path, cfg, kind, seed=9)
    gf = GGUFFile.open(path)
    whole = gguf.load_direct_quantized(gf)
    world = 2
    for rank in range(world):
        want = shard_weights(whole, rank, world)
        names = dict(gguf._LINEAR_NAMES)
        for key, tname in names.items():
            info = gf.get_tensor_info(f"blk.0.{tname}")
            raw, K, N = gguf.shard_blocks(info, gf.get_tensor_data(info), key, rank, world)
            got = gguf.quantized_weight_from_info(TensorInfo(tname, 2, (K, N, 1, 1), info.type_, 0), raw)
            ref = want.layers[0][key]
            assert (got.rows, got.cols) == (ref.rows, ref.cols)
            assert np.array_equal(got.data, ref.data) and np.array_equal(got.scales.view(np.uint32), ref.scales.view(np.uint32))
        if not tied:
            info = gf.get_tensor_info("output.weight")
            raw, K, N = gguf.shard_blocks(info, gf.get_tensor_data(info), "out_proj", rank, world)
            got = gguf.quantized_weight_from_info(TensorInfo("o", 2, (K, N, 1, 1), info.type_, 0), raw)
            assert np.array_equal(got.data, want.out_proj.data) and np.array_equal(got.scales, want.out_proj.scales)
    info = gf.get_tensor_info("blk.0.attn_q.weight")
    raw, K, N = gguf.shard_blocks(info, gf.get_tensor_data(info), "wq", 0, 1)
    assert (K, N) == (128, 128) and raw.size == info.data_size()
