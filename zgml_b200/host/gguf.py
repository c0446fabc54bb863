"""GGUF container and the GGUF -> device upload path (host side).

Mirror of zgml's reader, same names and error behaviour:
  * `GGMLType.block_size / type_size`, `TensorInfo.n_elems / data_size`, `GGUFFile.parse_buffer / open / get_meta /
    get_meta_u32 / get_meta_string / get_tensor_info / get_tensor_data / get_tensor_f32 / align_up`
    — src/gguf.zig:30-112,159-181,183-346,461-463 (v2 and v3 headers, `general.alignment`, default 32).
  * `config_from_gguf`, `quantized_weight_from_info`, `load_tensor_f32`, `load_direct_quantized`
    — src/models/gguf_loader.zig:99-154,171-204,214-236,340-391 (tensor names :258-283).
  * `upload_quantized_tensor`, `load_resident` — what changes for the device: the raw Q8_0 / Q4_0 block bytes of
    `get_tensor_data` go straight to `zg_cuda_qweight_upload_gguf` (expanded and repacked on the GPU), one tensor at a
    time; the host never holds the 1.125 B / weight int8 expansion that `quantizedWeightFromInfo` makes.
  * `GGUFWriter`, `write_llama_gguf` — our tool for the synthetic random-init models of BASELINE.json's configs, in zgml's
    convention (2-D linears `dims = [K, N]`, blocks over the flat `[K, N]` array; `token_embd.weight` `[d_model, vocab]`).

No arithmetic on weights beyond the format decode the reference's loader performs; no oracle involved.
"""
import enum
import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from ..backend import QuantizedWeightUpload, ResidentQuantizedWeight
from .llama import LINEARS, LlamaConfig, LlamaWeights, check_shardable, linear_shapes

GGUF_MAGIC = 0x46554747  # "GGUF" little-endian, src/gguf.zig:191
DEFAULT_ALIGNMENT = 32   # src/gguf.zig:192


class GGUFError(ValueError):
    """error.InvalidFormat / InvalidMagic / UnsupportedVersion / UnsupportedGGMLType / TensorNotFound / UnsupportedType"""


class GGMLType(enum.IntEnum):  # src/gguf.zig:30-61
    f32 = 0
    f16 = 1
    q4_0 = 2
    q4_1 = 3
    q5_0 = 6
    q5_1 = 7
    q8_0 = 8
    q8_1 = 9
    q2_k = 10
    q3_k = 11
    q4_k = 12
    q5_k = 13
    q6_k = 14
    q8_k = 15
    iq2_xxs = 16
    iq2_xs = 17
    iq3_xxs = 18
    iq1_s = 19
    iq4_nl = 20
    iq3_s = 21
    iq2_s = 22
    iq4_xs = 23
    i8 = 24
    i16 = 25
    i32 = 26
    i64 = 27
    f64 = 28
    iq1_m = 29

    @property
    def block_size(self) -> int:  # blockSize, src/gguf.zig:65-77
        if self in (GGMLType.f32, GGMLType.f16, GGMLType.f64, GGMLType.i8, GGMLType.i16, GGMLType.i32, GGMLType.i64):
            return 1
        if self in (GGMLType.q4_0, GGMLType.q4_1, GGMLType.q5_0, GGMLType.q5_1, GGMLType.q8_0, GGMLType.q8_1, GGMLType.iq4_nl):
            return 32
        return 256

    @property
    def type_size(self) -> int:  # typeSize, src/gguf.zig:81-112
        return _TYPE_SIZE[self]


_TYPE_SIZE = {GGMLType.f32: 4, GGMLType.f16: 2, GGMLType.f64: 8, GGMLType.i8: 1, GGMLType.i16: 2, GGMLType.i32: 4, GGMLType.i64: 8,
              GGMLType.q4_0: 18, GGMLType.q4_1: 20, GGMLType.q5_0: 22, GGMLType.q5_1: 24, GGMLType.q8_0: 34, GGMLType.q8_1: 40,
              GGMLType.q2_k: 84, GGMLType.q3_k: 110, GGMLType.q4_k: 144, GGMLType.q5_k: 176, GGMLType.q6_k: 210, GGMLType.q8_k: 292,
              GGMLType.iq2_xxs: 66, GGMLType.iq2_xs: 74, GGMLType.iq2_s: 82, GGMLType.iq3_xxs: 98, GGMLType.iq3_s: 110,
              GGMLType.iq1_s: 50, GGMLType.iq1_m: 56, GGMLType.iq4_nl: 18, GGMLType.iq4_xs: 136}


class MetaValueType(enum.IntEnum):  # src/gguf.zig:116-130
    uint8 = 0
    int8 = 1
    uint16 = 2
    int16 = 3
    uint32 = 4
    int32 = 5
    float32 = 6
    bool_ = 7
    string = 8
    array = 9
    uint64 = 10
    int64 = 11
    float64 = 12


_SCALAR_FMT = {MetaValueType.uint8: "<B", MetaValueType.int8: "<b", MetaValueType.uint16: "<H", MetaValueType.int16: "<h",
               MetaValueType.uint32: "<I", MetaValueType.int32: "<i", MetaValueType.float32: "<f", MetaValueType.bool_: "<B",
               MetaValueType.uint64: "<Q", MetaValueType.int64: "<q", MetaValueType.float64: "<d"}


@dataclass(frozen=True)
class ArrayValue:  # src/gguf.zig:134-138: element type + length + the raw element bytes
    elem_type: MetaValueType
    len: int
    data: bytes


@dataclass(frozen=True)
class MetaValue:  # tagged union of src/gguf.zig:142-156
    type: MetaValueType
    value: object


@dataclass(frozen=True)
class TensorInfo:  # src/gguf.zig:159-181
    name: str
    n_dims: int
    dims: Tuple[int, int, int, int]
    type_: GGMLType
    offset: int  # within the data section

    def n_elems(self) -> int:
        n = 1
        for d in self.dims[:self.n_dims]:
            n *= d
        return n

    def data_size(self) -> int:
        return (self.n_elems() // self.type_.block_size) * self.type_.type_size


class _Cursor:
    def __init__(self, buf):
        self.buf, self.pos = buf, 0

    def take(self, n: int) -> bytes:
        if self.pos + n > len(self.buf):
            raise GGUFError("GGUF: unexpected end of buffer")  # the reference panics here (src/gguf.zig:364)
        out = bytes(self.buf[self.pos:self.pos + n])
        self.pos += n
        return out

    def val(self, fmt: str):
        return struct.unpack(fmt, self.take(struct.calcsize(fmt)))[0]

    def string(self) -> str:
        return self.take(self.val("<Q")).decode("utf-8", errors="surrogateescape")

    def skip(self, vtype: MetaValueType):  # skipMetaValue, src/gguf.zig:437-458
        if vtype == MetaValueType.string:
            self.take(self.val("<Q"))
        elif vtype == MetaValueType.array:
            et, n = MetaValueType(self.val("<I")), self.val("<Q")
            for _ in range(n):
                self.skip(et)
        else:
            self.take(struct.calcsize(_SCALAR_FMT[vtype]))

    def meta_value(self) -> MetaValue:  # readMetaValue, src/gguf.zig:400-435
        vtype = MetaValueType(self.val("<I"))
        if vtype == MetaValueType.string:
            return MetaValue(vtype, self.string())
        if vtype == MetaValueType.array:
            et, n = MetaValueType(self.val("<I")), self.val("<Q")
            start = self.pos
            for _ in range(n):
                self.skip(et)
            return MetaValue(vtype, ArrayValue(et, n, bytes(self.buf[start:self.pos])))
        v = self.val(_SCALAR_FMT[vtype])
        return MetaValue(vtype, bool(v) if vtype == MetaValueType.bool_ else v)


class GGUFFile:
    """A parsed GGUF file: raw buffer, metadata KV pairs, tensor infos, aligned data offset (src/gguf.zig:183-346)."""

    def __init__(self, raw, version, metadata, tensors, data_offset):
        self.raw_data, self.version, self.metadata, self.tensors, self.data_offset = raw, version, metadata, tensors, data_offset

    @staticmethod
    def align_up(offset: int, alignment: int) -> int:  # src/gguf.zig:461-463
        return (offset + alignment - 1) & ~(alignment - 1)

    @classmethod
    def open(cls, path: str) -> "GGUFFile":  # src/gguf.zig:195-210; memory-mapped, tensor data is never copied
        raw = np.memmap(path, dtype=np.uint8, mode="r")
        if raw.size < 24:
            raise GGUFError("InvalidFormat")
        return cls.parse_buffer(raw)

    @classmethod
    def parse_buffer(cls, buf) -> "GGUFFile":  # src/gguf.zig:214-292
        raw = buf if isinstance(buf, np.ndarray) else np.frombuffer(bytes(buf), dtype=np.uint8)
        c = _Cursor(raw)
        if c.val("<I") != GGUF_MAGIC:
            raise GGUFError("InvalidMagic")
        version = c.val("<I")
        if version < 2 or version > 3:
            raise GGUFError("UnsupportedVersion")
        count_fmt = "<Q" if version >= 3 else "<I"   # v3 counts are u64, v2 u32
        tensor_count, kv_count = c.val(count_fmt), c.val(count_fmt)
        metadata: Dict[str, MetaValue] = {}
        for _ in range(kv_count):
            key = c.string()
            metadata[key] = c.meta_value()
        tensors: Dict[str, TensorInfo] = {}
        for _ in range(tensor_count):
            name = c.string()
            n_dims = c.val("<I")
            dims = [1, 1, 1, 1]
            for d in range(n_dims):
                dims[d] = c.val("<Q")
            type_raw = c.val("<I")
            try:
                type_ = GGMLType(type_raw)
            except ValueError:
                raise GGUFError("UnsupportedGGMLType") from None
            tensors[name] = TensorInfo(name, n_dims, tuple(dims), type_, c.val("<Q"))
        alignment = DEFAULT_ALIGNMENT
        a = metadata.get("general.alignment")
        if a is not None and a.type in (MetaValueType.uint32, MetaValueType.uint64, MetaValueType.int32):
            alignment = int(a.value)
        return cls(raw, version, metadata, tensors, cls.align_up(c.pos, alignment))

    def get_meta(self, key: str) -> Optional[MetaValue]:
        return self.metadata.get(key)

    def get_meta_u32(self, key: str) -> Optional[int]:  # src/gguf.zig:307-315
        v = self.metadata.get(key)
        if v is None:
            return None
        if v.type == MetaValueType.uint32:
            return v.value
        if v.type == MetaValueType.int32:
            return v.value if v.value >= 0 else None
        if v.type == MetaValueType.uint64:
            return v.value if v.value <= 0xFFFFFFFF else None
        return None

    def get_meta_string(self, key: str) -> Optional[str]:
        v = self.metadata.get(key)
        return v.value if v is not None and v.type == MetaValueType.string else None

    def get_tensor_info(self, name: str) -> Optional[TensorInfo]:
        return self.tensors.get(name)

    def get_tensor_data(self, info: TensorInfo) -> np.ndarray:  # src/gguf.zig:333-337: a view, no copy
        start = self.data_offset + info.offset
        return self.raw_data[start:start + info.data_size()]

    def get_tensor_f32(self, info: TensorInfo) -> np.ndarray:  # src/gguf.zig:341-346
        assert info.type_ == GGMLType.f32
        return np.frombuffer(self.get_tensor_data(info), dtype="<f4")


class GGUFWriter:
    """Writes the container `GGUFFile` reads (header, KV pairs, tensor infos, padding, aligned tensor data)."""

    def __init__(self, version: int = 3, alignment: int = DEFAULT_ALIGNMENT):
        assert version in (2, 3)
        self.version, self.alignment = version, alignment
        self._meta: List[bytes] = []
        self._tensors: List[Tuple[str, Tuple[int, ...], GGMLType, np.ndarray]] = []
        if alignment != DEFAULT_ALIGNMENT:
            self.add_meta("general.alignment", MetaValueType.uint32, alignment)

    @staticmethod
    def _str(s: str) -> bytes:
        b = s.encode("utf-8")
        return struct.pack("<Q", len(b)) + b

    def add_meta(self, key: str, vtype: MetaValueType, value, elem_type: Optional[MetaValueType] = None):
        out = self._str(key) + struct.pack("<I", int(vtype))
        if vtype == MetaValueType.string:
            out += self._str(value)
        elif vtype == MetaValueType.array:
            out += struct.pack("<IQ", int(elem_type), len(value))
            for v in value:
                out += self._str(v) if elem_type == MetaValueType.string else struct.pack(_SCALAR_FMT[elem_type], v)
        else:
            out += struct.pack(_SCALAR_FMT[vtype], value)
        self._meta.append(out)

    def add_tensor(self, name: str, dims, type_: GGMLType, data):
        data = np.ascontiguousarray(data).view(np.uint8).ravel()
        info = TensorInfo(name, len(dims), tuple(list(dims) + [1] * (4 - len(dims))), type_, 0)
        if data.size != info.data_size():
            raise GGUFError(f"{name}: {data.size} bytes given, dataSize() is {info.data_size()}")
        self._tensors.append((name, tuple(dims), type_, data))

    def tobytes(self) -> bytes:
        cnt = "<QQ" if self.version >= 3 else "<II"
        head = struct.pack("<II", GGUF_MAGIC, self.version) + struct.pack(cnt, len(self._tensors), len(self._meta)) + b"".join(self._meta)
        infos, offset = b"", 0
        for name, dims, type_, data in self._tensors:
            infos += self._str(name) + struct.pack("<I", len(dims)) + b"".join(struct.pack("<Q", d) for d in dims)
            infos += struct.pack("<IQ", int(type_), offset)
            offset = GGUFFile.align_up(offset + data.size, self.alignment)
        head += infos
        out = bytearray(head) + bytes(GGUFFile.align_up(len(head), self.alignment) - len(head))
        for _, _, _, data in self._tensors:
            out += data.tobytes()
            out += bytes(GGUFFile.align_up(len(out), self.alignment) - len(out))
        return bytes(out)

    def write(self, path: str):
        with open(path, "wb") as f:
            f.write(self.tobytes())


# ── block decode (what the reference's loader does on the host) ─────────────────────────────────────────────────
def is_direct_quantized_matmul_type(t: GGMLType) -> bool:  # src/models/gguf_loader.zig:95-97
    return t in (GGMLType.q8_0, GGMLType.q4_0)


def _blocks(raw: np.ndarray, t: GGMLType):
    nb = raw.size // t.type_size
    blk = np.asarray(raw[:nb * t.type_size]).reshape(nb, t.type_size)
    scales = np.ascontiguousarray(blk[:, 0:2]).view("<f2").astype(np.float32).ravel()    # f16 -> f32, exact
    if t == GGMLType.q8_0:
        q = np.ascontiguousarray(blk[:, 2:]).view(np.int8)                               # q[i] = int8(raw[2 + i])
    else:
        b = blk[:, 2:]                                                                   # byte 2 + i / 2: even i -> low nibble, odd -> high
        q = np.empty((nb, 32), np.int8)
        q[:, 0::2] = (b & 0x0F).astype(np.int8) - 8
        q[:, 1::2] = (b >> 4).astype(np.int8) - 8
    return q, scales


def quantized_weight_from_info(info: TensorInfo, raw: np.ndarray) -> QuantizedWeightUpload:
    """quantizedWeightFromInfo (src/models/gguf_loader.zig:99-154): Q8_0 / Q4_0 blocks -> i8 + f32 scales,
    rows = dims[0] = K, cols = dims[1] = N, block size 32.  The host-expanded form (1.125 B / weight)."""
    if info.n_dims < 2:
        raise GGUFError("UnsupportedShape")
    if not is_direct_quantized_matmul_type(info.type_):
        raise GGUFError("UnsupportedType")
    q, scales = _blocks(raw, info.type_)
    return QuantizedWeightUpload(np.ascontiguousarray(q).ravel(), scales, info.dims[0], info.dims[1], 32)


def load_tensor_f32(gf: GGUFFile, name: str) -> np.ndarray:
    """loadTensor (src/models/gguf_loader.zig:171-204): f32 as is, f16 widened, Q8_0 / Q4_0 dequantized (f32(q) * scale)."""
    info = gf.get_tensor_info(name)
    if info is None:
        raise GGUFError(f"TensorNotFound: {name}")
    raw = gf.get_tensor_data(info)
    if info.type_ == GGMLType.f32:
        return np.frombuffer(raw, dtype="<f4").astype(np.float32)
    if info.type_ == GGMLType.f16:
        return np.frombuffer(raw, dtype="<f2").astype(np.float32)
    if is_direct_quantized_matmul_type(info.type_):
        q, scales = _blocks(raw, info.type_)
        return (q.astype(np.float32) * scales[:, None]).ravel()
    raise GGUFError("UnsupportedType")


# ── LLaMA glue ───────────────────────────────────────────────────────────────────────────────────────────────────────
_LINEAR_NAMES = {"wq": "attn_q.weight", "wk": "attn_k.weight", "wv": "attn_v.weight", "wo": "attn_output.weight",
                 "w_gate": "ffn_gate.weight", "w_up": "ffn_up.weight", "w_down": "ffn_down.weight"}   # gguf_loader.zig:258-283


def config_from_gguf(gf: GGUFFile) -> LlamaConfig:  # configFromGGUF, src/models/gguf_loader.zig:214-236 (same defaults)
    arch = gf.get_meta_string("general.architecture") or "llama"

    def u32(suffix, default):
        v = gf.get_meta_u32(f"{arch}.{suffix}")
        return default if v is None else v

    n_heads = u32("attention.head_count", 32)
    rope = gf.get_meta(f"{arch}.rope.freq_base")
    rope_base = float(rope.value) if rope is not None and rope.type in (MetaValueType.float32, MetaValueType.float64, MetaValueType.uint32) else 10000.0
    vocab = gf.get_meta_u32(f"{arch}.vocab_size")
    if vocab is None:
        ti = gf.get_tensor_info("token_embd.weight")
        vocab = ti.dims[1] if ti is not None else 32000
    return LlamaConfig(vocab_size=vocab, d_model=u32("embedding_length", 4096), n_layers=u32("block_count", 32), n_heads=n_heads,
                       n_kv_heads=u32("attention.head_count_kv", n_heads), d_ff=u32("feed_forward_length", 11008),
                       max_seq_len=u32("context_length", 2048), rope_base=rope_base,
                       tied_lm_head=gf.get_tensor_info("output.weight") is None)


def _check_linear(gf: GGUFFile, name: str, K: int, N: int) -> TensorInfo:
    info = gf.get_tensor_info(name)
    if info is None:
        raise GGUFError(f"TensorNotFound: {name}")
    if info.n_elems() != K * N:
        raise GGUFError(f"SizeMismatch: {name}")   # loadRuntimeTensor, src/models/gguf_loader.zig:297
    if not is_direct_quantized_matmul_type(info.type_):
        raise GGUFError(f"UnsupportedType: {name} is {info.type_.name}; the device path takes Q8_0 / Q4_0 linears")
    return info


def _f32_params(gf: GGUFFile, cfg: LlamaConfig):
    emb = load_tensor_f32(gf, "token_embd.weight").reshape(cfg.vocab_size, cfg.d_model)
    n1 = [load_tensor_f32(gf, f"blk.{i}.attn_norm.weight") for i in range(cfg.n_layers)]
    n2 = [load_tensor_f32(gf, f"blk.{i}.ffn_norm.weight") for i in range(cfg.n_layers)]
    return emb, n1, n2, load_tensor_f32(gf, "output_norm.weight")


def load_direct_quantized(gf: GGUFFile, cfg: Optional[LlamaConfig] = None) -> LlamaWeights:
    """loadDirectQuantized (src/models/gguf_loader.zig:340-391) in the reference's HOST form: every Q8_0 / Q4_0 linear
    expanded to i8 + f32 scales, embeddings and norms dequantized to f32."""
    cfg = cfg or config_from_gguf(gf)
    shapes = linear_shapes(cfg)
    layers = []
    for i in range(cfg.n_layers):
        layer = {}
        for key in LINEARS:
            info = _check_linear(gf, f"blk.{i}.{_LINEAR_NAMES[key]}", *shapes[key])
            layer[key] = quantized_weight_from_info(info, gf.get_tensor_data(info))
        layers.append(layer)
    emb, n1, n2, nf = _f32_params(gf, cfg)
    out_proj = None
    if not cfg.tied_lm_head:
        info = _check_linear(gf, "output.weight", cfg.d_model, cfg.vocab_size)
        out_proj = quantized_weight_from_info(info, gf.get_tensor_data(info))
    return LlamaWeights(cfg, emb, layers, n1, n2, nf, out_proj)


_COLUMN_SHARDED = ("wq", "wk", "wv", "w_gate", "w_up", "out_proj")   # this rank's output columns; wo / w_down: its input rows


def shard_blocks(info: TensorInfo, raw: np.ndarray, key: str, rank: int, world: int):
    """This rank's slab of a Q8_0 / Q4_0 [K, N] tensor, cut at BLOCK level from the raw GGUF bytes (no int8 expansion).
    Blocks run along n inside one k row (flat index k * N + n), so a column slab [n0, n1) with n0, n1 multiples of 32 is
    the block range [n0 / 32, n1 / 32) of every row, and a row slab is one contiguous run.
    Returns (bytes, K_local, N_local); the same slabs as `llama.shard_weights` makes from the expanded form."""
    K, N = info.dims[0], info.dims[1]
    bb = info.type_.type_size
    blk = np.asarray(raw).reshape(K, N // 32, bb)
    if world == 1:
        return blk.reshape(-1), K, N
    if key in _COLUMN_SHARDED:
        w = N // world
        assert w % 32 == 0
        return np.ascontiguousarray(blk[:, rank * w // 32:(rank + 1) * w // 32]).reshape(-1), K, w
    r = K // world
    return np.ascontiguousarray(blk[rank * r:(rank + 1) * r]).reshape(-1), r, N


def upload_quantized_tensor(be, gf: GGUFFile, name: str, key: str = "", rank: int = 0, world: int = 1):
    """One Q8_0 / Q4_0 tensor (or this rank's slab of it), raw block bytes -> packed device weight
    (`zg_cuda_qweight_upload_gguf`)."""
    from ..backend import QuantizedWeight
    info = gf.get_tensor_info(name)
    if info is None:
        raise GGUFError(f"TensorNotFound: {name}")
    if info.n_dims < 2:
        raise GGUFError("UnsupportedShape")
    if not is_direct_quantized_matmul_type(info.type_):
        raise GGUFError("UnsupportedType")
    raw, K, N = shard_blocks(info, gf.get_tensor_data(info), key, rank, world)
    return QuantizedWeight.from_gguf_blocks(be, raw, int(info.type_), K, N)


def load_resident(be, gf: GGUFFile, cfg: Optional[LlamaConfig] = None, rank: int = 0, world: int = 1):
    """The device form of loadDirectQuantized: linears go tensor by tensor from the (memory-mapped) file into HBM and
    the program borrows them (`ZG_QWEIGHT_RESIDENT`).  With world > 1 every rank reads only its own slabs of the file
    (row-sharded model).  Returns (LlamaWeights, handles to free after the session)."""
    cfg = cfg or config_from_gguf(gf)
    check_shardable(cfg, world)
    shapes = linear_shapes(cfg)
    handles, layers = [], []

    def up(name, key, K, N):
        _check_linear(gf, name, K, N)
        h = upload_quantized_tensor(be, gf, name, key, rank, world)
        handles.append(h)
        return ResidentQuantizedWeight(h)

    for i in range(cfg.n_layers):
        layers.append({key: up(f"blk.{i}.{_LINEAR_NAMES[key]}", key, *shapes[key]) for key in LINEARS})
    emb, n1, n2, nf = _f32_params(gf, cfg)
    out_proj = None if cfg.tied_lm_head else up("output.weight", "out_proj", cfg.d_model, cfg.vocab_size)
    if world == 1:
        return LlamaWeights(cfg, emb, layers, n1, n2, nf, out_proj), handles
    V = cfg.vocab_size // world
    head_rows = np.ascontiguousarray(emb[rank * V:(rank + 1) * V]) if cfg.tied_lm_head else None
    return LlamaWeights(cfg, emb, layers, n1, n2, nf, out_proj, (rank, world), head_rows), handles


def write_llama_gguf(path: str, cfg: LlamaConfig, kind: str = "q8_0", seed: int = 0, embed_scale: float = 0.05, version: int = 3):
    """A random-init GGUF of `cfg`'s architecture in zgml's convention: linears Q8_0 / Q4_0 with
    `dims = [K, N]`, norms = 1 (f32), `token_embd.weight` f32 `[d_model, vocab]` ~ U(-embed_scale, embed_scale)."""
    from .llama import synthetic_gguf_blocks
    t = GGMLType.q8_0 if kind == "q8_0" else GGMLType.q4_0
    r = np.random.default_rng(seed)
    w = GGUFWriter(version)
    w.add_meta("general.architecture", MetaValueType.string, "llama")
    w.add_meta("general.name", MetaValueType.string, f"zgml-b200-synthetic-{kind}")
    for key, v in (("embedding_length", cfg.d_model), ("block_count", cfg.n_layers), ("attention.head_count", cfg.n_heads),
                   ("attention.head_count_kv", cfg.n_kv_heads), ("feed_forward_length", cfg.d_ff), ("context_length", cfg.max_seq_len),
                   ("vocab_size", cfg.vocab_size)):
        w.add_meta(f"llama.{key}", MetaValueType.uint32, v)
    w.add_meta("llama.rope.freq_base", MetaValueType.float32, cfg.rope_base)
    w.add_tensor("token_embd.weight", (cfg.d_model, cfg.vocab_size), GGMLType.f32,
                 r.uniform(-embed_scale, embed_scale, (cfg.vocab_size, cfg.d_model)).astype("<f4"))
    shapes = linear_shapes(cfg)
    ones = np.ones(cfg.d_model, "<f4")
    for i in range(cfg.n_layers):
        w.add_tensor(f"blk.{i}.attn_norm.weight", (cfg.d_model,), GGMLType.f32, ones)
        w.add_tensor(f"blk.{i}.ffn_norm.weight", (cfg.d_model,), GGMLType.f32, ones)
        for key in LINEARS:
            K, N = shapes[key]
            w.add_tensor(f"blk.{i}.{_LINEAR_NAMES[key]}", (K, N), t, synthetic_gguf_blocks(r, K, N, kind))
    w.add_tensor("output_norm.weight", (cfg.d_model,), GGMLType.f32, ones)
    if not cfg.tied_lm_head:
        w.add_tensor("output.weight", (cfg.d_model, cfg.vocab_size), t, synthetic_gguf_blocks(r, cfg.d_model, cfg.vocab_size, kind))
    w.write(path)
