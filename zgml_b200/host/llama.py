"""Host-side mirror of zgml's LLaMA device inference for the CUDA backend.

What the reference does in Zig, restated in Python over the same `Backend` interface:

  * `build_program(cfg, weights, token_len)` — the DeviceProgram that
    `DeviceInference.init` (src/device_inference.zig:61-238) lowers from the frozen plan
    of `LLaMA.forwardCachedMasked` (src/models/llama.zig:143-168) and
    `LLaMABlock.forwardCachedMasked` (src/models/llama_transformer.zig:192-253): per layer
    rmsnorm -> repeat(gamma) -> mul, q/k/v qmatmul, per-KV-head rope + KV-cache slice_assign,
    per-head rope + attention + slice_assign_rows into the concatenated buffer, o qmatmul,
    residual add, second norm, gate/up qmatmul, SiLU chain (src/nn.zig:38-44) as a fused
    elementwise op, down qmatmul, residual add; final norm; LM head (dense f32 matmul against
    the tied embedding, or a quantized output projection).  Field values follow the
    `*DeviceOp` helpers of src/device_inference.zig:464-672 (column-major tensors, ne[0] fastest).
  * `DeviceLlamaSession.step / prefill` — the per-token host work of
    `InferencePlan.execute` (src/llama_inference.zig:405-466): embedding row copy, causal
    mask rewrite, RoPE cos|sin patch, then `patchSliceAssignOffset(pos)`,
    `patchAttentionSeqKV(pos + n)` (src/device_inference.zig:240-256), `refreshProgram`,
    `executeProgram`.

Runs against any object with the Backend methods (`compile_program`, `refresh_program`,
`execute_program`, `free_program`): the CUDA backend in production, the oracle executor in
tests (tests/llama_reference.py).  No arithmetic on weights happens here.
"""
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

from .. import abi
from ..backend import DeviceOp, DeviceProgram, ProgramIO, QuantizedWeightUpload, ResidentQuantizedWeight


@dataclass(frozen=True)
class LlamaConfig:  # src/models/llama.zig:34-45
    vocab_size: int
    d_model: int
    n_layers: int
    n_heads: int
    n_kv_heads: int
    d_ff: int
    max_seq_len: int
    rms_norm_eps: float = 1e-5
    rope_base: float = 10000.0
    tied_lm_head: bool = True

    @property
    def d_head(self) -> int:
        return self.d_model // self.n_heads

    @property
    def kv_dim(self) -> int:
        return self.d_head * self.n_kv_heads


# shapes of BASELINE.json's configs (public HF configs; benchmarks/llama_smollm_bench.zig:31-42 for 135M)
SMOLLM_135M = LlamaConfig(49152, 576, 30, 9, 3, 1536, 2048, 1e-5, 1e4, True)
SMOLLM_1_7B = LlamaConfig(49152, 2048, 24, 32, 32, 8192, 2048, 1e-5, 1e4, True)
LLAMA3_8B = LlamaConfig(128256, 4096, 32, 32, 8, 14336, 2048, 1e-5, 5e5, False)
LLAMA3_70B = LlamaConfig(128256, 8192, 80, 64, 8, 28672, 2048, 1e-5, 5e5, False)

LINEARS = ("wq", "wk", "wv", "wo", "w_gate", "w_up", "w_down")


def linear_shapes(cfg: LlamaConfig) -> Dict[str, tuple]:
    """(K, N) of every per-layer linear: zgml rows = K (input), cols = N (output)."""
    return {"wq": (cfg.d_model, cfg.d_model), "wk": (cfg.d_model, cfg.kv_dim), "wv": (cfg.d_model, cfg.kv_dim),
            "wo": (cfg.d_model, cfg.d_model), "w_gate": (cfg.d_model, cfg.d_ff), "w_up": (cfg.d_model, cfg.d_ff),
            "w_down": (cfg.d_ff, cfg.d_model)}


@dataclass
class LlamaWeights:
    """Host form of a GGUF-direct quantized LLaMA (src/models/gguf_loader.zig:340-391): every 2-D
    linear as i8 + f32 block scales (what `quantizedWeightFromInfo` produces), norms and the token
    embedding as f32."""
    cfg: LlamaConfig
    token_embed: np.ndarray                     # f32 [vocab, d_model] (row v = embedding of token v)
    layers: List[Dict[str, QuantizedWeightUpload]]
    norm1: List[np.ndarray]
    norm2: List[np.ndarray]
    norm_f: np.ndarray
    out_proj: Optional[QuantizedWeightUpload] = None   # untied LM head [d_model, vocab]
    # Row-sharded form: `shard = (rank, world)`; every linear then holds only this rank's slab
    # (see `shard_weights`), `head_rows` the rank's vocab rows of the tied embedding.  token_embed stays whole
    # (host-side row lookup); anything indexable by token id works.
    shard: Optional[tuple] = None
    head_rows: Optional[np.ndarray] = None


def rope_tables(cfg: LlamaConfig):
    """RoPE.init (src/nn.zig:286-313): f32 `pos / pow(base, 2i/d)`, half-split pairing, both halves filled."""
    d = cfg.d_head
    pos = np.arange(cfg.max_seq_len, dtype=np.float32)[:, None]
    i = np.arange(d // 2, dtype=np.float32)[None, :]
    freq = (pos / np.power(np.float32(cfg.rope_base), (2 * i) / np.float32(d), dtype=np.float32)).astype(np.float32)
    cos = np.cos(freq).astype(np.float32)
    sin = np.sin(freq).astype(np.float32)
    return np.concatenate([cos, cos], axis=1), np.concatenate([sin, sin], axis=1)   # [max_seq, d] each


def check_shardable(cfg: LlamaConfig, world: int):
    """Slab boundaries must fall on multiples of 32 so that no quant block is split."""
    if world == 1:
        return
    if cfg.n_heads % world or cfg.n_kv_heads % world:
        raise ValueError(f"{cfg.n_heads}/{cfg.n_kv_heads} heads do not divide over {world} GPUs")
    widths = {"q": cfg.d_model // world, "kv": cfg.kv_dim // world, "ff": cfg.d_ff // world}
    if not cfg.tied_lm_head:
        widths["vocab"] = cfg.vocab_size // world
    if cfg.d_ff % world or cfg.vocab_size % world or any(v % 32 for v in widths.values()):
        raise ValueError(f"shard widths {widths} over {world} GPUs are not multiples of the 32-element quant block")


def slice_columns(qw: QuantizedWeightUpload, n0: int, n1: int) -> QuantizedWeightUpload:
    """Output-column slab [n0, n1) of a [K, N] weight.  Blocks run along n inside one k row (flat index k*N + n,
    src/quant.zig:525), so with N, n0, n1 multiples of the block size the slab keeps whole blocks: exact."""
    K, N, bs = qw.rows, qw.cols, qw.block_size
    assert N % bs == 0 and n0 % bs == 0 and n1 % bs == 0 and 0 <= n0 < n1 <= N
    data = np.ascontiguousarray(qw.data.reshape(K, N)[:, n0:n1]).ravel()
    scales = np.ascontiguousarray(qw.scales.reshape(K, N // bs)[:, n0 // bs:n1 // bs]).ravel()
    return QuantizedWeightUpload(data, scales, K, n1 - n0, bs)


def slice_rows(qw: QuantizedWeightUpload, k0: int, k1: int) -> QuantizedWeightUpload:
    """Input-row slab [k0, k1): a contiguous run of the flat array, block-aligned whenever N is."""
    K, N, bs = qw.rows, qw.cols, qw.block_size
    assert (k0 * N) % bs == 0 and (k1 * N) % bs == 0 and 0 <= k0 < k1 <= K
    return QuantizedWeightUpload(np.ascontiguousarray(qw.data[k0 * N:k1 * N]), np.ascontiguousarray(qw.scales[k0 * N // bs:k1 * N // bs]),
                                 k1 - k0, N, bs)


def shard_weights(w: LlamaWeights, rank: int, world: int) -> LlamaWeights:
    """This rank's slabs of a whole model (tests and small models; big models are generated per shard)."""
    cfg = w.cfg
    check_shardable(cfg, world)
    if world == 1:
        return w
    Dq, kvd, F, V = cfg.d_model // world, cfg.kv_dim // world, cfg.d_ff // world, cfg.vocab_size // world
    col = {"wq": Dq, "wk": kvd, "wv": kvd, "w_gate": F, "w_up": F}
    row = {"wo": Dq, "w_down": F}
    layers = []
    for L in w.layers:
        d = {n: slice_columns(L[n], rank * c, (rank + 1) * c) for n, c in col.items()}
        d.update({n: slice_rows(L[n], rank * r, (rank + 1) * r) for n, r in row.items()})
        layers.append(d)
    out_proj = None if w.out_proj is None else slice_columns(w.out_proj, rank * V, (rank + 1) * V)
    head_rows = np.ascontiguousarray(w.token_embed[rank * V:(rank + 1) * V]) if cfg.tied_lm_head else None
    return LlamaWeights(cfg, w.token_embed, layers, w.norm1, w.norm2, w.norm_f, out_proj, (rank, world), head_rows)


class _Buffers:
    def __init__(self):
        self.sizes: List[int] = []
        self.uploads: List[ProgramIO] = []

    def new(self, n: int, init: Optional[np.ndarray] = None) -> int:
        self.sizes.append(max(int(n), 1))
        idx = len(self.sizes) - 1
        if init is not None:
            self.uploads.append(ProgramIO(idx, np.ascontiguousarray(init, dtype=np.float32).ravel()))
        return idx


@dataclass
class LlamaProgram:
    program: DeviceProgram
    token_len: int
    buf_token_input: int
    buf_attn_mask: int
    buf_rope: List[int]
    buf_logits: int
    slice_assign_ops: List[int]
    attention_ops: List[int]
    n_qmatmul: int
    logits_offset: int = 0          # f32 element offset of the last position's logits in buf_logits


def build_program(cfg: LlamaConfig, w: LlamaWeights, token_len: int = 1) -> LlamaProgram:
    """Unsharded (w.shard is None): the reference lowering.  Sharded over `world` GPUs: q/k/v/gate/up
    hold this rank's output columns (its heads, its slice of d_ff), wo/down this rank's input rows followed by an
    all-reduce of the d_model partial sums, the LM head this rank's vocab slice followed by an all-gather of the
    last position's logits.  Activations are replicated; the KV cache is sharded by KV head."""
    rank, world = w.shard if w.shard else (0, 1)
    check_shardable(cfg, world)
    T, D, dh, S = token_len, cfg.d_model, cfg.d_head, cfg.max_seq_len
    hd = dh // 2
    n_rep = cfg.n_heads // cfg.n_kv_heads
    n_heads, n_kv = cfg.n_heads // world, cfg.n_kv_heads // world      # local heads
    Dq, kvd, F, V = n_heads * dh, n_kv * dh, cfg.d_ff // world, cfg.vocab_size // world   # local widths
    B = _Buffers()
    ops: List[abi.ZgOp] = []
    qws: List[QuantizedWeightUpload] = []
    sa_idx: List[int] = []
    at_idx: List[int] = []

    def qmm(dst, src, qw, K, N, src_rs=0):
        qws.append(qw)
        ops.append(DeviceOp.qmatmul(dst, src, len(qws) - 1, T, N, K, 0, src_rs if src_rs else K, 0, N))

    def rms(dst_norm, src, gamma_buf, bare, gamma_rep):
        # x.rmsNorm -> bare ; gamma.repeatLike(bare) ; bare.mul(rep)   (llama_transformer.zig:118-125)
        ops.append(DeviceOp.rmsnorm(bare, src, T, D, cfg.rms_norm_eps))
        ops.append(DeviceOp.repeat(gamma_rep, gamma_buf, D * T, (D, 1, 1, 1), (D, T, 1, 1), (1, D, D, D), (1, D, D * T, D * T)))
        ops.append(DeviceOp.elementwise("mul", dst_norm, bare, gamma_rep, D * T))

    token_input = B.new(D * T)
    attn_mask = B.new(S * T)
    ones_ff = B.new(F * T, np.ones(F * T, np.float32))      # `one.repeatLike(exp_neg)` of nn.silu
    # activations are reused by every layer (the reference's workspace planner aliases them too)
    bare, gamma_rep, norm = B.new(D * T), B.new(D * T), B.new(D * T)
    q_proj, k_proj, v_proj = B.new(Dq * T), B.new(kvd * T), B.new(kvd * T)
    k_rot = [B.new(dh * T) for _ in range(n_kv)]
    q_rot = [B.new(dh * T) for _ in range(n_heads)]
    attn_out = [B.new(dh * T) for _ in range(n_heads)]
    attn_buf, attn_proj, after_attn = B.new(Dq * T), B.new(D * T), B.new(D * T)
    gate, up, silu, hidden, down = B.new(F * T), B.new(F * T), B.new(F * T), B.new(F * T), B.new(D * T)
    x_bufs = [B.new(D * T), B.new(D * T)]                   # layer outputs ping-pong
    rope_bufs: List[int] = []

    x = token_input
    for li in range(cfg.n_layers):
        L = w.layers[li]
        g1, g2 = B.new(D, w.norm1[li]), B.new(D, w.norm2[li])
        k_cache, v_cache = B.new(dh * S * n_kv), B.new(dh * S * n_kv)
        cs = B.new(2 * dh * T)                              # packed cos|sin leaf of this layer (patched per step)
        rope_bufs.append(cs)

        rms(norm, x, g1, bare, gamma_rep)
        qmm(q_proj, norm, L["wq"], D, Dq)
        qmm(k_proj, norm, L["wk"], D, kvd)
        qmm(v_proj, norm, L["wv"], D, kvd)
        for kv in range(n_kv):
            base = kv * S * dh                              # head slab = contiguous column range of the consolidated cache
            ops.append(DeviceOp.rope(k_rot[kv], k_proj, cs, hd, T, kv * dh, 0, 0, 1, kvd, 2 * dh))
            sa_idx.append(len(ops))
            ops.append(DeviceOp.slice_assign(k_cache, k_rot[kv], dh, T, base, base, 1, dh, 0, 1, dh, dh))
            sa_idx.append(len(ops))
            ops.append(DeviceOp.slice_assign(v_cache, v_proj, dh, T, base, base, 1, dh, kv * dh, 1, kvd, dh))
        scale = float(np.float32(1.0) / np.sqrt(np.float32(dh)))
        for h in range(n_heads):
            kv = h // n_rep
            base = kv * S * dh
            ops.append(DeviceOp.rope(q_rot[h], q_proj, cs, hd, T, h * dh, 0, 0, 1, Dq, 2 * dh))
            at_idx.append(len(ops))
            ops.append(DeviceOp.attention(attn_out[h], q_rot[h], k_cache, v_cache, attn_mask, True, dh, T, S, scale,
                                          0, base, base, 0, 0, 1, dh, 1, dh, 1, dh, 1, S, 1, dh))
            # sliceAssignRows(attn_out, h * d_head): patch_stride 0 (device_inference.zig:695-701)
            ops.append(DeviceOp.slice_assign(attn_buf, attn_out[h], dh, T, 0, h * dh, 1, Dq, 0, 1, dh, 0))
        qmm(attn_proj, attn_buf, L["wo"], Dq, D)
        if world > 1:
            ops.append(DeviceOp.allreduce(attn_proj, D * T))
        ops.append(DeviceOp.elementwise("add", after_attn, x, attn_proj, D * T))
        rms(norm, after_attn, g2, bare, gamma_rep)
        qmm(gate, norm, L["w_gate"], D, F)
        qmm(up, norm, L["w_up"], D, F)
        # silu(gate) = gate * recip(exp(-gate) + 1), then * up
        ops.append(DeviceOp.fused_elementwise([("neg", False, 0, 0), ("exp", False, 0, 0), ("add", False, ones_ff, 0),
                                               ("recip", False, 0, 0), ("mul", True, gate, 0)], F * T, silu, gate))
        ops.append(DeviceOp.elementwise("mul", hidden, silu, up, F * T))
        qmm(down, hidden, L["w_down"], F, D)
        if world > 1:
            ops.append(DeviceOp.allreduce(down, D * T))
        out = x_bufs[li % 2]
        ops.append(DeviceOp.elementwise("add", out, after_attn, down, D * T))
        x = out

    gf = B.new(D, w.norm_f)
    rms(norm, x, gf, bare, gamma_rep)
    logits = B.new(V * T)
    if cfg.tied_lm_head:
        # x.matMul(false, token_embed, true) (models/llama.zig:162-165): dense f32, never quantized
        emb = B.new(V * D, w.head_rows if w.shard else w.token_embed)
        ops.append(DeviceOp.matmul(logits, norm, emb, T, V, D, D, 1, 1, D, dst_row_stride=V))
    else:
        qmm(logits, norm, w.out_proj, D, V)
    logits_off = (T - 1) * V                                # last-column logits (llama_inference.zig:463-465)
    if world > 1:
        full = B.new(cfg.vocab_size)
        ops.append(DeviceOp.allgather(full, logits, V, 0, logits_off))
        logits, logits_off = full, 0
    prog = DeviceProgram(ops, B.sizes, B.uploads, qws)
    return LlamaProgram(prog, T, token_input, attn_mask, rope_bufs, logits, sa_idx, at_idx, len(qws), logits_off)


class DeviceLlamaSession:
    """LlamaInferenceSession.step/prefill over a device backend (src/llama_inference.zig:681-727,
    benchmarks/llama_smollm_bench.zig:194-315 `runDeviceVariant`)."""

    def __init__(self, backend, cfg: LlamaConfig, weights: LlamaWeights, token_len: int = 1):
        self.be, self.cfg, self.w = backend, cfg, weights
        self.lp = build_program(cfg, weights, token_len)
        self.ops = self.lp.program.ops_array()              # the caller-owned, per-step mutated op array
        self.n_ops = len(self.lp.program.ops)
        self.handle = backend.compile_program(self.lp.program)
        if self.handle is None:
            raise RuntimeError("compile_program returned null")
        T, D, S, dh = token_len, cfg.d_model, cfg.max_seq_len, cfg.d_head
        self.token_input = np.zeros(D * T, np.float32)
        self.attn_mask = np.zeros(S * T, np.float32)
        self.rope_cs = np.zeros(2 * dh * T, np.float32)
        self.logits = np.zeros(cfg.vocab_size, np.float32)
        self.cos, self.sin = rope_tables(cfg)
        self.pos = 0
        self.inputs = [ProgramIO(self.lp.buf_token_input, self.token_input), ProgramIO(self.lp.buf_attn_mask, self.attn_mask)]
        self.inputs += [ProgramIO(b, self.rope_cs) for b in self.lp.buf_rope]
        self.outputs = [ProgramIO(self.lp.buf_logits, self.logits, offset=self.lp.logits_offset * 4)]
        self._init_patch_views()

    def reset(self):
        self.pos = 0

    def _patch_host_inputs(self, token_ids, pos):
        cfg, T = self.cfg, self.lp.token_len
        D, S, dh = cfg.d_model, cfg.max_seq_len, cfg.d_head
        assert len(token_ids) == T and pos + T <= S
        for i, tid in enumerate(token_ids):                 # 1. embedding rows
            self.token_input[i * D:(i + 1) * D] = self.w.token_embed[tid]
        for j in range(T):                                  # 2. causal mask, column j = position pos + j
            col = self.attn_mask[j * S:(j + 1) * S]
            col[:pos + j + 1] = 0.0
            col[pos + j + 1:] = -np.inf
        for j in range(T):                                  # 3. packed cos|sin for positions [pos, pos + T)
            self.rope_cs[j * 2 * dh:j * 2 * dh + dh] = self.cos[pos + j]
            self.rope_cs[j * 2 * dh + dh:(j + 1) * 2 * dh] = self.sin[pos + j]

    def _init_patch_views(self):
        """u32 view of the caller-owned op array + word indices of the per-step patched fields, so that
        patchSliceAssignOffset / patchAttentionSeqKV are two vectorised stores instead of a Python loop."""
        import ctypes as C
        self._ops_u32 = np.frombuffer((C.c_char * C.sizeof(self.ops)).from_buffer(self.ops), dtype=np.uint32)
        op_words = C.sizeof(abi.ZgOp) // 4
        u_off = abi.ZgOp.u.offset
        sa_t, at_t = type(self.ops[0].u.slice_assign), type(self.ops[0].u.attention)
        w_sa = (u_off + sa_t.dst_offset.offset) // 4
        w_at = (u_off + at_t.seq_kv.offset) // 4
        sa = [i for i in self.lp.slice_assign_ops if self.ops[i].u.slice_assign.patch_stride]
        self._sa_words = np.array([i * op_words + w_sa for i in sa], dtype=np.int64)
        self._sa_base = np.array([self.ops[i].u.slice_assign.dst_base_offset for i in sa], dtype=np.uint32)
        self._sa_stride = np.array([self.ops[i].u.slice_assign.patch_stride for i in sa], dtype=np.uint32)
        self._at_words = np.array([i * op_words + w_at for i in self.lp.attention_ops], dtype=np.int64)

    def _patch_ops(self, pos):
        self._ops_u32[self._sa_words] = self._sa_base + np.uint32(pos) * self._sa_stride   # patchSliceAssignOffset
        self._ops_u32[self._at_words] = np.uint32(pos + self.lp.token_len)                 # patchAttentionSeqKV(pos + T)

    def execute_at(self, token_ids, pos) -> np.ndarray:
        self._patch_host_inputs(token_ids, pos)
        self._patch_ops(pos)
        self.be.refresh_program(self.handle, _OpsView(self.ops, self.n_ops))
        self.be.execute_program(self.handle, self.inputs, self.outputs)
        return self.logits

    def step(self, token_id: int) -> np.ndarray:
        assert self.lp.token_len == 1
        out = self.execute_at([token_id], self.pos)
        self.pos += 1
        return out

    def close(self):
        if self.handle is not None:
            self.be.free_program(self.handle)
            self.handle = None


class _OpsView:
    """ctypes op array + length, accepted by CudaBackend.refresh_program without a copy."""

    def __init__(self, arr, n):
        self.arr, self.n = arr, n

    def __len__(self):
        return self.n


def synthetic_weights(cfg: LlamaConfig, kind: str = "q8_0", seed: int = 0, embed_scale: float = 0.05) -> LlamaWeights:
    """Random-init GGUF-direct weights in the reference's host form: per linear
    q ~ U{-127..127} (q8_0) / U{-8..7} (q4_0), one f16-representable scale per 32 flat elements sized so the
    dequantized weight is ~U(-b, b) with b = sqrt(6 / K) (kaimingUniform, src/nn.zig:91-105); norms = 1;
    token embedding ~U(-embed_scale, embed_scale)."""
    r = np.random.default_rng(seed)
    qmax = 127 if kind == "q8_0" else 7
    layers = []

    def make(K, N):
        n = K * N
        if kind == "q8_0":
            data = r.integers(-127, 128, n, dtype=np.int8)
        else:
            data = r.integers(-8, 8, n, dtype=np.int8)
        b = np.sqrt(6.0 / K)
        scales = (r.uniform(0.5, 1.0, n // 32) * (b / qmax)).astype(np.float16).astype(np.float32)
        return QuantizedWeightUpload(data, scales, K, N, 32)

    shapes = linear_shapes(cfg)
    for _ in range(cfg.n_layers):
        layers.append({name: make(*shapes[name]) for name in LINEARS})
    emb = r.uniform(-embed_scale, embed_scale, (cfg.vocab_size, cfg.d_model)).astype(np.float32)
    ones = [np.ones(cfg.d_model, np.float32) for _ in range(cfg.n_layers)]
    out_proj = None if cfg.tied_lm_head else make(cfg.d_model, cfg.vocab_size)
    return LlamaWeights(cfg, emb, layers, ones, [o.copy() for o in ones], np.ones(cfg.d_model, np.float32), out_proj)


class SyntheticEmbedding:
    """Token-embedding rows generated on demand (row v ~ U(-scale, scale), seeded by v): big-vocabulary synthetic
    models need no [vocab, d_model] f32 table on the host."""

    def __init__(self, vocab_size: int, d_model: int, seed: int = 0, scale: float = 0.05):
        self.shape, self.seed, self.scale = (vocab_size, d_model), seed, scale

    def __getitem__(self, tid: int) -> np.ndarray:
        return np.random.default_rng([self.seed, int(tid)]).uniform(-self.scale, self.scale, self.shape[1]).astype(np.float32)


GGML_Q4_0, GGML_Q8_0 = 2, 8


def synthetic_gguf_blocks(r: np.random.Generator, K: int, N: int, kind: str) -> np.ndarray:
    """Raw GGUF block bytes of a random [K, N] weight (src/gguf.zig:65-112: Q8_0 = f16 scale + 32 i8 = 34 B,
    Q4_0 = f16 scale + 16 nibble bytes = 18 B per 32 flat elements).  Everything comes from ONE stream of random
    bytes (this has to produce tens of GB for the 70B shape): quants are the bytes themselves, the f16 scale keeps
    its 10 random mantissa bits under a fixed exponent, i.e. uniform in [2^e, 2^(e+1)) with 2^(e+1) <= sqrt(6/K)/qmax,
    so the dequantized magnitude is ~ sqrt(6 / K) like kaimingUniform (src/nn.zig:91-105)."""
    nb = K * N // 32
    bb, qmax = (34, 127) if kind == "q8_0" else (18, 7)
    raw = r.integers(0, 1 << 64, (nb * bb + 7) // 8, dtype=np.uint64).view(np.uint8)[:nb * bb].reshape(nb, bb)
    if kind == "q8_0":
        q = raw[:, 2:].view(np.int8)
        np.maximum(q, -127, out=q)
    e = int(np.floor(np.log2(np.sqrt(6.0 / K) / qmax))) - 1 + 15          # biased f16 exponent (normal range)
    assert 1 <= e <= 30
    raw[:, 1] = (raw[:, 1] & 3) | np.uint8(e << 2)                          # sign 0 | exponent | top 2 mantissa bits
    return raw.ravel()


_MASK64 = (1 << 64) - 1


def _mix64(z: np.ndarray) -> np.ndarray:
    """splitmix64 finalizer on uint64 arrays (wrap-around arithmetic) — zg_mix64 in csrc/qweight.cu."""
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def tensor_id(layer: int, name: str) -> int:
    """Identity of a linear inside a synthetic model: what `zg_cuda_qweight_synth_gguf` hashes with the seed."""
    return 0xFFFF0000 if name == "out_proj" else layer * 16 + LINEARS.index(name)


def synth_gguf_blocks(seed: int, tid: int, kind: str, K_full: int, N_full: int, k0: int = 0, k1: Optional[int] = None,
                      n0: int = 0, n1: Optional[int] = None) -> np.ndarray:
    """Host twin of `zg_cuda_qweight_synth_gguf` (csrc/qweight.cu k_synth_gguf): the raw GGUF blocks of the slab
    [k0, k1) x [n0, n1) of the global [K_full, N_full] tensor `tid` of the synthetic model `seed`.  Every byte is a pure
    function of (seed, tensor, global block index), so shards of any world size and the unsharded host form are slices of
    the same weights."""
    k1 = K_full if k1 is None else k1
    n1 = N_full if n1 is None else n1
    assert N_full % 32 == 0 and n0 % 32 == 0 and n1 % 32 == 0 and 0 <= k0 < k1 <= K_full and 0 <= n0 < n1 <= N_full
    bb, words, qmax = (34, 5, 127.0) if kind == "q8_0" else (18, 3, 7.0)
    e = int(np.floor(np.log2(np.sqrt(6.0 / K_full) / qmax))) - 1 + 15
    e = min(max(e, 1), 30)
    with np.errstate(over="ignore"):
        key = _mix64(np.array([(seed * 0x9E3779B97F4A7C15 + tid) & _MASK64], dtype=np.uint64))[0]
        k = np.arange(k0, k1, dtype=np.uint64)[:, None]
        nb = np.arange(n0 // 32, n1 // 32, dtype=np.uint64)[None, :]
        gb = (k * np.uint64(N_full // 32) + nb).ravel()                       # global block index of every slab block
        w = _mix64(key + np.uint64(8) * gb[:, None] + np.arange(words, dtype=np.uint64)[None, :])
    raw = np.ascontiguousarray(w).view(np.uint8).reshape(len(gb), words * 8)[:, :bb].copy()   # little-endian words
    raw[:, 1] = (raw[:, 1] & 3) | np.uint8(e << 2)
    if kind == "q8_0":
        q = raw[:, 2:]
        q[q == 0x80] = 0x81
    return raw.ravel()


def synthetic_model_host(cfg: LlamaConfig, kind: str, seed: int, embed_scale: float = 0.05) -> LlamaWeights:
    """The WHOLE synthetic model of `synthetic_resident_shard` in the reference's host form (i8 + f32 scales per linear,
    `quantizedWeightFromInfo` applied to `synth_gguf_blocks`): what the CPU oracle runs against any sharded GPU run."""
    from .gguf import GGMLType, TensorInfo, quantized_weight_from_info
    t = GGMLType.q8_0 if kind == "q8_0" else GGMLType.q4_0
    shapes = linear_shapes(cfg)

    def make(li, name, K, N):
        raw = synth_gguf_blocks(seed, tensor_id(li, name), kind, K, N)
        return quantized_weight_from_info(TensorInfo(name, 2, (K, N, 1, 1), t, 0), raw)

    layers = [{n: make(li, n, *shapes[n]) for n in LINEARS} for li in range(cfg.n_layers)]
    ones = [np.ones(cfg.d_model, np.float32) for _ in range(cfg.n_layers)]
    if cfg.tied_lm_head:
        emb = np.random.default_rng([seed, 7]).uniform(-embed_scale, embed_scale, (cfg.vocab_size, cfg.d_model)).astype(np.float32)
        out_proj = None
    else:
        emb = SyntheticEmbedding(cfg.vocab_size, cfg.d_model, seed, embed_scale)
        out_proj = make(0, "out_proj", cfg.d_model, cfg.vocab_size)
    return LlamaWeights(cfg, emb, layers, ones, [o.copy() for o in ones], np.ones(cfg.d_model, np.float32), out_proj)


def synthetic_resident_shard(be, cfg: LlamaConfig, kind: str, seed: int, rank: int = 0, world: int = 1,
                             embed_scale: float = 0.05):
    """This rank's shard of a random-init GGUF model, generated tensor by tensor IN HBM (`zg_cuda_qweight_synth_gguf` +
    quantizedWeightFromInfo on device): no host copy at all, which is what makes Llama-3-70B-shape Q4_0 (39 GB of blocks)
    loadable in seconds next to 7 other ranks.  The model depends on `seed` only — every world size holds slices of the SAME
    weights (`synthetic_model_host` is their host form) — so logits and greedy tokens can be compared across world sizes
    and against the CPU oracle.  Returns (LlamaWeights of ResidentQuantizedWeight descriptors, the handles to free)."""
    from ..backend import QuantizedWeight
    check_shardable(cfg, world)
    ggml = GGML_Q8_0 if kind == "q8_0" else GGML_Q4_0
    D, kvd, F, V = cfg.d_model, cfg.kv_dim, cfg.d_ff, cfg.vocab_size
    full = linear_shapes(cfg)
    handles, layers = [], []

    def make(li, name):
        K, N = (D, V) if name == "out_proj" else full[name]
        if name in ("wo", "w_down"):                      # input-row slab
            k0, k1, n0, n1 = rank * (K // world), (rank + 1) * (K // world), 0, N
        else:                                             # output-column slab
            k0, k1, n0, n1 = 0, K, rank * (N // world), (rank + 1) * (N // world)
        h = QuantizedWeight.synth_gguf(be, seed, tensor_id(li, name), ggml, K, N, k0, k1, n0, n1)
        handles.append(h)
        return ResidentQuantizedWeight(h)

    for li in range(cfg.n_layers):
        layers.append({n: make(li, n) for n in LINEARS})
    ones = [np.ones(D, np.float32) for _ in range(cfg.n_layers)]
    out_proj, head_rows = None, None
    if cfg.tied_lm_head:
        emb = np.random.default_rng([seed, 7]).uniform(-embed_scale, embed_scale, (V, D)).astype(np.float32)
        head_rows = np.ascontiguousarray(emb[rank * (V // world):(rank + 1) * (V // world)])
    else:
        emb = SyntheticEmbedding(V, D, seed, embed_scale)
        out_proj = make(0, "out_proj")
    w = LlamaWeights(cfg, emb, layers, ones, [o.copy() for o in ones], np.ones(D, np.float32), out_proj,
                     (rank, world) if world > 1 else None, head_rows if world > 1 else None)
    return w, handles
