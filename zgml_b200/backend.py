"""Host-side mirror of zgml's backend interface over the C-ABI CUDA library.

`CudaBackend` has the six `Backend.VTable` slots of reference src/backend.zig:339-352
under the reference's own names (compile_program / refresh_program / execute_program /
free_program / get_runtime_profile / dense_matmul_f32) plus `supports_program`
(= DeviceProgram.isSupportedBy, src/backend.zig:277-297).  `DeviceOp.*` build ops with
the reference's field names and defaults; `QuantizedWeight` mirrors the
src/quant.zig:200-630 container for the GPU-resident packed form.

There is no CPU fallback: if the CUDA library is missing or no B200 is present every
entry point raises.
"""
import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import abi

_LIB = None
_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libzgml_cuda.so")


class BackendError(RuntimeError):
    pass


def library_path() -> str:
    return _LIB_PATH


def load_library():
    """dlopen libzgml_cuda.so and bind every symbol include/zgml_cuda.h declares."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_LIB_PATH):
        raise BackendError(
            f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback for the CUDA backend.")
    lib = C.CDLL(_LIB_PATH)
    for name, restype, argtypes in abi.SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _LIB = lib
    return lib


def last_error() -> str:
    return load_library().zg_cuda_last_error().decode()


# ── DeviceOp constructors (reference src/backend.zig:179-249) ──────────────────
class DeviceOp:
    @staticmethod
    def _mk(tag):
        op = abi.ZgOp()
        op.tag = tag
        return op

    @staticmethod
    def elementwise(op, dst, src0, src1, n, dst_offset=0, src0_offset=0, src1_offset=0):
        o = DeviceOp._mk(abi.OP_ELEMENTWISE)
        e = o.u.elementwise
        e.op = abi.EW_NAMES[op] if isinstance(op, str) else op
        e.dst, e.src0, e.src1, e.n = dst, src0, src1, n
        e.dst_offset, e.src0_offset, e.src1_offset = dst_offset, src0_offset, src1_offset
        return o

    @staticmethod
    def matmul(dst, a, b, M, N, K, a_row_stride, a_col_stride, b_row_stride, b_col_stride,
               a_offset=0, b_offset=0, dst_offset=0, dst_row_stride=None):
        o = DeviceOp._mk(abi.OP_MATMUL)
        m = o.u.matmul
        m.dst, m.a, m.b = dst, a, b
        g = m.geom
        g.M, g.N, g.K = M, N, K
        g.a_row_stride, g.a_col_stride, g.b_row_stride, g.b_col_stride = a_row_stride, a_col_stride, b_row_stride, b_col_stride
        g.a_offset, g.b_offset, g.dst_offset = a_offset, b_offset, dst_offset
        g.dst_row_stride = N if dst_row_stride is None else dst_row_stride
        return o

    @staticmethod
    def qmatmul(dst, input, weight_idx, M, N, K, input_offset=0, input_row_stride=0, dst_offset=0, dst_row_stride=0):
        o = DeviceOp._mk(abi.OP_QMATMUL)
        q = o.u.qmatmul
        q.dst, q.input, q.weight_idx, q.M, q.N, q.K = dst, input, weight_idx, M, N, K
        q.input_offset, q.input_row_stride, q.dst_offset, q.dst_row_stride = input_offset, input_row_stride, dst_offset, dst_row_stride
        return o

    @staticmethod
    def softmax(dst, src, rows, cols, src_offset=0, dst_offset=0):
        o = DeviceOp._mk(abi.OP_SOFTMAX)
        s = o.u.softmax
        s.dst, s.src, s.rows, s.cols, s.src_offset, s.dst_offset = dst, src, rows, cols, src_offset, dst_offset
        return o

    @staticmethod
    def layernorm(dst, src, rows, cols, eps=1e-5, src_offset=0, dst_offset=0):
        o = DeviceOp._mk(abi.OP_LAYERNORM)
        s = o.u.layernorm
        s.dst, s.src, s.rows, s.cols, s.eps, s.src_offset, s.dst_offset = dst, src, rows, cols, eps, src_offset, dst_offset
        return o

    @staticmethod
    def rmsnorm(dst, src, rows, cols, eps=1e-5, src_offset=0, dst_offset=0):
        o = DeviceOp._mk(abi.OP_RMSNORM)
        s = o.u.rmsnorm
        s.dst, s.src, s.rows, s.cols, s.eps, s.src_offset, s.dst_offset = dst, src, rows, cols, eps, src_offset, dst_offset
        return o

    @staticmethod
    def reduce(op, dst, src, n_out, reduce_size, src_offset=0, dst_offset=0):
        o = DeviceOp._mk(abi.OP_REDUCE)
        r = o.u.reduce
        r.op = abi.EW_NAMES[op] if isinstance(op, str) else op
        r.dst, r.src, r.n_out, r.reduce_size, r.src_offset, r.dst_offset = dst, src, n_out, reduce_size, src_offset, dst_offset
        return o

    @staticmethod
    def repeat(dst, src, n, src_ne, dst_ne, src_strides, dst_strides, src_offset=0, dst_offset=0):
        o = DeviceOp._mk(abi.OP_REPEAT)
        r = o.u.repeat
        r.dst, r.src, r.n = dst, src, n
        for i in range(4):
            r.src_ne[i], r.dst_ne[i], r.src_strides[i], r.dst_strides[i] = src_ne[i], dst_ne[i], src_strides[i], dst_strides[i]
        r.src_offset, r.dst_offset = src_offset, dst_offset
        return o

    @staticmethod
    def slice_assign(dst, src, rows, cols, dst_base_offset, dst_offset, dst_row_stride, dst_col_stride,
                     src_offset, src_row_stride, src_col_stride, patch_stride):
        o = DeviceOp._mk(abi.OP_SLICE_ASSIGN)
        s = o.u.slice_assign
        s.dst, s.src, s.rows, s.cols = dst, src, rows, cols
        s.dst_base_offset, s.dst_offset, s.dst_row_stride, s.dst_col_stride = dst_base_offset, dst_offset, dst_row_stride, dst_col_stride
        s.src_offset, s.src_row_stride, s.src_col_stride, s.patch_stride = src_offset, src_row_stride, src_col_stride, patch_stride
        return o

    @staticmethod
    def rope(dst, src, cos_sin, half_d, seq_len, src_off, cs_off, dst_off, src_rs, src_cs, cs_cs):
        o = DeviceOp._mk(abi.OP_ROPE)
        r = o.u.rope
        r.dst, r.src, r.cos_sin, r.half_d, r.seq_len = dst, src, cos_sin, half_d, seq_len
        r.src_off, r.cs_off, r.dst_off, r.src_rs, r.src_cs, r.cs_cs = src_off, cs_off, dst_off, src_rs, src_cs, cs_cs
        return o

    @staticmethod
    def attention(dst, q, k, v, mask, has_mask, d_head, seq_q, seq_kv, scale, q_off, k_off, v_off, mask_off,
                  dst_off, q_rs, q_cs, k_rs, k_cs, v_rs, v_cs, mask_rs, mask_cs, dst_rs, dst_cs):
        o = DeviceOp._mk(abi.OP_ATTENTION)
        a = o.u.attention
        a.dst, a.q, a.k, a.v, a.mask, a.has_mask = dst, q, k, v, mask, int(bool(has_mask))
        a.d_head, a.seq_q, a.seq_kv, a.scale = d_head, seq_q, seq_kv, scale
        a.q_off, a.k_off, a.v_off, a.mask_off, a.dst_off = q_off, k_off, v_off, mask_off, dst_off
        a.q_rs, a.q_cs, a.k_rs, a.k_cs, a.v_rs, a.v_cs = q_rs, q_cs, k_rs, k_cs, v_rs, v_cs
        a.mask_rs, a.mask_cs, a.dst_rs, a.dst_cs = mask_rs, mask_cs, dst_rs, dst_cs
        return o

    @staticmethod
    def fused_elementwise(steps, n, dst, src, dst_offset=0, src_offset=0):
        """steps: sequence of (op, is_swapped, secondary_buf, secondary_offset)."""
        o = DeviceOp._mk(abi.OP_FUSED_ELEMENTWISE)
        arr = (abi.ZgFusedEwStep * max(len(steps), 1))()
        for i, (op, sw, sb, so) in enumerate(steps):
            arr[i].op = abi.EW_NAMES[op] if isinstance(op, str) else op
            arr[i].is_swapped, arr[i].secondary_buf, arr[i].secondary_offset = int(bool(sw)), sb, so
        f = o.u.fused_elementwise
        f.steps = C.cast(arr, C.POINTER(abi.ZgFusedEwStep))
        f.n_steps = len(steps)
        f.n, f.dst, f.src, f.dst_offset, f.src_offset = n, dst, src, dst_offset, src_offset
        o._steps_keepalive = arr  # the caller owns `steps` (src/device_inference.zig:291-298)
        return o

    # Multi-GPU extensions (no zgml counterpart).  One NCCL collective each.
    @staticmethod
    def allreduce(buf, n, offset=0):
        o = DeviceOp._mk(abi.OP_ALLREDUCE)
        o.u.allreduce.buf, o.u.allreduce.offset, o.u.allreduce.n = buf, offset, n
        return o

    @staticmethod
    def allgather(dst, src, n, dst_offset=0, src_offset=0):
        o = DeviceOp._mk(abi.OP_ALLGATHER)
        g = o.u.allgather
        g.dst, g.src, g.n, g.dst_offset, g.src_offset = dst, src, n, dst_offset, src_offset
        return o


@dataclass
class ProgramIO:  # reference src/backend.zig:252-257 (offset/size in bytes)
    buf_idx: int
    host: np.ndarray
    offset: int = 0
    size: Optional[int] = None

    def to_c(self) -> abi.ZgIO:
        io = abi.ZgIO()
        io.buf_idx, io.offset = self.buf_idx, self.offset
        io.host_ptr = self.host.ctypes.data
        io.size = self.host.nbytes if self.size is None else self.size
        return io


@dataclass
class QuantizedWeightUpload:  # reference src/backend.zig:260-266
    data: np.ndarray    # int8 [rows*cols], row-major [K=rows, N=cols]
    scales: np.ndarray  # float32 [ceil(rows*cols / block_size)]
    rows: int
    cols: int
    block_size: int

    def to_c(self) -> abi.ZgQWeight:
        assert self.data.dtype == np.int8 and self.scales.dtype == np.float32
        assert self.data.flags.c_contiguous and self.scales.flags.c_contiguous
        q = abi.ZgQWeight()
        q.data, q.n_data = self.data.ctypes.data, self.data.size
        q.scales, q.n_scales = self.scales.ctypes.data, self.scales.size
        q.rows, q.cols, q.block_size = self.rows, self.cols, self.block_size
        return q


@dataclass
class ResidentQuantizedWeight:
    """Descriptor of a weight already packed in HBM (ZG_QWEIGHT_RESIDENT, include/zgml_cuda.h): the program
    borrows `weight` (a `QuantizedWeight`); the caller keeps it alive and frees it after free_program."""
    weight: "QuantizedWeight"

    @property
    def rows(self):
        return self.weight.rows

    @property
    def cols(self):
        return self.weight.cols

    block_size = abi.QWEIGHT_RESIDENT

    def to_c(self) -> abi.ZgQWeight:
        q = abi.ZgQWeight()
        q.data, q.rows, q.cols, q.block_size = self.weight.ptr, self.weight.rows, self.weight.cols, abi.QWEIGHT_RESIDENT
        return q


@dataclass
class DeviceProgram:  # reference src/backend.zig:270-275
    ops: List[abi.ZgOp]
    buffer_sizes: Sequence[int]
    initial_uploads: List[ProgramIO] = field(default_factory=list)
    qweights: List[QuantizedWeightUpload] = field(default_factory=list)

    @property
    def n_buffers(self) -> int:
        return len(self.buffer_sizes)

    def ops_array(self):
        arr = (abi.ZgOp * max(len(self.ops), 1))()
        for i, o in enumerate(self.ops):
            arr[i] = o
        return arr

    def to_c(self):
        """Returns (ZgProgram, keepalive)."""
        keep = {}
        keep["ops"] = self.ops_array()
        keep["sizes"] = (C.c_size_t * max(self.n_buffers, 1))(*[int(s) for s in self.buffer_sizes])
        keep["ups"] = (abi.ZgIO * max(len(self.initial_uploads), 1))()
        for i, io in enumerate(self.initial_uploads):
            keep["ups"][i] = io.to_c()
        keep["qws"] = (abi.ZgQWeight * max(len(self.qweights), 1))()
        for i, qw in enumerate(self.qweights):
            keep["qws"][i] = qw.to_c()
        p = abi.ZgProgram()
        p.ops, p.n_ops = C.cast(keep["ops"], C.POINTER(abi.ZgOp)), len(self.ops)
        p.n_buffers = self.n_buffers
        p.buffer_sizes = C.cast(keep["sizes"], C.POINTER(C.c_size_t))
        p.initial_uploads, p.n_uploads = C.cast(keep["ups"], C.POINTER(abi.ZgIO)), len(self.initial_uploads)
        p.qweights, p.n_qweights = C.cast(keep["qws"], C.POINTER(abi.ZgQWeight)), len(self.qweights)
        return p, keep


def _io_array(ios: Sequence[ProgramIO]):
    arr = (abi.ZgIO * max(len(ios), 1))()
    for i, io in enumerate(ios):
        arr[i] = io.to_c()
    return arr


class CompiledHandle:
    def __init__(self, ptr, n_ops):
        self.ptr = ptr
        self.n_ops = n_ops


class CudaBackend:
    """The `cuda` sibling of CpuBackend / MetalBackend (src/backend/cpu.zig:9-20)."""

    name_str = "cuda"
    device_type = "cuda"  # Device.cuda, src/backend.zig:11

    def __init__(self, device_ordinal: int = 0):
        self.lib = load_library()
        self.ctx = self.lib.zg_cuda_create(device_ordinal)
        if not self.ctx:
            raise BackendError(f"zg_cuda_create({device_ordinal}) failed: {last_error()}")
        caps = abi.ZgCapabilities()
        self.lib.zg_cuda_capabilities(C.byref(caps))
        self.capabilities = caps
        self.rank, self.world, self.comm_ready = 0, 1, False

    # Multi-GPU: one process per GPU, one NCCL communicator per backend (csrc/comm.cu)
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        if self.lib.zg_cuda_comm_unique_id(buf) != 0:
            raise BackendError(f"comm_unique_id failed: {last_error()}")
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        if self.lib.zg_cuda_comm_init(self.ctx, buf, rank, world) != 0:
            raise BackendError(f"comm_init failed: {last_error()}")
        self.rank, self.world, self.comm_ready = rank, world, True

    def comm_mode(self) -> str:
        return {0: "none", 1: "nccl", 2: "nvlink-peer+nccl"}[self.lib.zg_cuda_comm_mode(self.ctx)]

    def comm_init_torch(self):
        """Exchange the NCCL id over an initialised torch.distributed group (any backend) and join."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [self.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        self.comm_init(box[0], rank, world)

    def close(self):
        if self.ctx:
            self.lib.zg_cuda_destroy(self.ctx)
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # DeviceProgram.isSupportedBy, src/backend.zig:277-297 (host-side check)
    def supports_program(self, program: DeviceProgram) -> bool:
        caps = self.capabilities
        if not caps.compiled_programs:
            return False
        nb = program.n_buffers
        for op in program.ops:
            t = op.tag
            if t == abi.OP_QMATMUL:
                q = op.u.qmatmul
                if not caps.qmatmul or q.dst >= nb or q.input >= nb:
                    return False
                if q.weight_idx >= len(program.qweights):
                    return False
                qw = program.qweights[q.weight_idx]
                if qw.block_size == 0 or qw.rows != q.K or qw.cols != q.N:
                    return False
                if qw.block_size == abi.QWEIGHT_RESIDENT:
                    continue
                n_elems = q.K * q.N
                n_blocks = (n_elems + qw.block_size - 1) // qw.block_size
                if qw.data.size < n_elems or qw.scales.size < n_blocks:
                    return False
            elif t == abi.OP_ATTENTION:
                a = op.u.attention
                if not caps.attention_supported or a.d_head > caps.attention_max_d_head:
                    return False
                if max(a.dst, a.q, a.k, a.v, a.mask) >= nb:
                    return False
            elif t == abi.OP_REDUCE:
                if op.u.reduce.op not in (abi.EW_SUM, abi.EW_MAX):
                    return False
            elif t == abi.OP_ELEMENTWISE:
                if not (abi.EW_ADD <= op.u.elementwise.op <= abi.EW_GELU):
                    return False
            elif t == abi.OP_FUSED_ELEMENTWISE:
                f = op.u.fused_elementwise
                if not caps.fused_elementwise:
                    return False
                for i in range(f.n_steps):
                    if not (abi.EW_ADD <= f.steps[i].op <= abi.EW_GELU):
                        return False
            elif t in (abi.OP_ALLREDUCE, abi.OP_ALLGATHER):
                pass   # world 1 (no comm_init): identity / local copy
            elif t >= abi.OP_COUNT:
                return False
        return True

    def dense_matmul_f32(self, spec=None) -> bool:
        return bool(self.lib.zg_cuda_dense_matmul_f32(self.ctx, None, None, None, None))

    def compile_program(self, program: DeviceProgram) -> Optional[CompiledHandle]:
        if not self.supports_program(program):  # Backend.compileProgram, src/backend.zig:354-357
            return None
        cprog, keep = program.to_c()
        ptr = self.lib.zg_cuda_compile(self.ctx, C.byref(cprog))
        del keep  # backend copied everything it needs (weights, uploads, ops)
        if not ptr:
            return None
        return CompiledHandle(ptr, len(program.ops))

    def refresh_program(self, handle: CompiledHandle, ops) -> None:
        if isinstance(ops, (list, tuple)):
            arr = (abi.ZgOp * max(len(ops), 1))()
            for i, o in enumerate(ops):
                arr[i] = o
            n = len(ops)
        elif hasattr(ops, "arr"):   # (ctypes array, length) view: no copy, like the reference's caller-owned slice
            arr, n = ops.arr, len(ops)
        else:
            arr, n = ops, len(ops)
        self.lib.zg_cuda_refresh(self.ctx, handle.ptr, C.cast(arr, C.POINTER(abi.ZgOp)), n)

    def execute_program(self, handle: CompiledHandle, inputs: Sequence[ProgramIO], outputs: Sequence[ProgramIO]) -> None:
        ia, oa = _io_array(inputs), _io_array(outputs)
        self.lib.zg_cuda_execute(self.ctx, handle.ptr, ia, len(inputs), oa, len(outputs))

    def free_program(self, handle: CompiledHandle) -> None:
        if handle.ptr:
            self.lib.zg_cuda_free(self.ctx, handle.ptr)
            handle.ptr = None

    def get_runtime_profile(self, handle: CompiledHandle):
        p = self.lib.zg_cuda_profile(self.ctx, handle.ptr)
        return p.contents if p else None

    # extensions
    def set_profiling(self, enabled: bool):
        self.lib.zg_cuda_set_profiling(self.ctx, int(enabled))

    def set_graph_mode(self, enabled: bool):
        self.lib.zg_cuda_set_graph_mode(self.ctx, int(enabled))

    def set_stream(self, cuda_stream: int):
        self.lib.zg_cuda_set_stream(self.ctx, cuda_stream)

    def sync(self):
        self.lib.zg_cuda_sync(self.ctx)

    def launch_count(self) -> int:
        return int(self.lib.zg_cuda_launch_count())

    def quantize_kv(self, handle: CompiledHandle, block_size: int = 32, int8_query: bool = False) -> None:
        """LlamaInferenceSession.quantizeKV (src/llama_inference.zig:648-679) for a compiled program: Q8 caches replace the f32 KV
        buffers, cache writes run storeColumn, attention runs attentionQuantized."""
        if self.lib.zg_cuda_program_quantize_kv(self.ctx, handle.ptr, block_size, int(int8_query)) != 0:
            raise BackendError(f"quantize_kv failed: {last_error()}")

    def promote_dense_weights(self, handle: CompiledHandle, fmt: str = "f16") -> int:
        """Format hint for dense matmul operands (include/zgml_cuda.h zg_cuda_program_promote_dense; precedent
        src/backend/wgpu.zig:1068-1106): single-row matmuls against a constant k-contiguous operand (the tied LM head) stream a
        16-bit copy (f16: 11 significant bits, bf16: 8) and recompute every possible argmax from the f32 original.  Returns the number of ops promoted."""
        if fmt not in ("bf16", "f16"):
            raise BackendError("promote_dense_weights: format must be 'f16' or 'bf16'")
        n = self.lib.zg_cuda_program_promote_dense(self.ctx, handle.ptr, 1 if fmt == "bf16" else 2)
        if n < 0:
            raise BackendError(f"promote_dense_weights failed: {last_error()}")
        return int(n)

    def program_stats(self, handle: CompiledHandle) -> dict:
        """Schedule facts of a compiled program: kernels per execution, DeviceOps / layers inside the fused decode kernel."""
        f = self.lib.zg_cuda_program_stats
        return {"kernels": int(f(handle.ptr, 0)), "fused_decode_ops": int(f(handle.ptr, 1)), "fused_decode_layers": int(f(handle.ptr, 2)),
                "streamed_matvec_launches": int(f(handle.ptr, 3)), "dense_bytes_saved_per_execution": int(f(handle.ptr, 4))}


class QuantizedWeight:
    """GPU-resident packed QuantizedWeight (src/quant.zig:200-212): rows = K, cols = N."""

    def __init__(self, be: CudaBackend, ptr, rows, cols, block_size):
        self.be, self.ptr, self.rows, self.cols, self.block_size = be, ptr, rows, cols, block_size

    @classmethod
    def upload(cls, be: CudaBackend, data: np.ndarray, scales: np.ndarray, rows: int, cols: int,
               block_size: int = 32, fmt_hint: int = abi.QFMT_AUTO) -> "QuantizedWeight":
        up = QuantizedWeightUpload(np.ascontiguousarray(data, dtype=np.int8).ravel(),
                                   np.ascontiguousarray(scales, dtype=np.float32).ravel(), rows, cols, block_size)
        c = up.to_c()
        ptr = be.lib.zg_cuda_qweight_upload(be.ctx, C.byref(c), fmt_hint)
        if not ptr:
            raise BackendError(f"qweight upload failed: {last_error()}")
        return cls(be, ptr, rows, cols, block_size)

    @classmethod
    def from_slice(cls, be: CudaBackend, weights: np.ndarray, rows: int, cols: int, block_size: int = 32, return_flat: bool = False):
        """QuantizedWeight.fromSlice (src/quant.zig:216-256) on device.  `return_flat`: also the reference's flat
        (data i8, scales f32) form, for bit-exact comparison."""
        w = np.ascontiguousarray(weights, dtype=np.float32).ravel()
        assert w.size == rows * cols
        data = np.zeros(w.size, np.int8) if return_flat else None
        scales = np.zeros((w.size + block_size - 1) // block_size, np.float32) if return_flat else None
        ptr = be.lib.zg_cuda_qweight_from_f32(be.ctx, w.ctypes.data, rows, cols, block_size,
                                              data.ctypes.data if return_flat else None, scales.ctypes.data if return_flat else None)
        if not ptr:
            raise BackendError(f"qweight from_slice failed: {last_error()}")
        qw = cls(be, ptr, rows, cols, block_size)
        return (qw, data, scales) if return_flat else qw

    @classmethod
    def from_gguf_blocks(cls, be: CudaBackend, raw: np.ndarray, ggml_type: int, rows: int, cols: int) -> "QuantizedWeight":
        """quantizedWeightFromInfo (src/models/gguf_loader.zig:99-154) on device."""
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        ptr = be.lib.zg_cuda_qweight_upload_gguf(be.ctx, raw.ctypes.data, raw.nbytes, ggml_type, rows, cols)
        if not ptr:
            raise BackendError(f"gguf qweight upload failed: {last_error()}")
        return cls(be, ptr, rows, cols, 32)

    @classmethod
    def synth_gguf(cls, be: CudaBackend, seed: int, tensor_id: int, ggml_type: int, rows_full: int, cols_full: int,
                   k0: int, k1: int, n0: int, n1: int) -> "QuantizedWeight":
        """Random-init GGUF blocks of the slab [k0, k1) x [n0, n1) of a global tensor, generated in HBM
        (zg_cuda_qweight_synth_gguf): world-size independent, reproduced on the host by host/llama.py::synth_gguf_blocks."""
        ptr = be.lib.zg_cuda_qweight_synth_gguf(be.ctx, seed, tensor_id, ggml_type, rows_full, cols_full, k0, k1, n0, n1)
        if not ptr:
            raise BackendError(f"synthetic qweight failed: {last_error()}")
        return cls(be, ptr, k1 - k0, n1 - n0, 32)

    @property
    def format(self) -> int:
        return self.be.lib.zg_cuda_qweight_format(self.ptr)

    @property
    def device_bytes(self) -> int:
        return int(self.be.lib.zg_cuda_qweight_device_bytes(self.ptr))

    def dequantize_to(self) -> np.ndarray:  # src/quant.zig:594-618
        out = np.empty(self.rows * self.cols, dtype=np.float32)
        if self.be.lib.zg_cuda_qweight_dequantize(self.be.ctx, self.ptr, out.ctypes.data) != 0:
            raise BackendError(f"dequantize failed: {last_error()}")
        return out.reshape(self.rows, self.cols)

    def matmul(self, input: np.ndarray, M: int) -> np.ndarray:  # src/quant.zig:475-578 via host buffers
        x = np.ascontiguousarray(input, dtype=np.float32).reshape(M, self.rows)
        out = np.empty((M, self.cols), dtype=np.float32)
        if self.be.lib.zg_cuda_qmatmul_host(self.be.ctx, self.ptr, x.ctypes.data, out.ctypes.data, M) != 0:
            raise BackendError(f"qmatmul failed: {last_error()}")
        return out

    def matmul_bias(self, input: np.ndarray, bias: np.ndarray, M: int) -> np.ndarray:  # src/quant.zig:581-589
        x = np.ascontiguousarray(input, dtype=np.float32).reshape(M, self.rows)
        b = np.ascontiguousarray(bias, dtype=np.float32).ravel()
        assert b.size == self.cols
        out = np.empty((M, self.cols), dtype=np.float32)
        if self.be.lib.zg_cuda_qmatmul_bias_host(self.be.ctx, self.ptr, x.ctypes.data, b.ctypes.data, out.ctypes.data, M) != 0:
            raise BackendError(f"qmatmul_bias failed: {last_error()}")
        return out

    def prepare_transposed(self, return_host: bool = False):  # src/quant.zig:274-317
        """Build the W8A8 form on device ([N, K] int8, K-aligned blocks).  `return_host`: also (t_data, t_scales) copies,
        bit-identical to the reference's."""
        bpr = (self.rows + self.block_size - 1) // self.block_size
        t_data = np.empty(self.cols * self.rows, np.int8) if return_host else None
        t_scales = np.empty(self.cols * bpr, np.float32) if return_host else None
        if self.be.lib.zg_cuda_qweight_prepare_transposed(self.be.ctx, self.ptr, t_data.ctypes.data if return_host else None,
                                                          t_scales.ctypes.data if return_host else None) != 0:
            raise BackendError(f"prepare_transposed failed: {last_error()}")
        return (t_data, t_scales) if return_host else None

    def gemv(self, input: np.ndarray) -> np.ndarray:  # src/quant.zig:443-459 (quantizeInput + gemvRange), M = 1
        x = np.ascontiguousarray(input, dtype=np.float32).ravel()
        assert x.size == self.rows
        out = np.empty(self.cols, dtype=np.float32)
        if self.be.lib.zg_cuda_gemv_w8a8_host(self.be.ctx, self.ptr, x.ctypes.data, out.ctypes.data) != 0:
            raise BackendError(f"gemv failed: {last_error()}")
        return out

    def gemv_device(self, d_input: int, d_dst: int):
        if self.be.lib.zg_cuda_gemv_w8a8_device(self.be.ctx, self.ptr, d_input, d_dst) != 0:
            raise BackendError(f"gemv_device failed: {last_error()}")

    def matmul_device(self, d_input: int, d_dst: int, M: int, input_row_stride: int = 0, dst_row_stride: int = 0):
        if self.be.lib.zg_cuda_qmatmul_device(self.be.ctx, self.ptr, d_input, d_dst, M, input_row_stride, dst_row_stride) != 0:
            raise BackendError(f"qmatmul_device failed: {last_error()}")

    def free(self):
        if self.ptr:
            self.be.lib.zg_cuda_qweight_free(self.be.ctx, self.ptr)
            self.ptr = None


def quantize_input(be: CudaBackend, input: np.ndarray, block_size: int = 32):  # src/quant.zig:320-341
    """quantizeInput on device: (int8 [K], f32 scales [ceil(K / block_size)]), bit-identical to the reference's."""
    x = np.ascontiguousarray(input, dtype=np.float32).ravel()
    q = np.empty(x.size, np.int8)
    scales = np.empty((x.size + block_size - 1) // block_size, np.float32)
    if be.lib.zg_cuda_quantize_input_host(be.ctx, x.ctypes.data, x.size, block_size, q.ctypes.data, scales.ctypes.data) != 0:
        raise BackendError(f"quantize_input failed: {last_error()}")
    return q, scales


class QuantizedKVCache:
    """GPU-resident QuantizedKVCache (src/quant.zig:646-761): column-major Q8, one column per kv position."""

    def __init__(self, be: CudaBackend, d_head: int, n_cols: int, block_size: int = 32):  # init, src/quant.zig:658-678
        self.be, self.d_head, self.n_cols, self.block_size = be, d_head, n_cols, block_size
        self.blocks_per_col = d_head // block_size if block_size else 0
        self.ptr = be.lib.zg_cuda_kvcache_create(be.ctx, d_head, n_cols, block_size)
        if not self.ptr:
            raise BackendError(f"kvcache create failed: {last_error()}")

    def clear(self):  # src/quant.zig:683-686
        if self.be.lib.zg_cuda_kvcache_clear(self.be.ctx, self.ptr) != 0:
            raise BackendError(f"kvcache clear failed: {last_error()}")

    def store_column(self, col_idx: int, src: np.ndarray):  # src/quant.zig:689-701
        self.store_columns(col_idx, np.ascontiguousarray(src, dtype=np.float32).reshape(1, self.d_head))

    def store_columns(self, col_start: int, src: np.ndarray):  # n_write consecutive columns, src/llama_inference.zig:336-348
        x = np.ascontiguousarray(src, dtype=np.float32).reshape(-1, self.d_head)
        if self.be.lib.zg_cuda_kvcache_store_host(self.be.ctx, self.ptr, col_start, x.shape[0], x.ctypes.data) != 0:
            raise BackendError(f"kvcache store failed: {last_error()}")

    def download(self):
        q = np.empty(self.d_head * self.n_cols, np.int8)
        s = np.empty(self.blocks_per_col * self.n_cols, np.float32)
        if self.be.lib.zg_cuda_kvcache_download(self.be.ctx, self.ptr, q.ctypes.data, s.ctypes.data) != 0:
            raise BackendError(f"kvcache download failed: {last_error()}")
        return q, s

    def dequant_column(self, col_idx: int) -> np.ndarray:  # src/quant.zig:704-716 (host arithmetic on the downloaded column)
        q, s = self.download()
        d, bs = self.d_head, self.block_size
        col = q[col_idx * d:(col_idx + 1) * d].astype(np.float32).reshape(-1, bs)
        return (col * s[col_idx * self.blocks_per_col:(col_idx + 1) * self.blocks_per_col, None]).ravel()

    def free(self):
        if self.ptr:
            self.be.lib.zg_cuda_kvcache_free(self.be.ctx, self.ptr)
            self.ptr = None


def attention_quantized(be: CudaBackend, q: np.ndarray, seq_q: int, k_cache: QuantizedKVCache, k_col_start: int,
                        v_cache: QuantizedKVCache, v_col_start: int, seq_kv: int, scale: float, mask=None, mask_row_stride: int = 0,
                        mask_col_stride: int = 0, int8_query: bool = True, q_col_stride=None, dst_col_stride=None) -> np.ndarray:
    """attentionQuantized (src/quant.zig:924-1091) through host buffers; same argument meaning as the reference."""
    d = k_cache.d_head
    q_cs = d if q_col_stride is None else q_col_stride
    d_cs = d if dst_col_stride is None else dst_col_stride
    qh = np.ascontiguousarray(q, dtype=np.float32).ravel()
    dst = np.zeros(max(1, (seq_q - 1) * d_cs + d), np.float32)
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.float32).ravel()
    rc = be.lib.zg_cuda_attention_quantized_host(be.ctx, dst.ctypes.data, d_cs, qh.ctypes.data, q_cs, d, seq_q, k_cache.ptr, k_col_start,
                                                 v_cache.ptr, v_col_start, seq_kv, None if m is None else m.ctypes.data,
                                                 mask_row_stride, mask_col_stride, float(scale), 1 if int8_query else 0)
    if rc != 0:
        raise BackendError(f"attention_quantized failed: {last_error()}")
    return dst
