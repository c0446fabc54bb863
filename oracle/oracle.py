"""ctypes wrapper over oracle/zgml_oracle.c — TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of zgml's quantized path (see the header of
zgml_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module; the product (zgml_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from zgml_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "zgml_oracle.c")
_LIBS = {}


def build(native: bool = False, force: bool = False) -> str:
    """gcc -O3 -ffp-contract=off; `native` = -march=native (CPU baseline on the box it runs on)."""
    out = os.path.join(_HERE, "_build", "libzgml_oracle_native.so" if native else "libzgml_oracle.so")
    hdr = os.path.join(_HERE, "..", "include", "zgml_cuda.h")
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(_SRC), os.path.getmtime(hdr)):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    march = "-march=native" if native else "-march=x86-64-v3"
    subprocess.run(["gcc", "-O3", march, "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                    "-fvisibility=hidden", "-Wall", "-o", out, _SRC, "-lm", "-lpthread"], check=True)
    return out


def lib(native: bool = False):
    if native in _LIBS:
        return _LIBS[native]
    path = build(native=native)
    L = C.CDLL(path)
    vp, sz = C.c_void_p, C.c_size_t
    sig = {
        "zo_f16_to_f32": (C.c_float, [C.c_uint16]),
        "zo_from_slice": (None, [vp, sz, sz, sz, vp, vp]),
        "zo_prepare_transposed": (None, [vp, vp, sz, sz, sz, vp, vp]),
        "zo_quantize_input": (None, [vp, sz, sz, vp, vp]),
        "zo_kv_store_column": (None, [vp, vp, sz, sz, sz, vp]),
        "zo_kv_dequant_column": (None, [vp, vp, sz, sz, sz, vp]),
        "zo_attention_quantized": (C.c_int, [vp, sz, vp, sz, sz, sz, vp, vp, sz, vp, vp, sz, sz, sz, vp, sz, sz,
                                             C.c_float, C.c_int]),
        "zo_gemv_range": (None, [vp, vp, vp, vp, vp, sz, sz, sz, sz]),
        "zo_gemv": (C.c_int, [vp, vp, vp, vp, sz, sz, sz]),
        "zo_gemv_pool": (C.c_int, [vp, vp, vp, vp, sz, sz, sz, sz]),
        "zo_matmul": (None, [vp, vp, sz, vp, vp, sz, sz, sz]),
        "zo_matmul_bias": (None, [vp, vp, sz, vp, vp, vp, sz, sz, sz]),
        "zo_matmul_mt": (None, [vp, vp, sz, vp, vp, sz, sz, sz, sz]),
        "zo_dequantize_to": (None, [vp, vp, sz, sz, vp]),
        "zo_dequant_q4_0": (None, [vp, vp, sz]),
        "zo_dequant_q8_0": (None, [vp, vp, sz]),
        "zo_dequant_f16": (None, [vp, vp, sz]),
        "zo_qweight_from_gguf": (C.c_int, [vp, C.c_uint32, sz, vp, vp]),
        "zo_qmatmul_op": (None, [vp, vp, vp, vp, sz, sz, sz, sz, sz, sz, sz, sz]),
        "zo_run_program": (C.c_int, [C.POINTER(abi.ZgProgram), C.POINTER(abi.ZgIO), sz, C.POINTER(abi.ZgIO), sz]),
        "zo_state_create": (vp, [C.POINTER(abi.ZgProgram)]),
        "zo_state_execute": (None, [vp, C.POINTER(abi.ZgProgram), C.POINTER(abi.ZgOp), sz,
                                    C.POINTER(abi.ZgIO), sz, C.POINTER(abi.ZgIO), sz]),
        "zo_state_destroy": (None, [vp]),
        "zo_set_exec_threads": (None, [sz]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _LIBS[native] = L
    return L


def _p(a):
    return a.ctypes.data


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class QuantizedWeight:
    """quant.zig QuantizedWeight(f32): data i8 [K*N], scales f32, rows=K, cols=N."""

    def __init__(self, data, scales, rows, cols, block_size):
        self.data = np.ascontiguousarray(data, dtype=np.int8).ravel()
        self.scales = np.ascontiguousarray(scales, dtype=np.float32).ravel()
        self.rows, self.cols, self.block_size = rows, cols, block_size
        self.t_data = None
        self.t_scales = None

    @classmethod
    def from_slice(cls, weights, rows, cols, block_size=32):  # src/quant.zig:216-256
        w = f32(weights).ravel()
        assert w.size == rows * cols
        n_blocks = (w.size + block_size - 1) // block_size
        data = np.zeros(w.size, np.int8)
        scales = np.zeros(n_blocks, np.float32)
        lib().zo_from_slice(_p(w), rows, cols, block_size, _p(data), _p(scales))
        return cls(data, scales, rows, cols, block_size)

    @classmethod
    def from_gguf(cls, raw, ggml_type, rows, cols):  # src/models/gguf_loader.zig:99-154
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        n = rows * cols
        data = np.zeros(n, np.int8)
        scales = np.zeros((n + 31) // 32, np.float32)
        rc = lib().zo_qweight_from_gguf(_p(raw), ggml_type, n, _p(data), _p(scales))
        if rc != 0:
            raise ValueError("UnsupportedType")
        return cls(data, scales, rows, cols, 32)

    def prepare_transposed(self):  # src/quant.zig:274-317
        K, N, bs = self.rows, self.cols, self.block_size
        bpr = (K + bs - 1) // bs
        self.t_data = np.zeros(N * K, np.int8)
        self.t_scales = np.zeros(N * bpr, np.float32)
        lib().zo_prepare_transposed(_p(self.data), _p(self.scales), K, N, bs, _p(self.t_data), _p(self.t_scales))

    def gemv(self, x):  # src/quant.zig:443-459
        x = f32(x).ravel()
        out = np.zeros(self.cols, np.float32)
        rc = lib().zo_gemv(_p(self.t_data), _p(self.t_scales), _p(x), _p(out), self.cols, self.rows, self.block_size)
        if rc != 0:
            raise ValueError("K exceeds the reference's stack buffers")
        return out

    def gemv_pool(self, x, n_workers, native=False):  # src/quant.zig:135-196
        x = f32(x).ravel()
        out = np.zeros(self.cols, np.float32)
        lib(native).zo_gemv_pool(_p(self.t_data), _p(self.t_scales), _p(x), _p(out), self.cols, self.rows,
                                 self.block_size, n_workers)
        return out

    def matmul(self, x, M, threads=1, native=False):  # src/quant.zig:475-578
        x = f32(x).ravel()
        out = np.zeros(M * self.cols, np.float32)
        if threads <= 1:
            lib(native).zo_matmul(_p(self.data), _p(self.scales), self.block_size, _p(x), _p(out), M, self.cols, self.rows)
        else:
            lib(native).zo_matmul_mt(_p(self.data), _p(self.scales), self.block_size, _p(x), _p(out), M, self.cols,
                                     self.rows, threads)
        return out.reshape(M, self.cols)

    def matmul_bias(self, x, bias, M):  # src/quant.zig:581-589
        x = f32(x).ravel()
        bias = f32(bias).ravel()
        out = np.zeros(M * self.cols, np.float32)
        lib().zo_matmul_bias(_p(self.data), _p(self.scales), self.block_size, _p(x), _p(bias), _p(out), M, self.cols, self.rows)
        return out.reshape(M, self.cols)

    def dequantize_to(self):  # src/quant.zig:594-618
        out = np.zeros(self.rows * self.cols, np.float32)
        lib().zo_dequantize_to(_p(self.data), _p(self.scales), out.size, self.block_size, _p(out))
        return out.reshape(self.rows, self.cols)

    def qmatmul_op(self, inp, dst, M, input_offset=0, input_row_stride=0, dst_offset=0, dst_row_stride=0):
        """DeviceOp.qmatmul on flat f32 buffers — src/backend/reference.zig:499-566."""
        lib().zo_qmatmul_op(_p(inp), _p(dst), _p(self.data), _p(self.scales), self.block_size, M, self.cols,
                            self.rows, input_offset, input_row_stride, dst_offset, dst_row_stride)


def quantize_input(x, bs):  # src/quant.zig:320-341
    x = f32(x).ravel()
    K = x.size
    q = np.zeros(K, np.int8)
    s = np.zeros((K + bs - 1) // bs, np.float32)
    lib().zo_quantize_input(_p(x), K, bs, _p(q), _p(s))
    return q, s


class QuantizedKVCache:
    """QuantizedKVCache (src/quant.zig:646-761): column-major Q8 cache, column = d_head int8 + d_head/bs scales."""

    def __init__(self, d_head, n_cols, block_size):  # init, src/quant.zig:658-678
        assert d_head % block_size == 0
        self.d_head, self.n_cols, self.block_size = d_head, n_cols, block_size
        self.blocks_per_col = d_head // block_size
        self.q_data = np.zeros(d_head * n_cols, np.int8)
        self.scales = np.zeros(self.blocks_per_col * n_cols, np.float32)

    def store_column(self, col, src):  # src/quant.zig:689-701
        src = f32(src).ravel()
        assert src.size == self.d_head and col < self.n_cols
        lib().zo_kv_store_column(_p(self.q_data), _p(self.scales), self.d_head, self.block_size, col, _p(src))

    def dequant_column(self, col):  # src/quant.zig:704-716
        out = np.zeros(self.d_head, np.float32)
        lib().zo_kv_dequant_column(_p(self.q_data), _p(self.scales), self.d_head, self.block_size, col, _p(out))
        return out


def attention_quantized(q, seq_q, k_cache, k_col_start, v_cache, v_col_start, seq_kv, scale, mask=None, mask_row_stride=0,
                        mask_col_stride=0, use_sdot=True, q_col_stride=None, dst_col_stride=None):
    """attentionQuantized (src/quant.zig:924-1091).  q: [d_head, seq_q] column-major.  use_sdot: the aarch64 branch
    (int8 query); False: the portable f32-query branch.  Returns dst [seq_q, d_head] rows = query columns."""
    d = k_cache.d_head
    q_cs = d if q_col_stride is None else q_col_stride
    d_cs = d if dst_col_stride is None else dst_col_stride
    q = f32(q).ravel()
    dst = np.zeros(max(1, (seq_q - 1) * d_cs + d), np.float32)
    m = None if mask is None else f32(mask).ravel()
    rc = lib().zo_attention_quantized(_p(dst), d_cs, _p(q), q_cs, d, seq_q, _p(k_cache.q_data), _p(k_cache.scales), k_col_start,
                                      _p(v_cache.q_data), _p(v_cache.scales), v_col_start, k_cache.block_size, seq_kv,
                                      None if m is None else _p(m), mask_row_stride, mask_col_stride, scale, 1 if use_sdot else 0)
    if rc != 0:
        raise ValueError("attentionQuantized: d_head > 512 or d_head % block_size != 0")
    return dst


def dequant_q4_0(raw, n):
    raw = np.ascontiguousarray(raw, np.uint8)
    out = np.zeros(n, np.float32)
    lib().zo_dequant_q4_0(_p(out), _p(raw), n)
    return out


def dequant_q8_0(raw, n):
    raw = np.ascontiguousarray(raw, np.uint8)
    out = np.zeros(n, np.float32)
    lib().zo_dequant_q8_0(_p(out), _p(raw), n)
    return out


def dequant_f16(raw, n):
    raw = np.ascontiguousarray(raw, np.uint8)
    out = np.zeros(n, np.float32)
    lib().zo_dequant_f16(_p(out), _p(raw), n)
    return out


def run_program(program, inputs=(), outputs=()):
    """CpuBackend compile+execute+free (src/backend/cpu.zig:55-119) on a zgml_b200.DeviceProgram."""
    from zgml_b200.backend import _io_array
    cprog, keep = program.to_c()
    ia, oa = _io_array(inputs), _io_array(outputs)
    rc = lib().zo_run_program(C.byref(cprog), ia, len(inputs), oa, len(outputs))
    del keep
    if rc != 0:
        raise MemoryError("oracle program allocation failed")


def set_exec_threads(n: int, native: bool = False):
    """Threads the program executor splits a qmatmul's output columns over (bench.py --impl reference: all host cores;
    default 1 like the reference, src/inference_utils.zig:192).  Results are bit-identical for any n."""
    lib(native).zo_set_exec_threads(int(n))


class ProgramState:
    """Persistent reference execution of one DeviceProgram (buffers live across steps)."""

    def __init__(self, program, native=False):
        self._L = lib(native)
        self.program = program
        self.cprog, self._keep = program.to_c()
        self.state = self._L.zo_state_create(C.byref(self.cprog))

    def execute(self, ops_array, n_ops, inputs=(), outputs=()):
        from zgml_b200.backend import _io_array
        ia, oa = _io_array(inputs), _io_array(outputs)
        self._L.zo_state_execute(self.state, C.byref(self.cprog), C.cast(ops_array, C.POINTER(abi.ZgOp)), n_ops,
                                 ia, len(inputs), oa, len(outputs))

    def close(self):
        if self.state:
            self._L.zo_state_destroy(self.state)
            self.state = None
