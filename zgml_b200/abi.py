"""ctypes mirror of include/zgml_cuda.h (the C-ABI drop-in boundary).

Field names and order are copied from the header, which copies them from zgml's
`DeviceOp` / `ProgramIO` / `QuantizedWeightUpload` / `DeviceProgram`
(reference src/backend.zig:146-275).
"""
import ctypes as C

# zgml Op ordinals (reference src/op.zig:11-61)
EW_ADD, EW_MUL = 7, 8
EW_NEG, EW_ABS, EW_SGN, EW_STEP, EW_RELU = 9, 10, 11, 12, 13
EW_SQRT, EW_RECIP, EW_EXP, EW_LOG, EW_GELU = 14, 15, 16, 17, 18
EW_SUM, EW_MAX, EW_REPEAT = 19, 20, 21
EW_NAMES = {"add": 7, "mul": 8, "neg": 9, "abs": 10, "sgn": 11, "step": 12, "relu": 13,
            "sqrt": 14, "recip": 15, "exp": 16, "log": 17, "gelu": 18, "sum": 19, "max": 20}

(OP_ELEMENTWISE, OP_MATMUL, OP_QMATMUL, OP_SOFTMAX, OP_LAYERNORM, OP_RMSNORM, OP_REDUCE,
 OP_REPEAT, OP_SLICE_ASSIGN, OP_ROPE, OP_ATTENTION, OP_FUSED_ELEMENTWISE) = range(12)
OP_COUNT = 12
OP_ALLREDUCE, OP_ALLGATHER = 12, 13   # multi-GPU extensions (need comm_init)
QWEIGHT_RESIDENT = (1 << 64) - 1
OP_TAG_NAMES = ["elementwise", "matmul", "qmatmul", "softmax", "layernorm", "rmsnorm", "reduce",
                "repeat", "slice_assign", "rope", "attention", "fused_elementwise"]

QFMT_AUTO, QFMT_I8_F32, QFMT_I8_F16, QFMT_I4_F16, QFMT_GENERIC = 0, 1, 2, 3, 4
QFMT_NAMES = {1: "i8_f32", 2: "i8_f16", 3: "i4_f16", 4: "generic"}
QFMT_BLOCK_BYTES = {1: 36, 2: 34, 3: 18}

u32, f32, sz = C.c_uint32, C.c_float, C.c_size_t


class ZgMatMulGeometry(C.Structure):
    _fields_ = [(n, sz) for n in ("M", "N", "K", "a_row_stride", "a_col_stride", "b_row_stride",
                                  "b_col_stride", "a_offset", "b_offset", "dst_offset", "dst_row_stride")]


class ZgFusedEwStep(C.Structure):
    _fields_ = [("op", u32), ("is_swapped", u32), ("secondary_buf", u32), ("secondary_offset", u32)]


class _Elementwise(C.Structure):
    _fields_ = [(n, u32) for n in ("op", "dst", "src0", "src1", "n", "dst_offset", "src0_offset", "src1_offset")]


class _MatMul(C.Structure):
    _fields_ = [("dst", u32), ("a", u32), ("b", u32), ("_pad", u32), ("geom", ZgMatMulGeometry)]


class _QMatMul(C.Structure):
    _fields_ = [(n, u32) for n in ("dst", "input", "weight_idx", "M", "N", "K", "input_offset",
                                   "input_row_stride", "dst_offset", "dst_row_stride")]


class _Softmax(C.Structure):
    _fields_ = [(n, u32) for n in ("dst", "src", "rows", "cols", "src_offset", "dst_offset")]


class _Norm(C.Structure):
    _fields_ = [("dst", u32), ("src", u32), ("rows", u32), ("cols", u32), ("eps", f32),
                ("src_offset", u32), ("dst_offset", u32)]


class _Reduce(C.Structure):
    _fields_ = [(n, u32) for n in ("op", "dst", "src", "n_out", "reduce_size", "src_offset", "dst_offset")]


class _Repeat(C.Structure):
    _fields_ = [("dst", u32), ("src", u32), ("n", u32), ("src_ne", u32 * 4), ("dst_ne", u32 * 4),
                ("src_strides", u32 * 4), ("dst_strides", u32 * 4), ("src_offset", u32), ("dst_offset", u32)]


class _SliceAssign(C.Structure):
    _fields_ = [(n, u32) for n in ("dst", "src", "rows", "cols", "dst_base_offset", "dst_offset",
                                   "dst_row_stride", "dst_col_stride", "src_offset", "src_row_stride",
                                   "src_col_stride", "patch_stride")]


class _Rope(C.Structure):
    _fields_ = [(n, u32) for n in ("dst", "src", "cos_sin", "half_d", "seq_len", "src_off", "cs_off",
                                   "dst_off", "src_rs", "src_cs", "cs_cs")]


class _Attention(C.Structure):
    _fields_ = ([(n, u32) for n in ("dst", "q", "k", "v", "mask", "has_mask", "d_head", "seq_q", "seq_kv")]
                + [("scale", f32)]
                + [(n, u32) for n in ("q_off", "k_off", "v_off", "mask_off", "dst_off", "q_rs", "q_cs",
                                      "k_rs", "k_cs", "v_rs", "v_cs", "mask_rs", "mask_cs", "dst_rs", "dst_cs")])


class _FusedElementwise(C.Structure):
    _fields_ = [("steps", C.POINTER(ZgFusedEwStep)), ("n_steps", sz), ("n", u32), ("dst", u32),
                ("src", u32), ("dst_offset", u32), ("src_offset", u32)]


class _AllReduce(C.Structure):
    _fields_ = [(n, u32) for n in ("buf", "offset", "n")]


class _AllGather(C.Structure):
    _fields_ = [(n, u32) for n in ("dst", "src", "n", "dst_offset", "src_offset")]


class _OpUnion(C.Union):
    _fields_ = [("elementwise", _Elementwise), ("matmul", _MatMul), ("qmatmul", _QMatMul),
                ("softmax", _Softmax), ("layernorm", _Norm), ("rmsnorm", _Norm), ("reduce", _Reduce),
                ("repeat", _Repeat), ("slice_assign", _SliceAssign), ("rope", _Rope),
                ("attention", _Attention), ("fused_elementwise", _FusedElementwise),
                ("allreduce", _AllReduce), ("allgather", _AllGather)]


class ZgOp(C.Structure):
    _fields_ = [("tag", u32), ("_pad", u32), ("u", _OpUnion)]


class ZgIO(C.Structure):
    _fields_ = [("buf_idx", u32), ("offset", u32), ("host_ptr", C.c_void_p), ("size", u32), ("_pad", u32)]


class ZgQWeight(C.Structure):
    _fields_ = [("data", C.c_void_p), ("n_data", sz), ("scales", C.c_void_p), ("n_scales", sz),
                ("rows", sz), ("cols", sz), ("block_size", sz)]


class ZgProgram(C.Structure):
    _fields_ = [("ops", C.POINTER(ZgOp)), ("n_ops", sz), ("n_buffers", sz), ("buffer_sizes", C.POINTER(sz)),
                ("initial_uploads", C.POINTER(ZgIO)), ("n_uploads", sz), ("qweights", C.POINTER(ZgQWeight)),
                ("n_qweights", sz)]


class ZgProfile(C.Structure):
    _fields_ = [("time_ns", C.c_uint64 * OP_COUNT), ("backend_op_count", C.c_uint64),
                ("fallback_op_count", C.c_uint64), ("backend_dispatch_count", C.c_uint64),
                ("sync_time_ns", C.c_uint64), ("sync_count", C.c_uint64), ("call_count", u32), ("_pad", u32)]


class ZgCapabilities(C.Structure):
    _fields_ = [(n, u32) for n in ("compiled_programs", "host_visible_program_memory", "dense_matmul_f32",
                                   "dense_matmul_f16", "qmatmul", "fused_elementwise",
                                   "max_fused_elementwise_steps", "dynamic_program_refresh",
                                   "prefill_attention", "decode_attention", "quantized_kv",
                                   "attention_supported", "attention_max_seq_kv", "attention_max_d_head")]


# Every symbol include/zgml_cuda.h declares: (name, restype, argtypes)
vp = C.c_void_p
SYMBOLS = [
    ("zg_cuda_create", vp, [C.c_int]),
    ("zg_cuda_destroy", None, [vp]),
    ("zg_cuda_capabilities", None, [C.POINTER(ZgCapabilities)]),
    ("zg_cuda_dense_matmul_f32", C.c_int, [vp, vp, vp, vp, C.POINTER(ZgMatMulGeometry)]),
    ("zg_cuda_compile", vp, [vp, C.POINTER(ZgProgram)]),
    ("zg_cuda_refresh", None, [vp, vp, C.POINTER(ZgOp), sz]),
    ("zg_cuda_execute", None, [vp, vp, C.POINTER(ZgIO), sz, C.POINTER(ZgIO), sz]),
    ("zg_cuda_free", None, [vp, vp]),
    ("zg_cuda_profile", C.POINTER(ZgProfile), [vp, vp]),
    ("zg_cuda_set_profiling", None, [vp, C.c_int]),
    ("zg_cuda_last_error", C.c_char_p, []),
    ("zg_cuda_set_stream", None, [vp, vp]),
    ("zg_cuda_sync", None, [vp]),
    ("zg_cuda_launch_count", C.c_uint64, []),
    ("zg_cuda_set_graph_mode", None, [vp, C.c_int]),
    ("zg_cuda_program_buffer", vp, [vp, u32]),
    ("zg_cuda_program_stats", C.c_uint64, [vp, C.c_int]),
    ("zg_cuda_execute_device", None, [vp, vp]),
    ("zg_cuda_qweight_upload", vp, [vp, C.POINTER(ZgQWeight), C.c_int]),
    ("zg_cuda_qweight_upload_gguf", vp, [vp, vp, sz, u32, sz, sz]),
    ("zg_cuda_qweight_synth_gguf", vp, [vp, C.c_uint64, C.c_uint64, u32, sz, sz, sz, sz, sz, sz]),
    ("zg_cuda_qweight_free", None, [vp, vp]),
    ("zg_cuda_qweight_format", C.c_int, [vp]),
    ("zg_cuda_qweight_device_bytes", sz, [vp]),
    ("zg_cuda_qweight_dequantize", C.c_int, [vp, vp, vp]),
    ("zg_cuda_qmatmul_device", C.c_int, [vp, vp, vp, vp, u32, u32, u32]),
    ("zg_cuda_qmatmul_host", C.c_int, [vp, vp, vp, vp, u32]),
    ("zg_cuda_qweight_from_f32", vp, [vp, vp, sz, sz, sz, vp, vp]),
    ("zg_cuda_qmatmul_bias_host", C.c_int, [vp, vp, vp, vp, vp, u32]),
    ("zg_cuda_qweight_prepare_transposed", C.c_int, [vp, vp, vp, vp]),
    ("zg_cuda_quantize_input_host", C.c_int, [vp, vp, sz, sz, vp, vp]),
    ("zg_cuda_gemv_w8a8_device", C.c_int, [vp, vp, vp, vp]),
    ("zg_cuda_gemv_w8a8_host", C.c_int, [vp, vp, vp, vp]),
    ("zg_cuda_kvcache_create", vp, [vp, sz, sz, sz]),
    ("zg_cuda_kvcache_free", None, [vp, vp]),
    ("zg_cuda_kvcache_clear", C.c_int, [vp, vp]),
    ("zg_cuda_kvcache_store_device", C.c_int, [vp, vp, sz, sz, vp]),
    ("zg_cuda_kvcache_store_host", C.c_int, [vp, vp, sz, sz, vp]),
    ("zg_cuda_kvcache_download", C.c_int, [vp, vp, vp, vp]),
    ("zg_cuda_attention_quantized_device", C.c_int, [vp, vp, sz, vp, sz, sz, sz, vp, sz, vp, sz, sz, vp, sz, sz, C.c_float, C.c_int]),
    ("zg_cuda_attention_quantized_host", C.c_int, [vp, vp, sz, vp, sz, sz, sz, vp, sz, vp, sz, sz, vp, sz, sz, C.c_float, C.c_int]),
    ("zg_cuda_program_quantize_kv", C.c_int, [vp, vp, sz, C.c_int]),
    ("zg_cuda_program_promote_dense", C.c_int, [vp, vp, C.c_int]),
    ("zg_cuda_comm_unique_id", C.c_int, [vp]),
    ("zg_cuda_comm_init", C.c_int, [vp, vp, C.c_int, C.c_int]),
    ("zg_cuda_comm_destroy", None, [vp]),
    ("zg_cuda_comm_mode", C.c_int, [vp]),
    ("zg_cuda_trace", C.c_int, [vp, C.c_int]),
    ("zg_cuda_trace_read", sz, [vp, vp, sz]),
    ("zg_cuda_malloc", vp, [vp, sz]),
    ("zg_cuda_free_device", None, [vp, vp]),
    ("zg_cuda_memcpy_h2d", C.c_int, [vp, vp, vp, sz]),
    ("zg_cuda_memcpy_d2h", C.c_int, [vp, vp, vp, sz]),
    ("zg_cuda_memset", C.c_int, [vp, vp, C.c_int, sz]),
]
