"""The 11 single-op DevicePrograms of the reference's backend conformance harness
(reference src/backend/conformance.zig:62-346), restated with this repo's DeviceOp
constructors.  Used twice: tests/test_oracle_golden.py pins the oracle executor on them
(CPU), tests/test_gpu_conformance.py runs them on the CUDA backend against the oracle
exactly like `assertBackendMatchesReference` (conformance.zig:52-60) does.

Each entry: (name, DeviceProgram, out_buf_idx, out_len, numpy_expected or None).
"""
import numpy as np

from zgml_b200 import DeviceOp, DeviceProgram, ProgramIO, QuantizedWeightUpload


def f32(*v):
    return np.array(v, np.float32)


def core_cases():
    cases = []

    # 1. matmul — conformance.zig:63-79
    a, b = f32(1, 2, 3, 4, 5, 6), f32(7, 8, 9, 10, 11, 12)
    cases.append(("matmul", DeviceProgram([DeviceOp.matmul(2, 0, 1, 2, 2, 3, 3, 1, 2, 1, dst_row_stride=2)], [6, 6, 4],
                                          [ProgramIO(0, a), ProgramIO(1, b)]), 2, 4, f32(58, 64, 139, 154)))

    # 2. qmatmul with offsets/strides — conformance.zig:81-112
    inp, dst = f32(99, 1, 2, 3, 99, -1, 0.5, 4, 99), np.full(9, -7, np.float32)
    qw = QuantizedWeightUpload(np.array([2, -1, 3, 4, -2, 1, -3, 5, 2], np.int8), f32(0.5, 0.25, 1.0), 3, 3, 4)
    cases.append(("qmatmul", DeviceProgram([DeviceOp.qmatmul(1, 0, 0, 2, 3, 3, 1, 4, 1, 4)], [9, 9],
                                           [ProgramIO(0, inp), ProgramIO(1, dst)], [qw]), 1, 9,
                  f32(-7, 2.75, 2.25, 8.0, -7, -3.0, 5.25, 6.625, -7)))

    # 3. elementwise add — conformance.zig:114-132
    a, b = f32(1, 2, 3, 4), f32(10, 20, 30, 40)
    cases.append(("add", DeviceProgram([DeviceOp.elementwise("add", 2, 0, 1, 4)], [4, 4, 4],
                                       [ProgramIO(0, a), ProgramIO(1, b)]), 2, 4, f32(11, 22, 33, 44)))

    # 4. reduce sum + max — conformance.zig:134-158
    src = f32(1, -2, 3, 4, 5, -6)
    cases.append(("reduce", DeviceProgram([DeviceOp.reduce("sum", 1, 0, 2, 3), DeviceOp.reduce("max", 1, 0, 2, 3, dst_offset=2)],
                                          [6, 4], [ProgramIO(0, src)]), 1, 4, f32(2, 3, 3, 5)))

    # 5. repeat — conformance.zig:160-175
    src = f32(7, 8)
    cases.append(("repeat", DeviceProgram([DeviceOp.repeat(1, 0, 6, (2, 1, 1, 1), (2, 3, 1, 1), (1, 2, 2, 2), (1, 2, 6, 6))],
                                          [2, 6], [ProgramIO(0, src)]), 1, 6, f32(7, 8, 7, 8, 7, 8)))

    # 6. slice_assign — conformance.zig:177-201
    src, dst = f32(99, 2, 3, 5, 6, 77), f32(10, 11, 12, 13, 14, 15, 16, 17)
    cases.append(("slice_assign", DeviceProgram([DeviceOp.slice_assign(1, 0, 2, 2, 0, 2, 1, 2, 1, 1, 2, 2)], [6, 8],
                                                [ProgramIO(0, src), ProgramIO(1, dst)]), 1, 8,
                  f32(10, 11, 2, 3, 5, 6, 16, 17)))

    # 7. softmax — conformance.zig:203-216
    src = f32(1, 2, 3, -1, 0, 1)
    e = np.exp(src.reshape(2, 3).astype(np.float64) - src.reshape(2, 3).max(1, keepdims=True))
    cases.append(("softmax", DeviceProgram([DeviceOp.softmax(1, 0, 2, 3)], [6, 6], [ProgramIO(0, src)]), 1, 6,
                  (e / e.sum(1, keepdims=True)).astype(np.float32).ravel()))

    # 8. layernorm + rmsnorm — conformance.zig:218-242
    src = f32(1, 2, 3, 4, -1, 0, 1, 2)
    x = src.reshape(2, 4).astype(np.float64)
    ln = (x - x.mean(1, keepdims=True)) / np.sqrt(x.var(1, keepdims=True) + 1e-5)
    rn = x / np.sqrt((x * x).mean(1, keepdims=True) + 1e-5)
    cases.append(("norms", DeviceProgram([DeviceOp.layernorm(1, 0, 2, 4, 1e-5), DeviceOp.rmsnorm(1, 0, 2, 4, 1e-5, dst_offset=8)],
                                         [8, 16], [ProgramIO(0, src)]), 1, 16,
                  np.concatenate([ln.ravel(), rn.ravel()]).astype(np.float32)))

    # 9. rope — conformance.zig:244-268 (col 0: cos=1,sin=0 -> identity; col 1: cos=0,sin=1 -> (-hi, lo))
    src, cs = f32(1, 2, 3, 4, 5, 6, 7, 8), f32(1, 1, 0, 0, 0, 0, 1, 1)
    cases.append(("rope", DeviceProgram([DeviceOp.rope(2, 0, 1, 2, 2, 0, 0, 0, 1, 4, 4)], [8, 8, 8],
                                        [ProgramIO(0, src), ProgramIO(1, cs)]), 2, 8, f32(1, 2, 3, 4, -7, -8, 5, 6)))

    # 10. masked attention — conformance.zig:270-321
    q = f32(0.2, 0.1, -0.3, 0.4, -0.1, 0.5, 0.2, -0.4)
    k = f32(0.1, 0.2, 0.3, 0.4, -0.2, 0.3, 0.1, -0.1, 0.5, -0.4, 0.2, 0.1)
    v = f32(1, 2, 3, 4, -1, 0.5, 2, -0.5, 0.25, -0.75, 1.5, 2.5)
    mask = f32(0, 0, -np.inf, 0, -0.25, 0)
    s = (q.reshape(2, 4).astype(np.float64) @ k.reshape(3, 4).astype(np.float64).T) * 0.5 + mask.reshape(2, 3)
    p = np.exp(s - s.max(1, keepdims=True))
    p /= p.sum(1, keepdims=True)
    want = (p @ v.reshape(3, 4).astype(np.float64)).astype(np.float32).ravel()
    cases.append(("attention", DeviceProgram(
        [DeviceOp.attention(4, 0, 1, 2, 3, True, 4, 2, 3, 0.5, 0, 0, 0, 0, 0, 1, 4, 1, 4, 1, 4, 1, 3, 1, 4)],
        [8, 12, 12, 6, 8], [ProgramIO(0, q), ProgramIO(1, k), ProgramIO(2, v), ProgramIO(3, mask)]), 4, 8, want))

    # 11. fused chain relu -> sqrt -> add — conformance.zig:323-345
    src, addend = f32(1, -2, 4, 9), f32(10, 20, 30, 40)
    cases.append(("fused", DeviceProgram(
        [DeviceOp.fused_elementwise([("relu", False, 0, 0), ("sqrt", False, 0, 0), ("add", False, 1, 0)], 4, 2, 0)],
        [4, 4, 4], [ProgramIO(0, src), ProgramIO(1, addend)]), 2, 4, f32(11, 20, 32, 43)))
    return cases
