"""CPU-side check that the C-ABI library builds, loads and exports every symbol that
include/zgml_cuda.h declares (no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re

from zgml_b200 import abi
from zgml_b200.build import build_cuda

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "zgml_cuda.h")).read()
    return sorted(set(re.findall(r"\b(zg_cuda_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    path = build_cuda()
    lib = C.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in zgml_cuda.h but not exported"
    bound = {n for n, _, _ in abi.SYMBOLS}
    assert set(declared) == bound, (set(declared) ^ bound)


def test_struct_layouts_match_header_sizes():
    # sizes the C compiler gives the PODs (LP64): guards the ctypes mirror against drift
    import subprocess, tempfile
    src = r'''
    #include <stdio.h>
    #include "zgml_cuda.h"
    int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(ZgOp), sizeof(ZgIO), sizeof(ZgQWeight),
        sizeof(ZgProgram), sizeof(ZgProfile), sizeof(ZgMatMulGeometry), sizeof(ZgCapabilities)); return 0; }
    '''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", exe, c], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(abi.ZgOp), C.sizeof(abi.ZgIO), C.sizeof(abi.ZgQWeight), C.sizeof(abi.ZgProgram),
            C.sizeof(abi.ZgProfile), C.sizeof(abi.ZgMatMulGeometry), C.sizeof(abi.ZgCapabilities)]
    assert sizes == want


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    from zgml_b200 import BackendError, CudaBackend
    try:
        CudaBackend(0)
    except BackendError as e:
        assert "CUDA" in str(e) or "device" in str(e)
    else:
        raise AssertionError("CudaBackend() must raise without a GPU: there is no CPU fallback")
