"""zgml_b200 — B200 (sm_100a) CUDA backend for zgml's quantized-weight matmul path.

Only what the path needs: csrc/ (CUDA kernels + the C-ABI of include/zgml_cuda.h)
and the host-side mirror of zgml's backend interface (backend.py, host/llama.py,
host/gguf.py).  No CPU fallback: the CUDA library must be built and a B200 present.
"""
from . import abi  # noqa: F401
from .backend import (BackendError, CompiledHandle, CudaBackend, DeviceOp, DeviceProgram,  # noqa: F401
                      ProgramIO, QuantizedWeight, QuantizedWeightUpload, library_path, load_library)
