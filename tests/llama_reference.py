"""Oracle-backed stand-in for a zgml Backend: the same compile/refresh/execute/free surface as
CudaBackend, executed by the CPU reference executor (oracle/zgml_oracle.c).  Test infrastructure."""
import ctypes as C

from oracle import oracle
from zgml_b200 import abi


class OracleBackend:
    name_str = "oracle-cpu"

    def __init__(self, native=False):
        self.native = native

    def compile_program(self, program):
        st = oracle.ProgramState(program, native=self.native)
        st.ops_arr, st.n_ops = program.ops_array(), len(program.ops)
        return st

    def refresh_program(self, handle, ops):
        if hasattr(ops, "arr"):
            handle.ops_arr, handle.n_ops = ops.arr, len(ops)
        else:
            arr = (abi.ZgOp * max(len(ops), 1))()
            for i, o in enumerate(ops):
                arr[i] = o
            handle.ops_arr, handle.n_ops = arr, len(ops)

    def execute_program(self, handle, inputs, outputs):
        handle.execute(handle.ops_arr, handle.n_ops, inputs, outputs)

    def free_program(self, handle):
        handle.close()
