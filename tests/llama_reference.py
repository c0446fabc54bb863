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


class ShardedOracleBackend(OracleBackend):
    """One rank of a row-sharded run on the CPU: the oracle executor runs the op segments between collectives, the
    collectives themselves (ZG_OP_ALLREDUCE / ZG_OP_ALLGATHER, include/zgml_cuda.h) go through torch.distributed
    (gloo).  Checks the host-side sharding logic without a GPU."""
    name_str = "oracle-cpu-sharded"

    def __init__(self, dist, native=False):
        super().__init__(native)
        self.dist = dist

    def execute_program(self, handle, inputs, outputs):
        import numpy as np
        import torch
        from zgml_b200.backend import ProgramIO
        ops, n = handle.ops_arr, handle.n_ops
        world = self.dist.get_world_size()
        op_sz = C.sizeof(abi.ZgOp)

        def run(lo, hi, ins=(), outs=()):
            ptr = C.cast(C.byref(ops, lo * op_sz), C.POINTER(abi.ZgOp))
            handle.execute(ptr, hi - lo, ins, outs)

        lo, first = 0, True
        for i in range(n):
            tag = ops[i].tag
            if tag not in (abi.OP_ALLREDUCE, abi.OP_ALLGATHER):
                continue
            if tag == abi.OP_ALLREDUCE:
                a = ops[i].u.allreduce
                host = np.zeros(a.n, np.float32)
                run(lo, i, inputs if first else (), [ProgramIO(a.buf, host, offset=a.offset * 4)])
                t = torch.from_numpy(host)
                self.dist.all_reduce(t)
                run(i, i, [ProgramIO(a.buf, host, offset=a.offset * 4)])
            else:
                g = ops[i].u.allgather
                host = np.zeros(g.n, np.float32)
                run(lo, i, inputs if first else (), [ProgramIO(g.src, host, offset=g.src_offset * 4)])
                parts = [torch.zeros(g.n) for _ in range(world)]
                self.dist.all_gather(parts, torch.from_numpy(host))
                full = torch.cat(parts).numpy()
                run(i, i, [ProgramIO(g.dst, full, offset=g.dst_offset * 4)])
            lo, first = i + 1, False
        run(lo, n, inputs if first else (), outputs)
