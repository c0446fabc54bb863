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


class QuantKVOracleBackend(OracleBackend):
    """The reference's quantized-KV session mode on the CPU (LlamaInferenceSession.quantizeKV + executeOneQuantized,
    src/llama_inference.zig:277-377,648-679): the oracle executor runs the op segments, every patched slice_assign into a
    buffer that attention reads as K / V goes to QuantizedKVCache.storeColumn and those attention ops to
    attentionQuantized (oracle/zgml_oracle.c restatements of src/quant.zig:633-1091)."""
    name_str = "oracle-cpu-quantized-kv"

    def __init__(self, block_size=32, int8_query=False, native=False):
        super().__init__(native)
        self.bs, self.int8_query = block_size, int8_query

    def compile_program(self, program):
        st = super().compile_program(program)
        ops = program.ops
        st.kv = {}
        for o in ops:
            if o.tag == abi.OP_ATTENTION:
                a = o.u.attention
                for b in (a.k, a.v):
                    if b not in st.kv:
                        st.kv[b] = oracle.QuantizedKVCache(a.d_head, program.buffer_sizes[b] // a.d_head, self.bs)
        return st

    def execute_program(self, handle, inputs, outputs):
        import numpy as np
        from zgml_b200.backend import ProgramIO
        ops, n = handle.ops_arr, handle.n_ops
        op_sz = C.sizeof(abi.ZgOp)
        sizes = handle.program.buffer_sizes

        def run(lo, hi, ins=(), outs=()):
            ptr = C.cast(C.byref(ops, lo * op_sz), C.POINTER(abi.ZgOp))
            handle.execute(ptr, hi - lo, ins, outs)

        def fetch(buf):
            host = np.zeros(sizes[buf], np.float32)
            run(0, 0, (), [ProgramIO(buf, host)])
            return host

        lo, first = 0, True
        for i in range(n):
            tag = ops[i].tag
            if tag == abi.OP_SLICE_ASSIGN and ops[i].u.slice_assign.dst in handle.kv:
                sa = ops[i].u.slice_assign
                run(lo, i, inputs if first else ())
                src = fetch(sa.src)
                cache, d = handle.kv[sa.dst], sa.rows
                for c in range(sa.cols):                     # storeColumn per written column (llama_inference.zig:336-348)
                    cache.store_column(sa.dst_offset // d + c, src[sa.src_offset + c * sa.src_col_stride:][:d])
            elif tag == abi.OP_ATTENTION and ops[i].u.attention.k in handle.kv:
                a = ops[i].u.attention
                run(lo, i, inputs if first else ())
                q, dst = fetch(a.q), fetch(a.dst)
                mask = fetch(a.mask)[a.mask_off:] if a.has_mask else None
                out = oracle.attention_quantized(q[a.q_off:], a.seq_q, handle.kv[a.k], a.k_off // a.d_head, handle.kv[a.v], a.v_off // a.d_head,
                                                 a.seq_kv, a.scale, mask, a.mask_rs, a.mask_cs if a.seq_q > 1 else 0, use_sdot=self.int8_query,
                                                 q_col_stride=a.q_cs, dst_col_stride=a.dst_cs)
                for c in range(a.seq_q):
                    dst[a.dst_off + c * a.dst_cs:][:a.d_head] = out[c * a.dst_cs:][:a.d_head]
                run(i, i, [ProgramIO(a.dst, dst)])
            else:
                continue
            lo, first = i + 1, False
        run(lo, n, inputs if first else (), outputs)
