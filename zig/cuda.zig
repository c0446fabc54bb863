//! src/backend/cuda.zig — zgml `Backend` over libzgml_cuda.so (B200 / sm_100a).
//!
//! This file is the ONLY reference-side change besides one `pub const cuda = @import(...)`
//! line: it flattens zgml's slices / tagged unions (src/backend.zig:179-275) into the PODs of
//! include/zgml_cuda.h and forwards the six vtable slots (src/backend.zig:339-352) to the
//! `zg_cuda_*` C entry points, the same way src/backend/metal.zig wraps metal_shim.h.
//! It contains no arithmetic and no fallbacks.
//!
//! NOT compiled in this repository's CI: the build image has no zig toolchain (DESIGN.md §1).
//! The C-ABI it binds is exercised through the identical ctypes binding in zgml_b200/abi.py.
//!
//! build.zig additions:   exe.linkSystemLibrary("zgml_cuda"); exe.addLibraryPath(.{ .cwd_relative = "<repo>/zgml_b200/lib" });

const std = @import("std");
const backend_mod = @import("../backend.zig");
const profile_mod = @import("../profile.zig");

// ── C-ABI (include/zgml_cuda.h) ─────────────────────────────────────────────
const ZgMatMulGeometry = extern struct {
    M: usize, N: usize, K: usize,
    a_row_stride: usize, a_col_stride: usize, b_row_stride: usize, b_col_stride: usize,
    a_offset: usize, b_offset: usize, dst_offset: usize, dst_row_stride: usize,
};
const ZgFusedEwStep = extern struct { op: u32, is_swapped: u32, secondary_buf: u32, secondary_offset: u32 };
const ZgOp = extern struct {
    tag: u32,
    _pad: u32 = 0,
    u: extern union {
        elementwise: extern struct { op: u32, dst: u32, src0: u32, src1: u32, n: u32, dst_offset: u32, src0_offset: u32, src1_offset: u32 },
        matmul: extern struct { dst: u32, a: u32, b: u32, _pad: u32 = 0, geom: ZgMatMulGeometry },
        qmatmul: extern struct { dst: u32, input: u32, weight_idx: u32, M: u32, N: u32, K: u32, input_offset: u32, input_row_stride: u32, dst_offset: u32, dst_row_stride: u32 },
        softmax: extern struct { dst: u32, src: u32, rows: u32, cols: u32, src_offset: u32, dst_offset: u32 },
        norm: extern struct { dst: u32, src: u32, rows: u32, cols: u32, eps: f32, src_offset: u32, dst_offset: u32 }, // layernorm + rmsnorm
        reduce: extern struct { op: u32, dst: u32, src: u32, n_out: u32, reduce_size: u32, src_offset: u32, dst_offset: u32 },
        repeat: extern struct { dst: u32, src: u32, n: u32, src_ne: [4]u32, dst_ne: [4]u32, src_strides: [4]u32, dst_strides: [4]u32, src_offset: u32, dst_offset: u32 },
        slice_assign: extern struct { dst: u32, src: u32, rows: u32, cols: u32, dst_base_offset: u32, dst_offset: u32, dst_row_stride: u32, dst_col_stride: u32, src_offset: u32, src_row_stride: u32, src_col_stride: u32, patch_stride: u32 },
        rope: extern struct { dst: u32, src: u32, cos_sin: u32, half_d: u32, seq_len: u32, src_off: u32, cs_off: u32, dst_off: u32, src_rs: u32, src_cs: u32, cs_cs: u32 },
        attention: extern struct {
            dst: u32, q: u32, k: u32, v: u32, mask: u32, has_mask: u32, d_head: u32, seq_q: u32, seq_kv: u32, scale: f32,
            q_off: u32, k_off: u32, v_off: u32, mask_off: u32, dst_off: u32,
            q_rs: u32, q_cs: u32, k_rs: u32, k_cs: u32, v_rs: u32, v_cs: u32, mask_rs: u32, mask_cs: u32, dst_rs: u32, dst_cs: u32,
        },
        fused_elementwise: extern struct { steps: ?[*]const ZgFusedEwStep, n_steps: usize, n: u32, dst: u32, src: u32, dst_offset: u32, src_offset: u32 },
        // extensions (tags 12 / 13): collectives of a row-sharded program; zgml's DeviceOp has no such variants yet
        allreduce: extern struct { buf: u32, offset: u32, n: u32 },
        allgather: extern struct { dst: u32, src: u32, n: u32, dst_offset: u32, src_offset: u32 },
    },
};
const ZgIO = extern struct { buf_idx: u32, offset: u32, host_ptr: ?*anyopaque, size: u32, _pad: u32 = 0 };
const ZgQWeight = extern struct { data: [*]const i8, n_data: usize, scales: [*]const f32, n_scales: usize, rows: usize, cols: usize, block_size: usize };
const ZgProgram = extern struct {
    ops: [*]const ZgOp, n_ops: usize, n_buffers: usize, buffer_sizes: [*]const usize,
    initial_uploads: [*]const ZgIO, n_uploads: usize, qweights: [*]const ZgQWeight, n_qweights: usize,
};
const ZgProfile = extern struct { time_ns: [12]u64, backend_op_count: u64, fallback_op_count: u64, backend_dispatch_count: u64, sync_time_ns: u64, sync_count: u64, call_count: u32, _pad: u32 };

extern fn zg_cuda_create(device_ordinal: c_int) ?*anyopaque;
extern fn zg_cuda_destroy(ctx: *anyopaque) void;
extern fn zg_cuda_compile(ctx: *anyopaque, program: *const ZgProgram) ?*anyopaque;
extern fn zg_cuda_refresh(ctx: *anyopaque, prog: *anyopaque, ops: [*]const ZgOp, n_ops: usize) void;
extern fn zg_cuda_execute(ctx: *anyopaque, prog: *anyopaque, inputs: [*]const ZgIO, n_in: usize, outputs: [*]const ZgIO, n_out: usize) void;
extern fn zg_cuda_free(ctx: *anyopaque, prog: *anyopaque) void;
extern fn zg_cuda_profile(ctx: *anyopaque, prog: *anyopaque) ?*const ZgProfile;
extern fn zg_cuda_last_error() [*:0]const u8;
// extensions: multi-GPU communicator (one process per GPU) and weights packed in HBM ahead of compile
extern fn zg_cuda_comm_unique_id(id128: *[128]u8) c_int;
extern fn zg_cuda_comm_init(ctx: *anyopaque, id128: *const [128]u8, rank: c_int, world: c_int) c_int;
extern fn zg_cuda_comm_destroy(ctx: *anyopaque) void;
extern fn zg_cuda_qweight_upload_gguf(ctx: *anyopaque, raw: [*]const u8, raw_bytes: usize, ggml_type: u32, rows: usize, cols: usize) ?*anyopaque;
extern fn zg_cuda_qweight_free(ctx: *anyopaque, w: *anyopaque) void;
// W8A8 decode path (src/quant.zig:274-459): transposed re-quantization at load, then quantizeInput + gemvRange per call
extern fn zg_cuda_qweight_prepare_transposed(ctx: *anyopaque, w: *anyopaque, h_t_data: ?[*]i8, h_t_scales: ?[*]f32) c_int;
extern fn zg_cuda_quantize_input_host(ctx: *anyopaque, h_input: [*]const f32, K: usize, block_size: usize, h_q: [*]i8, h_scales: [*]f32) c_int;
extern fn zg_cuda_gemv_w8a8_device(ctx: *anyopaque, w: *const anyopaque, d_input: *const anyopaque, d_dst: *anyopaque) c_int;
// quantized KV cache (src/quant.zig:633-1091)
extern fn zg_cuda_kvcache_create(ctx: *anyopaque, d_head: usize, n_cols: usize, block_size: usize) ?*anyopaque;
extern fn zg_cuda_kvcache_free(ctx: *anyopaque, cache: *anyopaque) void;
extern fn zg_cuda_kvcache_clear(ctx: *anyopaque, cache: *anyopaque) c_int;
extern fn zg_cuda_kvcache_store_device(ctx: *anyopaque, cache: *anyopaque, col_start: usize, n_write: usize, d_src: *const anyopaque) c_int;
extern fn zg_cuda_attention_quantized_device(ctx: *anyopaque, d_dst: *anyopaque, dst_col_stride: usize, d_q: *const anyopaque, q_col_stride: usize, d_head: usize, seq_q: usize, k_cache: *const anyopaque, k_col_start: usize, v_cache: *const anyopaque, v_col_start: usize, seq_kv: usize, d_mask: ?*const anyopaque, mask_row_stride: usize, mask_col_stride: usize, scale: f32, int8_query: c_int) c_int;
extern fn zg_cuda_program_quantize_kv(ctx: *anyopaque, prog: *anyopaque, block_size: usize, int8_query: c_int) c_int;
extern fn zg_cuda_program_promote_dense(ctx: *anyopaque, prog: *anyopaque, format: c_int) c_int;
extern fn zg_cuda_gemv_w8a8_host(ctx: *anyopaque, w: *const anyopaque, h_input: [*]const f32, h_dst: [*]f32) c_int;
pub const ZG_QWEIGHT_RESIDENT: usize = std.math.maxInt(usize); // ZgQWeight.block_size marker: `data` is a handle from zg_cuda_qweight_upload*

// ── flattening ──────────────────────────────────────────────────────────────
const Flat = struct {
    ops: []ZgOp,
    steps: []ZgFusedEwStep,

    fn deinit(self: Flat, alloc: std.mem.Allocator) void {
        alloc.free(self.ops);
        alloc.free(self.steps);
    }
};

/// DeviceOp tags are the union's declaration order (src/backend.zig:179-249) == ZgOpTag.
fn flatten(alloc: std.mem.Allocator, ops: []const backend_mod.DeviceOp) !Flat {
    var n_steps: usize = 0;
    for (ops) |op| switch (op) {
        .fused_elementwise => |f| n_steps += f.steps.len,
        else => {},
    };
    const out = try alloc.alloc(ZgOp, ops.len);
    errdefer alloc.free(out);
    const steps = try alloc.alloc(ZgFusedEwStep, n_steps);
    var si: usize = 0;
    for (ops, out) |op, *o| {
        o.* = switch (op) {
            .elementwise => |e| .{ .tag = 0, .u = .{ .elementwise = .{ .op = @intFromEnum(e.op), .dst = e.dst, .src0 = e.src0, .src1 = e.src1, .n = e.n, .dst_offset = e.dst_offset, .src0_offset = e.src0_offset, .src1_offset = e.src1_offset } } },
            // MatMulGeometry is an auto-layout Zig struct (src/backend.zig:146-158): copied field by field, never bit-cast
            .matmul => |m| .{ .tag = 1, .u = .{ .matmul = .{ .dst = m.dst, .a = m.a, .b = m.b, .geom = .{
                .M = m.geom.M, .N = m.geom.N, .K = m.geom.K,
                .a_row_stride = m.geom.a_row_stride, .a_col_stride = m.geom.a_col_stride,
                .b_row_stride = m.geom.b_row_stride, .b_col_stride = m.geom.b_col_stride,
                .a_offset = m.geom.a_offset, .b_offset = m.geom.b_offset,
                .dst_offset = m.geom.dst_offset, .dst_row_stride = m.geom.dst_row_stride,
            } } } },
            .qmatmul => |q| .{ .tag = 2, .u = .{ .qmatmul = .{ .dst = q.dst, .input = q.input, .weight_idx = q.weight_idx, .M = q.M, .N = q.N, .K = q.K, .input_offset = q.input_offset, .input_row_stride = q.input_row_stride, .dst_offset = q.dst_offset, .dst_row_stride = q.dst_row_stride } } },
            .softmax => |s| .{ .tag = 3, .u = .{ .softmax = .{ .dst = s.dst, .src = s.src, .rows = s.rows, .cols = s.cols, .src_offset = s.src_offset, .dst_offset = s.dst_offset } } },
            .layernorm => |l| .{ .tag = 4, .u = .{ .norm = .{ .dst = l.dst, .src = l.src, .rows = l.rows, .cols = l.cols, .eps = l.eps, .src_offset = l.src_offset, .dst_offset = l.dst_offset } } },
            .rmsnorm => |l| .{ .tag = 5, .u = .{ .norm = .{ .dst = l.dst, .src = l.src, .rows = l.rows, .cols = l.cols, .eps = l.eps, .src_offset = l.src_offset, .dst_offset = l.dst_offset } } },
            .reduce => |r| .{ .tag = 6, .u = .{ .reduce = .{ .op = @intFromEnum(r.op), .dst = r.dst, .src = r.src, .n_out = r.n_out, .reduce_size = r.reduce_size, .src_offset = r.src_offset, .dst_offset = r.dst_offset } } },
            .repeat => |r| .{ .tag = 7, .u = .{ .repeat = .{ .dst = r.dst, .src = r.src, .n = r.n, .src_ne = r.src_ne, .dst_ne = r.dst_ne, .src_strides = r.src_strides, .dst_strides = r.dst_strides, .src_offset = r.src_offset, .dst_offset = r.dst_offset } } },
            .slice_assign => |s| .{ .tag = 8, .u = .{ .slice_assign = .{ .dst = s.dst, .src = s.src, .rows = s.rows, .cols = s.cols, .dst_base_offset = s.dst_base_offset, .dst_offset = s.dst_offset, .dst_row_stride = s.dst_row_stride, .dst_col_stride = s.dst_col_stride, .src_offset = s.src_offset, .src_row_stride = s.src_row_stride, .src_col_stride = s.src_col_stride, .patch_stride = s.patch_stride } } },
            .rope => |r| .{ .tag = 9, .u = .{ .rope = .{ .dst = r.dst, .src = r.src, .cos_sin = r.cos_sin, .half_d = r.half_d, .seq_len = r.seq_len, .src_off = r.src_off, .cs_off = r.cs_off, .dst_off = r.dst_off, .src_rs = r.src_rs, .src_cs = r.src_cs, .cs_cs = r.cs_cs } } },
            .attention => |a| .{ .tag = 10, .u = .{ .attention = .{
                .dst = a.dst, .q = a.q, .k = a.k, .v = a.v, .mask = a.mask, .has_mask = @intFromBool(a.has_mask), .d_head = a.d_head, .seq_q = a.seq_q, .seq_kv = a.seq_kv, .scale = a.scale,
                .q_off = a.q_off, .k_off = a.k_off, .v_off = a.v_off, .mask_off = a.mask_off, .dst_off = a.dst_off,
                .q_rs = a.q_rs, .q_cs = a.q_cs, .k_rs = a.k_rs, .k_cs = a.k_cs, .v_rs = a.v_rs, .v_cs = a.v_cs, .mask_rs = a.mask_rs, .mask_cs = a.mask_cs, .dst_rs = a.dst_rs, .dst_cs = a.dst_cs,
            } } },
            .fused_elementwise => |f| blk: {
                const first = si;
                for (f.steps) |st| {
                    steps[si] = .{ .op = @intFromEnum(st.op), .is_swapped = @intFromBool(st.is_swapped), .secondary_buf = st.secondary_buf, .secondary_offset = st.secondary_offset };
                    si += 1;
                }
                break :blk .{ .tag = 11, .u = .{ .fused_elementwise = .{ .steps = if (f.steps.len > 0) steps[first..].ptr else null, .n_steps = f.steps.len, .n = f.n, .dst = f.dst, .src = f.src, .dst_offset = f.dst_offset, .src_offset = f.src_offset } } };
            },
        };
    }
    return .{ .ops = out, .steps = steps };
}

fn flattenIO(alloc: std.mem.Allocator, ios: []const backend_mod.ProgramIO) ![]ZgIO {
    const out = try alloc.alloc(ZgIO, ios.len);
    for (ios, out) |io, *o| o.* = .{ .buf_idx = io.buf_idx, .offset = io.offset, .host_ptr = io.host_ptr, .size = io.size };
    return out;
}

// ── backend ─────────────────────────────────────────────────────────────────
/// Per compiled program: the flattened copy of the caller's op array (kept across refreshes: only the two patched fields
/// change between tokens, src/device_inference.zig:242-256) and the RuntimeProfile handed out by get_runtime_profile.
const ProgramState = struct {
    flat: Flat,
    profile: profile_mod.RuntimeProfile = .{},
};

pub const CudaBackend = struct {
    ctx: *anyopaque,
    alloc: std.mem.Allocator,
    programs: std.AutoHashMapUnmanaged(usize, *ProgramState) = .{},

    pub fn init(alloc: std.mem.Allocator, device_ordinal: u32) !CudaBackend {
        const ctx = zg_cuda_create(@intCast(device_ordinal)) orelse {
            std.log.scoped(.cuda).err("zg_cuda_create: {s}", .{zg_cuda_last_error()});
            return error.NoCudaDevice; // no CPU fallback by design
        };
        return .{ .ctx = ctx, .alloc = alloc };
    }

    pub fn deinit(self: *CudaBackend) void {
        var it = self.programs.valueIterator();
        while (it.next()) |st| {
            st.*.flat.deinit(self.alloc);
            self.alloc.destroy(st.*);
        }
        self.programs.deinit(self.alloc);
        zg_cuda_destroy(self.ctx);
    }

    /// LlamaInferenceSession.quantizeKV (src/llama_inference.zig:648-679) for a compiled program: Q8 caches replace the f32 KV
    /// buffers, cache writes run storeColumn, attention runs attentionQuantized.  int8_query = the aarch64 branch.
    pub fn quantizeKV(self: *CudaBackend, handle: backend_mod.Backend.CompiledHandle, block_size: usize, int8_query: bool) !void {
        if (zg_cuda_program_quantize_kv(self.ctx, handle, block_size, @intFromBool(int8_query)) != 0) return error.UnsupportedCacheLayout;
    }

    /// Format hint for dense matmul operands (the tied LM head): the WGPU backend's f16 promotion (src/backend/wgpu.zig:1068-1106),
    /// opt-in and argmax-safe.  format: 1 = bf16, 2 = f16.  Returns the number of matmul ops promoted.
    pub fn promoteDenseWeights(self: *CudaBackend, handle: backend_mod.Backend.CompiledHandle, format: c_int) !usize {
        const n = zg_cuda_program_promote_dense(self.ctx, handle, format);
        if (n < 0) return error.PromotionFailed;
        return @intCast(n);
    }

    pub fn backend(self: *CudaBackend) backend_mod.Backend {
        var caps = backend_mod.Capabilities.reference_cpu; // src/backend.zig:60-70
        caps.host_visible_program_memory = false;
        caps.quantized_kv = true; // src/backend.zig:26: zg_cuda_program_quantize_kv
        return .{ .ctx = self, .vtable = &vtable, .name_str = "cuda-b200", .device_type = .cuda, .capabilities = caps };
    }
};

fn self_(ctx: *anyopaque) *CudaBackend {
    return @ptrCast(@alignCast(ctx));
}

fn denseMatMulF32(_: *anyopaque, _: backend_mod.DenseMatMulSpecF32) bool {
    return false; // host tensors are not device resident: decline like the fake backend (device_inference.zig:750-752)
}

fn compileProgram(ctx: *anyopaque, program: backend_mod.DeviceProgram) ?backend_mod.Backend.CompiledHandle {
    const be = self_(ctx);
    const flat = flatten(be.alloc, program.ops) catch return null;
    const ups = flattenIO(be.alloc, program.initial_uploads) catch { flat.deinit(be.alloc); return null; };
    defer be.alloc.free(ups);
    const qws = be.alloc.alloc(ZgQWeight, program.qweights.len) catch { flat.deinit(be.alloc); return null; };
    defer be.alloc.free(qws);
    for (program.qweights, qws) |qw, *o| o.* = .{ .data = qw.data.ptr, .n_data = qw.data.len, .scales = qw.scales.ptr, .n_scales = qw.scales.len, .rows = qw.rows, .cols = qw.cols, .block_size = qw.block_size };
    const zp = ZgProgram{
        .ops = flat.ops.ptr, .n_ops = flat.ops.len, .n_buffers = program.n_buffers, .buffer_sizes = program.buffer_sizes.ptr,
        .initial_uploads = ups.ptr, .n_uploads = ups.len, .qweights = qws.ptr, .n_qweights = qws.len,
    };
    const handle = zg_cuda_compile(be.ctx, &zp) orelse { // the library copies ops, steps and weights
        flat.deinit(be.alloc);
        return null;
    };
    // keep the flattened ops: refresh_program only patches them
    const st = be.alloc.create(ProgramState) catch { flat.deinit(be.alloc); zg_cuda_free(be.ctx, handle); return null; };
    st.* = .{ .flat = flat };
    be.programs.put(be.alloc, @intFromPtr(handle), st) catch { flat.deinit(be.alloc); be.alloc.destroy(st); zg_cuda_free(be.ctx, handle); return null; };
    return handle;
}

/// Called before every execute with the caller's mutated op array.  Only `slice_assign.dst_offset` and
/// `attention.seq_kv` change between tokens (patchSliceAssignOffset / patchAttentionSeqKV): they are copied into the
/// cached flat array; a different op count (a re-planned program) re-flattens.
fn refreshProgram(ctx: *anyopaque, handle: backend_mod.Backend.CompiledHandle, ops: []const backend_mod.DeviceOp) void {
    const be = self_(ctx);
    const st = be.programs.get(@intFromPtr(handle)) orelse return;
    if (ops.len != st.flat.ops.len) {
        const fresh = flatten(be.alloc, ops) catch return;
        st.flat.deinit(be.alloc);
        st.flat = fresh;
    } else {
        for (ops, st.flat.ops) |op, *o| switch (op) {
            .slice_assign => |s| o.u.slice_assign.dst_offset = s.dst_offset,
            .attention => |a| o.u.attention.seq_kv = a.seq_kv,
            else => {},
        };
    }
    zg_cuda_refresh(be.ctx, handle, st.flat.ops.ptr, st.flat.ops.len);
}

fn executeProgram(ctx: *anyopaque, handle: backend_mod.Backend.CompiledHandle, inputs: []const backend_mod.ProgramIO, outputs: []const backend_mod.ProgramIO) void {
    const be = self_(ctx);
    const in = flattenIO(be.alloc, inputs) catch return;
    defer be.alloc.free(in);
    const out = flattenIO(be.alloc, outputs) catch return;
    defer be.alloc.free(out);
    zg_cuda_execute(be.ctx, handle, in.ptr, in.len, out.ptr, out.len); // synchronous: outputs valid on return
}

fn freeProgram(ctx: *anyopaque, handle: backend_mod.Backend.CompiledHandle) void {
    const be = self_(ctx);
    if (be.programs.fetchRemove(@intFromPtr(handle))) |kv| {
        kv.value.flat.deinit(be.alloc);
        be.alloc.destroy(kv.value);
    }
    zg_cuda_free(be.ctx, handle);
}

/// VTable slot 6 (src/backend.zig:351): the library's ZgProfile (per-DeviceOp-tag device time + counters, filled when
/// profiling is enabled with zg_cuda_set_profiling) mapped onto RuntimeProfile (src/profile.zig:820-842).  The program-command
/// and schedule-region counters belong to the Metal planner (src/backend/program.zig) and stay zero.
fn getRuntimeProfile(ctx: *anyopaque, handle: backend_mod.Backend.CompiledHandle) ?*profile_mod.RuntimeProfile {
    const be = self_(ctx);
    const st = be.programs.get(@intFromPtr(handle)) orelse return null;
    const zp = zg_cuda_profile(be.ctx, handle) orelse return null; // profiling off: like cpu.zig:136-138
    var rp = &st.profile;
    const n = @min(rp.time_ns.len, zp.time_ns.len); // DeviceOp tags 0..11 in union declaration order on both sides
    for (0..n) |i| rp.time_ns[i] = zp.time_ns[i];
    rp.backend_op_count = zp.backend_op_count;
    rp.fallback_op_count = zp.fallback_op_count; // always 0: there is no CPU fallback
    rp.backend_dispatch_count = zp.backend_dispatch_count;
    rp.sync_time_ns = zp.sync_time_ns;
    rp.sync_count = zp.sync_count;
    rp.call_count = zp.call_count;
    return rp;
}

const vtable = backend_mod.Backend.VTable{
    .dense_matmul_f32 = denseMatMulF32,
    .compile_program = compileProgram,
    .refresh_program = refreshProgram,
    .execute_program = executeProgram,
    .free_program = freeProgram,
    .get_runtime_profile = getRuntimeProfile,
};

// The conformance test is added to src/backend/conformance.zig next to the cpu / metal / wgpu ones
// (:348-372; `runCoreCases` is file-private there):
//
//   test "cuda backend conforms to reference core ops" {
//       var cuda = cuda_mod.CudaBackend.init(std.testing.allocator, 0) catch return; // self-skips without a device
//       defer cuda.deinit();
//       try runCoreCases(cuda.backend(), 1e-5);
//   }
