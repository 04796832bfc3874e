//! flate_b200.zig -- the reference-side binding of libflate_b200.so (include/flate_b200.h).
//!
//! Drop this file into the reference checkout (ianic/flate @ 7b4d156) as `src/flate_b200.zig`, link the
//! executable with `-lflate_b200 -lcudart`, and point `src/flate.zig`, `src/gzip.zig` and `src/zlib.zig` at
//! the three facades at the bottom (INTEGRATION.md §3 shows the three-line change).  Every public name, the
//! argument meaning and the error sets are the reference's:
//!
//!   compress / compressor / Compressor.{compress, write, writer, flush, finish, setWriter}   deflate.zig:56-74,138,304-371
//!   huffman.* / store.*                                                                       deflate.zig:401-434,449-529
//!   decompress / decompressor / Inflate.{decompress, next, get, read, reader, reset, setReader}  inflate.zig:14-22,80,283-353
//!
//! The LZ77 match search, the Huffman coding, the bit packing and the inflate symbol loop run on the GPU behind the
//! C ABI; this file only moves bytes between the caller's reader/writer and the library and maps status codes to
//! Zig errors.  Written for the Zig the reference targets (0.12.0-dev: `callconv(.C)`, `std.io.Writer`);
//! there is no zig toolchain in the build image, so this file is the one part of the repository that is not
//! compiled by its test-suite -- the same entry points are exercised through ctypes by tests/ instead.
const std = @import("std");
const b200 = @This(); // this file, for qualified references from nested containers that re-declare the same names

// ---------------------------------------------------------------------------------------------------------
// extern block: every symbol include/flate_b200.h declares
// ---------------------------------------------------------------------------------------------------------
pub const Ctx = opaque {};
pub const DeflateHandle = opaque {};
pub const InflateHandle = opaque {};
pub const WriteFn = *const fn (user: ?*anyopaque, data: [*]const u8, len: usize) callconv(.C) c_int;
pub const ReadFn = *const fn (user: ?*anyopaque, buf: [*]u8, cap: usize) callconv(.C) usize;

pub extern "c" fn fb200_ctx_create(device: c_int, ctx: *?*Ctx) c_int;
pub extern "c" fn fb200_ctx_destroy(ctx: ?*Ctx) void;
pub extern "c" fn fb200_device_count() c_int;
pub extern "c" fn fb200_strerror(code: c_int) [*:0]const u8;
pub extern "c" fn fb200_last_cuda_error() [*:0]const u8;
pub extern "c" fn fb200_kernel_launches(ctx: ?*const Ctx) u64;
pub extern "c" fn fb200_ctx_set_parse_mode(ctx: ?*Ctx, mode: c_int) c_int;
pub extern "c" fn fb200_sparse_fallbacks(ctx: ?*const Ctx) u64;
pub extern "c" fn fb200_sparse_repairs(ctx: ?*const Ctx) u64;
pub extern "c" fn fb200_profile_enable(ctx: ?*Ctx, on: c_int) c_int;
pub extern "c" fn fb200_profile_phases() c_int;
pub extern "c" fn fb200_profile_phase_name(phase: c_int) [*:0]const u8;
pub extern "c" fn fb200_profile_read(ctx: ?*const Ctx, ms: [*]f64, count: [*]u64, n: c_int) c_int;
pub extern "c" fn fb200_compress_bound(n: usize, mode: c_int) usize;
pub extern "c" fn fb200_compress(ctx: ?*Ctx, container: c_int, mode: c_int, in: ?[*]const u8, n: usize, out: [*]u8, cap: usize, out_len: *usize) c_int;
pub extern "c" fn fb200_decompress(ctx: ?*Ctx, container: c_int, in: ?[*]const u8, n: usize, out: ?[*]u8, cap: usize, out_len: *usize, consumed: *usize) c_int;
pub extern "c" fn fb200_compress_device(ctx: ?*Ctx, container: c_int, mode: c_int, d_in: ?*const anyopaque, n: usize, d_out: ?*anyopaque, cap: usize, out_len: *usize, stream: ?*anyopaque) c_int;
pub extern "c" fn fb200_decompress_members_device(ctx: ?*Ctx, container: c_int, d_in: ?*const anyopaque, in_off: [*]const u64, in_len: [*]const u64, k: usize, d_out: ?*anyopaque, out_off: [*]const u64, out_cap: [*]const u64, out_len: ?[*]u64, consumed: ?[*]u64, status: ?[*]c_int, stream: ?*anyopaque) c_int;
pub extern "c" fn fb200_decompress_members(ctx: ?*Ctx, container: c_int, in: [*]const u8, in_off: [*]const u64, in_len: [*]const u64, k: usize, out: [*]u8, out_off: [*]const u64, out_cap: [*]const u64, out_len: [*]u64, consumed: ?[*]u64, status: ?[*]c_int) c_int;
pub extern "c" fn fb200_shard_align() usize;
pub extern "c" fn fb200_shard_overlap() usize;
pub extern "c" fn fb200_deflate_shard_search(ctx: ?*Ctx, level: c_int, d_in: ?*const anyopaque, n: usize, from: usize, to: usize, d_nx: ?*anyopaque, stream: ?*anyopaque) c_int;
pub extern "c" fn fb200_deflate_shard_finish(ctx: ?*Ctx, container: c_int, level: c_int, d_in: ?*const anyopaque, n: usize, d_nx: ?*const anyopaque, d_out: ?*anyopaque, cap: usize, out_len: *usize, stream: ?*anyopaque) c_int;
pub extern "c" fn fb200_simple_shard_plan(ctx: ?*Ctx, container: c_int, mode: c_int, d_in: ?*const anyopaque, shard_bytes: usize, is_last: c_int, pre_bits: *u64, has_stored: *c_int, post_bits: *u64, checksum: ?*u32, stream: ?*anyopaque) c_int;
pub extern "c" fn fb200_simple_shard_pack(ctx: ?*Ctx, start_bit: u64, d_out: ?*anyopaque, cap: usize, byte_lo: *u64, nbytes: *usize, end_bit: *u64, stream: ?*anyopaque) c_int;
pub extern "c" fn fb200_crc32_combine(crc1: u32, crc2: u32, len2: u64) u32;
pub extern "c" fn fb200_adler32_combine(adler1: u32, adler2: u32, len2: u64) u32;
pub extern "c" fn fb200_deflate_create(ctx: ?*Ctx, container: c_int, mode: c_int, w: WriteFn, user: ?*anyopaque, d: *?*DeflateHandle) c_int;
pub extern "c" fn fb200_deflate_write(d: ?*DeflateHandle, data: ?[*]const u8, n: usize) c_int;
pub extern "c" fn fb200_deflate_flush(d: ?*DeflateHandle) c_int;
pub extern "c" fn fb200_deflate_finish(d: ?*DeflateHandle) c_int;
pub extern "c" fn fb200_deflate_set_writer(d: ?*DeflateHandle, w: WriteFn, user: ?*anyopaque) void;
pub extern "c" fn fb200_deflate_destroy(d: ?*DeflateHandle) void;
pub extern "c" fn fb200_inflate_create(ctx: ?*Ctx, container: c_int, r: ReadFn, user: ?*anyopaque, s: *?*InflateHandle) c_int;
pub extern "c" fn fb200_inflate_next(s: ?*InflateHandle, data: *?[*]const u8, len: *usize) c_int;
pub extern "c" fn fb200_inflate_get(s: ?*InflateHandle, limit: usize, data: *?[*]const u8, len: *usize) c_int;
pub extern "c" fn fb200_inflate_read(s: ?*InflateHandle, buf: [*]u8, cap: usize, n: *usize) c_int;
pub extern "c" fn fb200_inflate_reset(s: ?*InflateHandle) c_int;
pub extern "c" fn fb200_inflate_set_reader(s: ?*InflateHandle, r: ReadFn, user: ?*anyopaque) void;
pub extern "c" fn fb200_inflate_rebind(s: ?*InflateHandle, r: ReadFn, user: ?*anyopaque) void;
pub extern "c" fn fb200_inflate_unused(s: ?*InflateHandle, data: *?[*]const u8, len: *usize) c_int;
pub extern "c" fn fb200_decompress_gzip_file(ctx: ?*Ctx, in: [*]const u8, n: usize, out: [*]u8, cap: usize, out_len: *usize, consumed: ?*usize, members: ?*usize) c_int;
pub const PoolHandle = opaque {};
pub extern "c" fn fb200_pool_create(device_mask: u64, pool: *?*PoolHandle) c_int;
pub extern "c" fn fb200_pool_devices(pool: ?*const PoolHandle) c_int;
pub extern "c" fn fb200_pool_destroy(pool: ?*PoolHandle) void;
pub extern "c" fn fb200_compress_batch(pool: ?*PoolHandle, container: c_int, mode: c_int, k: usize, in: [*]const [*]const u8, in_len: [*]const usize, out: [*]const [*]u8, out_cap: [*]const usize, out_len: [*]usize, status: ?[*]c_int) c_int;
pub extern "c" fn fb200_decompress_members_batch(pool: ?*PoolHandle, container: c_int, in: [*]const u8, in_off: [*]const u64, in_len: [*]const u64, k: usize, out: [*]u8, out_off: [*]const u64, out_cap: [*]const u64, out_len: [*]u64, consumed: ?[*]u64, status: ?[*]c_int) c_int;
pub extern "c" fn fb200_inflate_destroy(s: ?*InflateHandle) void;
pub extern "c" fn fb200_debug_tokens(ctx: ?*Ctx, level: c_int, in: ?[*]const u8, n: usize, tokens: [*]u32, cap: usize, ntok: *usize) c_int;
pub extern "c" fn fb200_debug_match_tables(ctx: ?*Ctx, level: c_int, in: [*]const u8, n: usize, r_full: [*]u32, r_quarter: [*]u32) c_int;
pub extern "c" fn fb200_debug_block_write(ctx: ?*Ctx, kind: c_int, tokens: ?[*]const u32, ntok: usize, eof: c_int, input: ?[*]const u8, input_len: usize, has_input: c_int, out: [*]u8, cap: usize, out_len: *usize) c_int;

// ---------------------------------------------------------------------------------------------------------
// status codes -> the reference's error sets (inflate.zig:72-78, huffman_decoder.zig:35-40, container.zig:45-51,
// bit_writer.zig:35, inflate.zig:302-304)
// ---------------------------------------------------------------------------------------------------------
pub const InflateError = error{
    EndOfStream,
    InvalidCode,
    InvalidMatch,
    InvalidBlockType,
    WrongStoredBlockNlen,
    InvalidDynamicBlockHeader,
    OversubscribedHuffmanTree,
    IncompleteHuffmanTree,
    MissingEndOfBlockCode,
    BadGzipHeader,
    BadZlibHeader,
    WrongGzipChecksum,
    WrongGzipSize,
    WrongZlibChecksum,
    InvalidState,
};
pub const DeviceError = error{ UnfinishedBits, NoSpaceLeft, InvalidArgument, CudaError, NoDevice };
pub const Error = InflateError || DeviceError;

pub fn check(rc: c_int) Error!void {
    return switch (rc) {
        0 => {},
        1 => error.EndOfStream,
        2 => error.InvalidCode,
        3 => error.InvalidMatch,
        4 => error.InvalidBlockType,
        5 => error.WrongStoredBlockNlen,
        6 => error.InvalidDynamicBlockHeader,
        7 => error.OversubscribedHuffmanTree,
        8 => error.IncompleteHuffmanTree,
        9 => error.MissingEndOfBlockCode,
        10 => error.BadGzipHeader,
        11 => error.BadZlibHeader,
        12 => error.WrongGzipChecksum,
        13 => error.WrongGzipSize,
        14 => error.WrongZlibChecksum,
        15 => error.UnfinishedBits,
        16 => error.InvalidState,
        17 => error.NoSpaceLeft,
        18 => error.InvalidArgument,
        20 => error.NoDevice,
        else => error.CudaError,
    };
}

// ---------------------------------------------------------------------------------------------------------
// one context per process (the reference keeps its 395 KB / 74.5 KB of state inside the struct; here the device
// workspace lives behind the handle).  Not safe for concurrent use, like one reference instance.
// ---------------------------------------------------------------------------------------------------------
var g_ctx: ?*Ctx = null;

/// Selects the GPU.  Optional: the first call into the library creates a context on device 0.
pub fn init(device: c_int) Error!void {
    deinit();
    try check(fb200_ctx_create(device, &g_ctx));
}
pub fn deinit() void {
    if (g_ctx) |c| fb200_ctx_destroy(c);
    g_ctx = null;
}
fn context() Error!*Ctx {
    if (g_ctx == null) try check(fb200_ctx_create(0, &g_ctx));
    return g_ctx.?;
}

pub const Container = enum(c_int) { raw = 0, gzip = 1, zlib = 2 }; // container.zig:17-21

/// deflate.zig:23-32
pub const Level = enum(c_int) {
    fast = 4,
    level_5 = 5,
    default = 6,
    level_7 = 7,
    level_8 = 8,
    best = 9,
    pub const level_4: Level = .fast;
    pub const level_6: Level = .default;
    pub const level_9: Level = .best;
};
pub const Options = struct { level: Level = .default }; // deflate.zig:18-20

const Mode = struct {
    const store: c_int = 0;
    const huffman: c_int = 1;
};

// ---------------------------------------------------------------------------------------------------------
// compression
// ---------------------------------------------------------------------------------------------------------

/// deflate.zig:56 -- same signature and behaviour.
pub fn compress(comptime container: Container, reader: anytype, writer: anytype, options: Options) !void {
    var c = try compressor(container, writer, options);
    defer c.deinit();
    try c.compress(reader);
    try c.finish();
}

/// deflate.zig:63
pub fn compressor(comptime container: Container, writer: anytype, options: Options) !Compressor(container, @TypeOf(writer)) {
    return try Compressor(container, @TypeOf(writer)).init(writer, options);
}

/// deflate.zig:71 Compressor / :121 Deflate.  `finish` must be called (deflate.zig:339-343); `deinit` releases the
/// device-side stream state (the one addition to the reference's interface).
pub fn Compressor(comptime container: Container, comptime WriterType: type) type {
    return CompressorImpl(container, WriterType, null);
}

fn CompressorImpl(comptime container: Container, comptime WriterType: type, comptime simple_mode: ?c_int) type {
    return struct {
        wrt: WriterType,
        handle: ?*DeflateHandle = null,
        write_err: ?anyerror = null, // the writer's error seen inside the callback, returned by the call that caused it

        const Self = @This();
        pub const Error = WriterType.Error || DeviceError;
        pub const Writer = std.io.Writer(*Self, Self.Error, write);

        fn onWrite(user: ?*anyopaque, data: [*]const u8, len: usize) callconv(.C) c_int {
            const self: *Self = @ptrCast(@alignCast(user.?));
            self.wrt.writeAll(data[0..len]) catch |err| {
                self.write_err = err;
                return 1;
            };
            return 0;
        }

        /// deflate.zig:138 init (writes the container header, like the reference) / :465 SimpleCompressor.init
        pub fn init(wrt: WriterType, options: Options) !Self {
            var self = Self{ .wrt = wrt };
            const mode: c_int = simple_mode orelse @intFromEnum(options.level);
            const rc = fb200_deflate_create(try context(), @intFromEnum(container), mode, onWrite, &self, &self.handle);
            try self.result(rc);
            return self;
        }
        pub fn deinit(self: *Self) void {
            if (self.handle) |h| fb200_deflate_destroy(h);
            self.handle = null;
        }

        // The struct is handed around by value like the reference's, so the callback context is refreshed before
        // every call that can reach the writer.
        fn bind(self: *Self) void {
            fb200_deflate_set_writer(self.handle, onWrite, self);
        }
        fn result(self: *Self, rc: c_int) Self.Error!void {
            if (self.write_err) |err| {
                self.write_err = null;
                return @as(WriterType.Error, @errorCast(err));
            }
            check(rc) catch |err| return switch (err) {
                error.UnfinishedBits, error.NoSpaceLeft, error.InvalidArgument, error.CudaError, error.NoDevice => |e| e,
                else => error.InvalidArgument, // inflate-side codes cannot come out of a compressor
            };
        }

        /// deflate.zig:304 -- reads `reader` to its end; the window/slide schedule of the reference is a function of the
        /// stream position only, so how the bytes arrive is not observable in the output.
        pub fn compress(self: *Self, reader: anytype) !void {
            var buf: [1 << 16]u8 = undefined;
            while (true) {
                const n = try reader.readAll(&buf);
                if (n > 0) _ = try self.write(buf[0..n]);
                if (n < buf.len) break;
            }
        }
        /// deflate.zig:363
        pub fn write(self: *Self, input: []const u8) Self.Error!usize {
            self.bind();
            try self.result(fb200_deflate_write(self.handle, input.ptr, input.len));
            return input.len;
        }
        /// deflate.zig:369
        pub fn writer(self: *Self) Writer {
            return .{ .context = self };
        }
        /// deflate.zig:335 -- completes the current block and appends the empty stored block 00 00 ff ff.
        pub fn flush(self: *Self) Self.Error!void {
            self.bind();
            try self.result(fb200_deflate_flush(self.handle));
        }
        /// deflate.zig:344 -- final block and container footer.
        pub fn finish(self: *Self) Self.Error!void {
            self.bind();
            try self.result(fb200_deflate_finish(self.handle));
        }
        /// deflate.zig:351 -- another writer, history preserved.
        pub fn setWriter(self: *Self, new_writer: WriterType) void {
            self.wrt = new_writer;
            self.bind();
        }
    };
}

fn SimpleNamespace(comptime mode: c_int) type {
    return struct {
        /// deflate.zig:402 / :421
        const NS = @This();
        pub fn compress(comptime container: Container, reader: anytype, writer: anytype) !void {
            var c = try NS.compressor(container, writer);
            defer c.deinit();
            try c.compress(reader);
            try c.finish();
        }
        /// deflate.zig:408 / :427 -- SimpleCompressor(.huffman | .store)
        pub fn Compressor(comptime container: Container, comptime WriterType: type) type {
            return CompressorImpl(container, WriterType, mode);
        }
        /// deflate.zig:412 / :431
        pub fn compressor(comptime container: Container, writer: anytype) !CompressorImpl(container, @TypeOf(writer), mode) {
            return try CompressorImpl(container, @TypeOf(writer), mode).init(writer, .{});
        }
    };
}
/// deflate.zig:401 -- Huffman coding only, no match search.
pub const huffman = SimpleNamespace(Mode.huffman);
/// deflate.zig:420 -- stored blocks only.
pub const store = SimpleNamespace(Mode.store);

// ---------------------------------------------------------------------------------------------------------
// decompression
// ---------------------------------------------------------------------------------------------------------

/// inflate.zig:14 -- same signature.
pub fn decompress(comptime container: Container, reader: anytype, writer: anytype) !void {
    var d = decompressor(container, reader);
    defer d.deinit();
    try d.decompress(writer);
}

/// inflate.zig:20
pub fn decompressor(comptime container: Container, reader: anytype) Inflate(container, @TypeOf(reader)) {
    return Inflate(container, @TypeOf(reader)).init(reader);
}

/// inflate.zig:43 Inflate.  `next`/`get` return slices borrowed from the library, valid until the next call
/// (inflate.zig:313-336); at most 65536 bytes per call like the reference's ring.
pub fn Inflate(comptime container: Container, comptime ReaderType: type) type {
    return struct {
        rdr: ReaderType,
        handle: ?*InflateHandle = null,
        read_err: ?anyerror = null,

        const Self = @This();
        pub const Error = ReaderType.Error || InflateError || DeviceError;
        pub const Reader = std.io.Reader(*Self, Self.Error, read);

        fn onRead(user: ?*anyopaque, buf: [*]u8, cap: usize) callconv(.C) usize {
            const self: *Self = @ptrCast(@alignCast(user.?));
            return self.rdr.read(buf[0..cap]) catch |err| {
                self.read_err = err; // reported by the call in progress; the library sees end of input
                return 0;
            };
        }

        /// inflate.zig:80 -- no allocation and no error here, like the reference: the handle is created on first use.
        pub fn init(rt: ReaderType) Self {
            return .{ .rdr = rt };
        }
        pub fn deinit(self: *Self) void {
            if (self.handle) |h| fb200_inflate_destroy(h);
            self.handle = null;
        }
        fn bind(self: *Self) Self.Error!void {
            if (self.handle == null) {
                try self.result(fb200_inflate_create(try context(), @intFromEnum(container), onRead, self, &self.handle));
            } else {
                fb200_inflate_rebind(self.handle, onRead, self);
            }
        }
        fn result(self: *Self, rc: c_int) Self.Error!void {
            if (self.read_err) |err| {
                self.read_err = null;
                return @as(ReaderType.Error, @errorCast(err));
            }
            try check(rc);
        }

        /// inflate.zig:283 -- replaces the inner reader; after the end of a member the next header is parsed.
        pub fn setReader(self: *Self, new_reader: ReaderType) void {
            self.rdr = new_reader;
            if (self.handle) |h| fb200_inflate_set_reader(h, onRead, self);
        }
        /// Bytes the library read from the inner reader past the end of the current member (it pulls the reader a
        /// chunk at a time; the reference's bit reader holds at most 8 such bytes, bit_reader.zig:18-44).  They are
        /// consumed by the next member after reset(); a caller that wants to go on reading the inner reader itself
        /// takes them from here first.  Valid until the next call on this decompressor.
        pub fn unreadBytes(self: *Self) []const u8 {
            var p: ?[*]const u8 = null;
            var n: usize = 0;
            if (self.handle == null or fb200_inflate_unused(self.handle, &p, &n) != 0 or p == null) return &[_]u8{};
            return p.?[0..n];
        }
        /// inflate.zig:292
        pub fn decompress(self: *Self, writer: anytype) !void {
            while (try self.next()) |buf| {
                try writer.writeAll(buf);
            }
        }
        /// inflate.zig:301 -- next member of the same reader; error.InvalidState unless the current one is at its end.
        pub fn reset(self: *Self) Self.Error!void {
            try self.bind();
            try self.result(fb200_inflate_reset(self.handle));
        }
        /// inflate.zig:313
        pub fn next(self: *Self) Self.Error!?[]const u8 {
            const out = try self.get(0);
            if (out.len == 0) return null;
            return out;
        }
        /// inflate.zig:326
        pub fn get(self: *Self, limit: usize) Self.Error![]const u8 {
            try self.bind();
            var data: ?[*]const u8 = null;
            var len: usize = 0;
            try self.result(fb200_inflate_get(self.handle, limit, &data, &len));
            if (len == 0) return &[_]u8{};
            return data.?[0..len];
        }
        /// inflate.zig:345
        pub fn read(self: *Self, buffer: []u8) Self.Error!usize {
            const out = try self.get(buffer.len);
            @memcpy(buffer[0..out.len], out);
            return out.len;
        }
        /// inflate.zig:351
        pub fn reader(self: *Self) Reader {
            return .{ .context = self };
        }
    };
}

// ---------------------------------------------------------------------------------------------------------
// the three facades: src/flate.zig (raw), src/gzip.zig, src/zlib.zig keep their public names by re-exporting these
//     pub usingnamespace b200.gzip;        // src/gzip.zig:5-66
// ---------------------------------------------------------------------------------------------------------
fn Facade(comptime container: Container) type {
    return struct {
        pub const Options = b200.Options;

        pub fn decompress(reader: anytype, writer: anytype) !void {
            try b200.decompress(container, reader, writer);
        }
        pub fn Decompressor(comptime ReaderType: type) type {
            return Inflate(container, ReaderType);
        }
        pub fn decompressor(reader: anytype) Inflate(container, @TypeOf(reader)) {
            return b200.decompressor(container, reader);
        }
        pub fn compress(reader: anytype, writer: anytype, options: b200.Options) !void {
            try b200.compress(container, reader, writer, options);
        }
        pub fn Compressor(comptime WriterType: type) type {
            return CompressorImpl(container, WriterType, null);
        }
        pub fn compressor(writer: anytype, options: b200.Options) !CompressorImpl(container, @TypeOf(writer), null) {
            return try b200.compressor(container, writer, options);
        }
        pub const huffman = struct {
            pub fn compress(reader: anytype, writer: anytype) !void {
                try SimpleNamespace(Mode.huffman).compress(container, reader, writer);
            }
            pub fn Compressor(comptime WriterType: type) type {
                return CompressorImpl(container, WriterType, Mode.huffman);
            }
            pub fn compressor(writer: anytype) !CompressorImpl(container, @TypeOf(writer), Mode.huffman) {
                return try SimpleNamespace(Mode.huffman).compressor(container, writer);
            }
        };
        pub const store = struct {
            pub fn compress(reader: anytype, writer: anytype) !void {
                try SimpleNamespace(Mode.store).compress(container, reader, writer);
            }
            pub fn Compressor(comptime WriterType: type) type {
                return CompressorImpl(container, WriterType, Mode.store);
            }
            pub fn compressor(writer: anytype) !CompressorImpl(container, @TypeOf(writer), Mode.store) {
                return try SimpleNamespace(Mode.store).compressor(container, writer);
            }
        };
    };
}
pub const flate = Facade(.raw); // src/flate.zig:10-71
pub const gzip = Facade(.gzip); // src/gzip.zig:5-66
pub const zlib = Facade(.zlib); // src/zlib.zig:5-66

// The reference's own round-trip test (bin/roundtrip.zig:14-75 behaviour) against this binding; runs where a zig
// toolchain and a GPU are present: `zig test src/flate_b200.zig -lc -lflate_b200 -lcudart`.
test "gzip round trip through the GPU library" {
    const data = "Hello world\n" ** 1000;
    var compressed = std.ArrayList(u8).init(std.testing.allocator);
    defer compressed.deinit();
    var in = std.io.fixedBufferStream(data);
    try gzip.compress(in.reader(), compressed.writer(), .{});
    var plain = std.ArrayList(u8).init(std.testing.allocator);
    defer plain.deinit();
    var cin = std.io.fixedBufferStream(compressed.items);
    try gzip.decompress(cin.reader(), plain.writer());
    try std.testing.expectEqualSlices(u8, data, plain.items);
}
