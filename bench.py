#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 DEFLATE engine (contract in the task statement).

One JSON line on stdout.  Its top level is BASELINE.json configs[1] (C2): raw deflate level 6 over 256 MiB of
enwik-like synthetic bytes per GPU; a "step" is one pass of the hot path over that batch.  With --gpus N (under
torchrun) every rank compresses its own independent 256 MiB chunk (the path shards by chunk: weak scaling) and the
per-shard outputs are all-gathered over NCCL, overlapped with the next step.  The other BASELINE configs ride in the
same line, each with its own value / e2e / roofline / cpu_baseline:

  inflate       C3: 1 GiB of plain output as 1 MiB gzip members, split over the ranks (strong scaling)
  c4_level9     C4: raw deflate level 9 (chain 4096) over the same 256 MiB
  c5_huffman    C5: huffman-only over ONE 4 GiB random+zeros stream; with N > 1 the stream is sharded by 65535-byte
                block ranges (exclusive scan of shard bit sizes, OR-merged boundary bytes, NCCL all-gather)
  mixed         tar-like input (text, 512-byte aligned zero padding, random binary): the sparse parse off its happy path
  single_stream N > 1: ONE level-6 stream sharded by position over all ranks

  value        : throughput with inputs resident in HBM (CUDA events on the launching stream)
  e2e          : same metric through the host-buffer C-ABI call (pinned host -> H2D -> kernels -> D2H inside the timing)
  roofline     : dominant kernel (largest live CUDA-event phase): algorithmic bytes / its time vs the measured HBM
                 peak; traffic = dram bytes per launch from the ncu --set full capture named in profiles/traffic.json
  cpu_baseline : the CPU oracle (port of the reference's algorithm; the reference is Zig, no zig toolchain here) on
                 this box's host cores, single thread like the reference, output compared byte for byte

--impl reference times the reference's CPU implementation of the C2 path (the oracle port) on the whole 256 MiB.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Keep stdout to the one JSON line WITHOUT switching NCCL's log off: everything that writes to file descriptor 1
# (NCCL_DEBUG=INFO prints there) goes to stderr, the JSON line goes to the original stdout.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)
sys.stdout = sys.stderr

MIB = 1 << 20
WORKLOAD_BYTES = 256 * MIB
MEMBER_BYTES = 1 * MIB
INFLATE_TOTAL = 1024 * MIB
C5_BYTES = 4096 * MIB
LEVEL = 6


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_of(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernel/config, from the ncu --set full
    captures of the committed state (profiles/traffic.json names the capture files)."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(tp)).get(key)
    except Exception:
        return None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML, every ~5 ms)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index=0):
        self.index = index
        self.sm = []
        self.mask = 0
        self.max_mhz = None
        self.stop_flag = False
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.nv is None:
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.stop_flag = True
        self.thread.join(timeout=2)
        sm = sorted(self.sm)
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm)}


def make_text(nbytes, rank):
    from flate_b200 import synth
    return synth.enwik_like(nbytes, seed=0x5EED0001 + 7919 * rank)


def make_tar_like(nbytes, seed=0x7A7):
    """Tar-like stream: members of text or random binary with a 512-byte header block and zero padding to a multiple
    of 512 bytes, plus the occasional long run of zero blocks (sparse files, the end-of-archive marker)."""
    import numpy as np
    from flate_b200 import synth
    out = np.zeros(nbytes, dtype=np.uint8)
    text = synth.enwik_like(min(nbytes, 64 * MIB), seed=seed)
    r = synth.splitmix64(seed, 1 << 16)
    pos, i = 0, 0
    while pos + 1024 < nbytes:
        kind = int(r[i % 65536] % 8)
        size = int(200 + r[(i + 1) % 65536] % (1 << (10 + int(r[(i + 2) % 65536] % 10))))
        i += 3
        hdr = text[(pos // 7) % (text.size - 512):][:100]
        out[pos: pos + 100] = hdr                       # name / mode / size fields; the rest of the header block is zero
        pos += 512
        size = min(size, nbytes - pos)
        if kind < 5:
            o = int(r[i % 65536] % max(1, text.size - size))
            out[pos: pos + size] = text[o: o + size]
        elif kind < 7:
            out[pos: pos + size] = synth.splitmix64(seed + i, (size + 7) // 8).view(np.uint8)[:size]
        # kind 7: a hole (zeros)
        pos += (size + 511) // 512 * 512
    return out


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def oracle_lib():
    from oracle import oracle as o
    try:
        return o, o.lib(o.build(native=True))
    except Exception:
        return o, o.lib()


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port) on the box's host cores, on the SAME configuration as
    our arm: the whole 256 MiB chunk of every rank per step.  One stream is one thread (the reference has no threads,
    SURVEY.md §2); the N-GPU workload is N independent chunks, which rank 0 runs side by side on min(N, host cores)
    threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    o, lib = oracle_lib()
    world = max(1, args.gpus)
    threads = max(1, min(world, host_threads()))
    n = args.bytes
    chunks = [make_text(n, r) for r in range(world)]
    times = []
    out_len = 0

    def one(d):
        return len(o.compress(d, o.RAW, LEVEL, _lib_override=lib))   # ctypes releases the GIL for the call

    with ThreadPoolExecutor(max_workers=threads) as pool:
        for i in range(args.warmup + args.steps):
            t = time.perf_counter()
            lens = list(pool.map(one, chunks))
            dt = time.perf_counter() - t
            out_len = lens[0]
            if i >= args.warmup:
                times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = world * n / 1e6 / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "deflate L6 MB/s in", "value": round(value, 2), "unit": "MB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "raw deflate level %d, %d MiB enwik-like synthetic bytes per GPU (BASELINE configs[1])"
                               % (LEVEL, n // MIB),
                   "ratio": round(n / out_len, 3), "compressed_bytes": out_len,
                   "threads": "%d chunk(s) on %d host thread(s)" % (world, threads)},
        "cpu_baseline": {"value": round(value, 2), "unit": "MB/s", "cores": threads, "kind": "port",
                         "sample": "the whole %d MiB chunk of every rank per step; oracle/flate_oracle.c -O3 -march=native"
                                   % (n // MIB)},
        "e2e": {"value": round(value, 2), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bytes", type=int, default=WORKLOAD_BYTES, help="per-GPU chunk size (default 256 MiB)")
    ap.add_argument("--level", type=int, default=LEVEL)
    ap.add_argument("--c5-bytes", type=int, default=C5_BYTES, help="size of the huffman-only stream (default 4 GiB)")
    ap.add_argument("--skip-inflate", action="store_true")
    ap.add_argument("--skip-c4", action="store_true")
    ap.add_argument("--skip-c5", action="store_true")
    ap.add_argument("--skip-mixed", action="store_true")
    ap.add_argument("--skip-stream", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import flate_b200
    from flate_b200 import sharding, synth
    ctx = flate_b200.Context(local_rank)
    lib = ctx.lib
    n = args.bytes
    level = args.level
    peak, peak_src = peaks()
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, warmup=0):
        """ms per step of fn on the device (CUDA events on the launching stream, max over ranks)."""
        for _ in range(warmup):
            fn()
        barrier()
        ev0.record(stream)
        for _ in range(steps):
            fn()
        ev1.record(stream)
        barrier()
        return max_over_ranks(ev0.elapsed_time(ev1) / steps)

    def timed_host(fn, steps, warmup=1):
        """ms per step of a host-buffer call (wall clock around calls that return after their last copy)."""
        for _ in range(warmup):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        return max_over_ranks((time.perf_counter() - t0) * 1e3 / steps)

    def roofline_of(phases, algo_bytes, traffic_key):
        dom = max((k for k in phases if phases[k][1]), key=lambda k: phases[k][0])
        dom_ms = phases[dom][0] / phases[dom][1]
        return {"bound": "hbm", "achieved": round(algo_bytes / 1e9 / (dom_ms / 1e3), 2), "peak": peak, "unit": "GB/s",
                "frac": round(algo_bytes / 1e9 / (dom_ms / 1e3) / peak, 5), "traffic": traffic_of(traffic_key),
                "traffic_key": traffic_key, "kernel": dom, "kernel_ms": round(dom_ms, 3), "peak_source": peak_src,
                "phases_ms": {k: round(v[0] / max(1, v[1]), 3) for k, v in phases.items() if v[1]}}

    text = make_text(n, rank)                       # this rank's independent chunk
    h_in = torch.from_numpy(text).pin_memory()
    d_in = h_in.cuda(non_blocking=True)
    cap = lib.fb200_compress_bound(n, level) + 64
    d_out = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
    h_out = torch.empty(cap + 64, dtype=torch.uint8).pin_memory()

    # =====================================================================================================
    # C2: raw deflate level 6, 256 MiB per GPU (the headline line)
    # =====================================================================================================
    d_outs = [d_out, torch.empty_like(d_out)] if world > 1 else [d_out]
    gather_bufs, pending = [], [None, None]
    out_len = ctx.compress_device(d_in.data_ptr(), n, d_out.data_ptr(), cap, mode=level, stream=sp)
    pad = 0
    if world > 1:
        # per-shard outputs are all-gathered (north star); pad to a common size agreed on once
        allsz = sharding.all_gather_sizes(out_len, d_out.device)
        pad = (int(max(allsz) * 1.02) + 4096) // 256 * 256
        gather_bufs = [torch.empty(world * pad, dtype=torch.uint8, device="cuda") for _ in range(2)]
    step_no = [0]
    last_len = [out_len]

    def full_step():
        # double-buffered: the all-gather of step k runs on NCCL's stream while step k+1 compresses
        i = (step_no[0] & 1) if world > 1 else 0
        step_no[0] += 1
        if pending[i] is not None:
            pending[i].wait()
            pending[i] = None
        buf = d_outs[i]
        last_len[0] = ctx.compress_device(d_in.data_ptr(), n, buf.data_ptr(), cap, mode=level, stream=sp)
        if world > 1:
            pending[i] = dist.all_gather_into_tensor(gather_bufs[i], buf[:pad], async_op=True)

    def drain_gathers():
        for i in range(2):
            if pending[i] is not None:
                pending[i].wait()
                pending[i] = None

    for _ in range(args.warmup):
        full_step()
    drain_gathers()
    ctx.profile(True)
    launches0 = ctx.kernel_launches
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev0.record(stream)
    for _ in range(args.steps):
        full_step()
    drain_gathers()  # the stream now waits for the last all-gathers: they are inside the timed region
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_step = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
    launches = ctx.kernel_launches - launches0
    phases = ctx.profile_read()
    ctx.profile(False)
    out_len = last_len[0]
    value = world * n / 1e6 / (ms_step / 1e3)

    def e2e_compress(h_src, nbytes, h_dst, dst_cap, mode):
        ln = C.c_size_t(0)
        rc = lib.fb200_compress(ctx.h, flate_b200.RAW, mode, h_src.data_ptr(), nbytes, h_dst.data_ptr(), dst_cap, C.byref(ln))
        if rc:
            raise RuntimeError("fb200_compress failed: %d" % rc)
        return ln.value

    e2e_len = [0]

    def e2e_step():
        e2e_len[0] = e2e_compress(h_in, n, h_out, cap, level)

    e2e_ms = timed_host(e2e_step, args.steps, warmup=2)
    e2e_value = world * n / 1e6 / (e2e_ms / 1e3)
    assert e2e_len[0] == out_len
    compressed = h_out[:out_len].numpy().tobytes()
    roofline = roofline_of(phases, n + out_len, "L6:sparse_parse")     # N_in + N_out per launch

    cpu = None
    if rank == 0 and not args.skip_cpu:
        o, olib = oracle_lib()
        t0 = time.perf_counter()
        want = o.compress(text, o.RAW, level, _lib_override=olib)
        dt = time.perf_counter() - t0
        assert want == compressed, "GPU output differs from the CPU oracle"
        cpu = {"value": round(n / 1e6 / dt, 2), "unit": "MB/s", "cores": 1, "kind": "port",
               "sample": "the whole %d MiB rank-0 chunk, one pass; oracle/flate_oracle.c (-O3 -march=native); output compared "
                         "byte for byte with the GPU's: identical; box has %d host threads, the reference is single-threaded"
                         % (n // MIB, host_threads())}

    # ---- one stream over all GPUs: position-sharded search + all-gather of the lazy-step tables ----
    single = None
    if world > 1:
        d_stream = d_in.clone()
        dist.broadcast(d_stream, src=0)                      # every rank works on rank 0's stream
        m = [0]

        def single_step():
            m[0] = sharding.compress_stream_sharded(ctx, d_stream, n, d_out, level=level)

        sms = timed(single_step, args.steps, warmup=2)
        if rank == 0:
            same = m[0] == len(compressed) and d_out[:m[0]].cpu().numpy().tobytes() == compressed
            single = {"metric": "deflate L6 MB/s in, ONE %d MiB stream sharded by position over %d GPUs" % (n // MIB, world),
                      "value": round(n / 1e6 / (sms / 1e3), 1), "unit": "MB/s", "ms_per_step": round(sms, 3),
                      "scaling": "strong", "identical_to_one_gpu_stream": bool(same),
                      "collective": "NCCL all-gather of the 4 B/position lazy-step tables"}
        del d_stream

    # =====================================================================================================
    # C3: inflate, 1 MiB gzip members, 1 GiB of plain output over all ranks
    # =====================================================================================================
    inflate = None
    if not args.skip_inflate:
        total = INFLATE_TOTAL // world
        nmem = max(1, total // MEMBER_BYTES)
        uniq = min(nmem, n // MEMBER_BYTES)
        # Members are level-6 gzip streams of 1 MiB slices of the text.  The reference's inflate rejects code-length runs
        # that cross the literal/distance boundary (inflate.zig:161-170) although its own block writer emits them
        # (block_writer.zig:78-171); we reproduce that, so such members (about 1 in 40) are re-cut from a shifted slice
        # until the stream is one the reference itself accepts.
        members, los = [], []
        for i in range(uniq):
            for shift in range(0, 64):
                lo = (i * MEMBER_BYTES + shift * 4099) % (n - MEMBER_BYTES + 1)
                mb = ctx.compress(text[lo:lo + MEMBER_BYTES], flate_b200.GZIP, LEVEL)
                try:
                    ctx.decompress(mb, flate_b200.GZIP, cap=MEMBER_BYTES + 64)
                    break
                except flate_b200.FlateError:
                    continue
            members.append(mb)
            los.append(lo)

        def inflate_leg(nmem):
            blob = np.frombuffer(b"".join(members[i % uniq] for i in range(nmem)), dtype=np.uint8)
            lens = np.array([len(members[i % uniq]) for i in range(nmem)], dtype=np.uint64)
            offs = np.zeros(nmem, dtype=np.uint64)
            offs[1:] = np.cumsum(lens)[:-1]
            h_blob = torch.from_numpy(blob.copy()).pin_memory()
            d_blob = h_blob.cuda()
            d_plain = torch.empty(nmem * MEMBER_BYTES + 64, dtype=torch.uint8, device="cuda")
            h_plain = torch.empty(nmem * MEMBER_BYTES + 64, dtype=torch.uint8).pin_memory()
            ooff = np.arange(nmem, dtype=np.uint64) * np.uint64(MEMBER_BYTES)
            ocap = np.full(nmem, MEMBER_BYTES, dtype=np.uint64)
            got = [0]

            def dev_step():
                rc, ol, used, st = ctx.decompress_members_device(d_blob.data_ptr(), offs, lens, d_plain.data_ptr(), ooff, ocap,
                                                                 flate_b200.GZIP, stream=sp)
                if rc:
                    raise RuntimeError("inflate failed: %d" % rc)
                got[0] = int(ol.sum())

            ol_h = np.zeros(nmem, dtype=np.uint64)
            used_h = np.zeros(nmem, dtype=np.uint64)
            st_h = np.zeros(nmem, dtype=np.int32)

            def host_step():
                rc = lib.fb200_decompress_members(ctx.h, flate_b200.GZIP, h_blob.data_ptr(), offs.ctypes.data, lens.ctypes.data,
                                                  nmem, h_plain.data_ptr(), ooff.ctypes.data, ocap.ctypes.data, ol_h.ctypes.data,
                                                  used_h.ctypes.data, st_h.ctypes.data)
                if rc:
                    raise RuntimeError("fb200_decompress_members failed: %d" % rc)

            isteps = max(2, args.steps)
            dev_step()
            ctx.profile(True)
            ims = timed(dev_step, isteps, warmup=2)
            iph = ctx.profile_read()
            ctx.profile(False)
            kms = iph["inflate_members"][0] / max(1, iph["inflate_members"][1])
            # parity: EVERY member's plain bytes equal the text slice it was made from (compared on the device)
            same = True
            for i in range(nmem):
                lo = los[i % uniq]
                same = same and bool(torch.equal(d_plain[i * MEMBER_BYTES:(i + 1) * MEMBER_BYTES], d_in[lo:lo + MEMBER_BYTES]))
            assert same, "inflate output differs from the original text"
            hms = timed_host(host_step, isteps, warmup=1)
            assert int(ol_h.sum()) == got[0] and bool(torch.equal(h_plain[:MEMBER_BYTES], h_in[los[0]:los[0] + MEMBER_BYTES]))
            return ims, hms, got[0], int(blob.size), kms

        ims, hms, plain_bytes, blob_len, kms = inflate_leg(nmem)
        algo = float(blob_len + plain_bytes)
        inflate = {"metric": "inflate MB/s out", "value": round(world * plain_bytes / 1e6 / (ims / 1e3), 1), "unit": "MB/s",
                   "ms_per_step": round(ims, 3), "scaling": "strong", "higher_is_better": True,
                   "workload": "%d gzip members x 1 MiB plain (level 6 text) per GPU, %d MiB plain over %d GPU(s) (BASELINE configs[2]); "
                               "every member compared with its source text on the device" % (nmem, world * plain_bytes // MIB, world),
                   "e2e": {"value": round(world * plain_bytes / 1e6 / (hms / 1e3), 1), "unit": "MB/s", "ms_per_step": round(hms, 3),
                           "h2d_bytes_per_step": blob_len, "d2h_bytes_per_step": plain_bytes,
                           "call": "fb200_decompress_members (pinned host buffers, copies pipelined in 8 batches)"},
                   "roofline": {"bound": "hbm", "achieved": round(algo / 1e9 / (kms / 1e3), 2), "peak": peak, "unit": "GB/s",
                                "frac": round(algo / 1e9 / (kms / 1e3) / peak, 5), "traffic": traffic_of("inflate:%d" % nmem),
                                "traffic_key": "inflate:%d" % nmem, "kernel": "inflate_members_par_kernel", "kernel_ms": round(kms, 3)}}
        if rank == 0 and not args.skip_cpu:
            o, olib = oracle_lib()
            t0 = time.perf_counter()
            for i in range(uniq):
                pl, _ = o.decompress(members[i], o.GZIP, cap=MEMBER_BYTES + 64)
                assert len(pl) == MEMBER_BYTES
            dt = time.perf_counter() - t0
            inflate["cpu_baseline"] = {"value": round(uniq * MEMBER_BYTES / 1e6 / dt, 1), "unit": "MB/s", "cores": 1, "kind": "port",
                                       "sample": "all %d distinct members, oracle inflate, 1 thread" % uniq}
        # one member alone (latency of a single stream)
        rc1 = [0]
        one_off, one_len = np.zeros(1, dtype=np.uint64), np.array([len(members[0])], dtype=np.uint64)
        one_cap = np.array([MEMBER_BYTES], dtype=np.uint64)
        d_one = torch.from_numpy(np.frombuffer(members[0], dtype=np.uint8).copy()).cuda()
        d_one_out = torch.empty(MEMBER_BYTES + 64, dtype=torch.uint8, device="cuda")

        def one_step():
            rc1[0] = ctx.decompress_members_device(d_one.data_ptr(), one_off, one_len, d_one_out.data_ptr(), one_off, one_cap,
                                                   flate_b200.GZIP, stream=sp)[0]

        oms = timed(one_step, 5, warmup=2)
        inflate["single_member"] = {"value": round(MEMBER_BYTES / 1e6 / (oms / 1e3), 1), "unit": "MB/s", "ms": round(oms, 3),
                                    "workload": "one 1 MiB gzip member alone on the GPU"}
        if world > 1:
            wn = max(1, INFLATE_TOTAL // MEMBER_BYTES)
            wms, whms, wplain, _, wk = inflate_leg(wn)
            inflate["weak"] = {"value": round(world * wplain / 1e6 / (wms / 1e3), 1), "unit": "MB/s", "ms_per_step": round(wms, 3),
                               "scaling": "weak", "kernel_ms": round(wk, 3), "workload": "%d gzip members x 1 MiB plain per GPU" % wn}

    # =====================================================================================================
    # C4: raw deflate level 9 on the same 256 MiB (deep hash-chain walk)
    # =====================================================================================================
    c4 = None
    if not args.skip_c4:
        l9 = [0]

        def l9_step():
            l9[0] = ctx.compress_device(d_in.data_ptr(), n, d_out.data_ptr(), cap, mode=9, stream=sp)

        l9_step()
        ctx.profile(True)
        ms9 = timed(l9_step, max(2, args.steps // 2), warmup=1)
        ph9 = ctx.profile_read()
        ctx.profile(False)
        e9 = [0]

        def l9_host():
            e9[0] = e2e_compress(h_in, n, h_out, cap, 9)

        hms9 = timed_host(l9_host, max(2, args.steps // 2), warmup=1)
        assert e9[0] == l9[0]
        c4 = {"metric": "deflate L9 MB/s in", "value": round(world * n / 1e6 / (ms9 / 1e3), 1), "unit": "MB/s",
              "ms_per_step": round(ms9, 3), "scaling": "weak", "higher_is_better": True,
              "workload": "raw deflate level 9 (good 32, lazy 258, nice 258, chain 4096), %d MiB enwik-like per GPU (BASELINE configs[3])"
                          % (n // MIB),
              "ratio": round(n / l9[0], 3),
              "e2e": {"value": round(world * n / 1e6 / (hms9 / 1e3), 1), "unit": "MB/s", "ms_per_step": round(hms9, 3),
                      "h2d_bytes_per_step": n, "d2h_bytes_per_step": l9[0]},
              "roofline": roofline_of(ph9, n + l9[0], "L9:sparse_parse")}
        if rank == 0 and not args.skip_cpu:
            o, olib = oracle_lib()
            t0 = time.perf_counter()
            want9 = o.compress(text, o.RAW, 9, _lib_override=olib)
            dt = time.perf_counter() - t0
            assert want9 == h_out[:e9[0]].numpy().tobytes(), "level 9: GPU output differs from the CPU oracle"
            c4["cpu_baseline"] = {"value": round(n / 1e6 / dt, 2), "unit": "MB/s", "cores": 1, "kind": "port",
                                  "sample": "the whole %d MiB chunk, one pass; output compared byte for byte: identical" % (n // MIB)}

        # worst case of SURVEY.md section 8(d): a low-entropy 4-gram alphabet, so that almost every 4-byte hash has been
        # seen hundreds of times inside the window and the chain walk goes to its full depth (rank 0, N = 1 only)
        if world == 1:
            wn = min(n, 32 * MIB)
            rng = np.random.default_rng(0x5EED0009)
            low = rng.integers(0, 4, wn, dtype=np.uint8) + ord("a")
            d_low = torch.from_numpy(low).cuda()
            lw = [0]

            def worst_step():
                lw[0] = ctx.compress_device(d_low.data_ptr(), wn, d_out.data_ptr(), cap, mode=9, stream=sp)

            msw = timed(worst_step, 2, warmup=1)
            c4["worst_case"] = {"metric": "deflate L9 MB/s in (low-entropy input, full-depth chain walks)", "value": round(wn / 1e6 / (msw / 1e3), 1),
                                "unit": "MB/s", "ms_per_step": round(msw, 3), "ratio": round(wn / lw[0], 3),
                                "workload": "%d MiB of uniformly random bytes over a 4-letter alphabet, level 9" % (wn // MIB)}
            if not args.skip_cpu:
                import zlib
                comp_w = d_out[:lw[0]].cpu().numpy().tobytes()
                assert zlib.decompress(comp_w, -15) == low.tobytes(), "worst case: the stream does not inflate back"
                o, olib = oracle_lib()
                samp = 2 * MIB
                t0 = time.perf_counter()
                want_w = o.compress(low[:samp], o.RAW, 9, _lib_override=olib)
                dt = time.perf_counter() - t0
                got_w = ctx.compress(low[:samp], flate_b200.RAW, 9)
                assert got_w == want_w, "worst case: GPU output differs from the CPU oracle on the sample"
                c4["worst_case"]["cpu_baseline"] = {"value": round(samp / 1e6 / dt, 2), "unit": "MB/s", "cores": 1, "kind": "port",
                                                    "sample": "first %d MiB as a stream of its own: output identical; the whole stream inflates back (zlib)" % (samp // MIB)}
            del d_low

    # =====================================================================================================
    # mixed: tar-like input, the sparse parse off its happy path (level 6)
    # =====================================================================================================
    mixed = None
    if not args.skip_mixed:
        tar = make_tar_like(n)
        d_tar = torch.from_numpy(tar).cuda()
        mx = [0]
        rep0, fb0 = ctx.sparse_repairs, ctx.sparse_fallbacks

        def mixed_step():
            mx[0] = ctx.compress_device(d_tar.data_ptr(), n, d_out.data_ptr(), cap, mode=level, stream=sp)

        mms = timed(mixed_step, max(2, args.steps // 2), warmup=1)
        calls = 1 + max(2, args.steps // 2)
        mixed = {"metric": "deflate L6 MB/s in (tar-like input)", "value": round(world * n / 1e6 / (mms / 1e3), 1), "unit": "MB/s",
                 "ms_per_step": round(mms, 3), "ratio": round(n / mx[0], 3),
                 "workload": "%d MiB tar-like: text and random members, 512-byte headers, zero padding and holes" % (n // MIB),
                 "sparse_repairs_per_call": round((ctx.sparse_repairs - rep0) / calls, 2),
                 "sparse_fallbacks_per_call": round((ctx.sparse_fallbacks - fb0) / calls, 2)}
        if rank == 0 and not args.skip_cpu:
            o, olib = oracle_lib()
            t0 = time.perf_counter()
            wantm = o.compress(tar, o.RAW, level, _lib_override=olib)
            dt = time.perf_counter() - t0
            assert wantm == d_out[:mx[0]].cpu().numpy().tobytes(), "tar-like: GPU output differs from the CPU oracle"
            mixed["cpu_baseline"] = {"value": round(n / 1e6 / dt, 2), "unit": "MB/s", "cores": 1, "kind": "port",
                                     "sample": "the whole stream, one pass; output identical"}
        del d_tar

    # =====================================================================================================
    # streaming: the same bytes through Compressor.write in 1 MiB pieces
    # =====================================================================================================
    streaming = None
    if world == 1 and not args.skip_stream:
        import psutil
        from flate_b200 import _lib as fb_lib
        sink = torch.zeros(4 * cap + 64, dtype=torch.uint8)   # where the writer puts what it is handed (touched: no page faults later)
        sink_ptr, sink_len = sink.data_ptr(), [0]

        def on_write(_user, data, nbytes):
            C.memmove(sink_ptr + sink_len[0], data, nbytes)
            sink_len[0] += nbytes
            return 0

        cb = fb_lib.WRITE_FN(on_write)
        piece = MIB

        def stream_once(passes):
            sink_len[0] = 0
            h = C.c_void_p()
            rc = lib.fb200_deflate_create(ctx.h, flate_b200.RAW, level, cb, None, C.byref(h))
            assert rc == 0, rc
            for _ in range(passes):
                for pos in range(0, n, piece):
                    rc = lib.fb200_deflate_write(h, h_in.data_ptr() + pos, min(piece, n - pos))
                    assert rc == 0, rc
            rc = lib.fb200_deflate_finish(h)
            assert rc == 0, rc
            lib.fb200_deflate_destroy(h)
            return sink_len[0]

        got1 = stream_once(1)   # warm-up, and the parity check: the stream equals the one-shot stream (itself equal to the oracle's)
        same = got1 == out_len and sink[:got1].numpy().tobytes() == compressed
        assert same, "streaming compressor: output differs from the one-shot stream"
        proc = psutil.Process()
        rss0 = proc.memory_info().rss
        passes = max(1, (1 << 30) // n)
        t0 = time.perf_counter()
        got4 = stream_once(passes)
        dt = time.perf_counter() - t0
        rss1 = proc.memory_info().rss
        streaming = {"metric": "deflate L6 MB/s in, Compressor.write in 1 MiB pieces", "value": round(passes * n / 1e6 / dt, 1), "unit": "MB/s",
                     "workload": "%d MiB (the C2 text %d times) through fb200_deflate_write in 1 MiB pieces from pinned host memory, "
                                 "output handed to a writer callback that copies it away" % (passes * n // MIB, passes),
                     "compressed_bytes": got4, "fraction_of_one_shot_e2e": round(passes * n / 1e6 / dt / e2e_value, 3),
                     "host_rss_growth_mib": round((rss1 - rss0) / MIB, 1),
                     "identical_to_one_shot_stream": bool(same), "checked_on": "%d MiB, one pass" % (n // MIB)}
        del sink

    # free the C2 buffers before the 4 GiB leg
    del d_outs, gather_bufs

    # =====================================================================================================
    # C5: huffman-only over ONE 4 GiB random+zeros stream; N > 1: sharded by 65535-byte block ranges
    # =====================================================================================================
    c5 = None
    if not args.skip_c5:
        n5 = args.c5_bytes
        ranges = sharding.simple_shard_ranges(n5, world)
        lo5, hi5 = ranges[rank]
        # every rank generates only its own range of the stream (the generator is position-addressable in 256 MiB pieces)
        piece = 256 * MIB

        def gen_range(lo, hi):
            out = np.empty(hi - lo, dtype=np.uint8)
            p = lo
            while p < hi:
                k = p // piece
                blk = synth.random_zero_mix(min(piece, n5 - k * piece), seed=0x5EED0005 + k)
                a, b = p - k * piece, min(hi, (k + 1) * piece) - k * piece
                out[p - lo: p - lo + (b - a)] = blk[a:b]
                p += b - a
            return out

        mine = gen_range(lo5, hi5)
        h5 = torch.from_numpy(mine).pin_memory()
        d5 = h5.cuda()
        sh_bytes = hi5 - lo5
        cap5 = lib.fb200_compress_bound(sh_bytes, 1) + 64
        out5 = [None, 0]
        if world == 1:
            d5_out = torch.empty(cap5 + 64, dtype=torch.uint8, device="cuda")

            def c5_step():
                out5[1] = ctx.compress_device(d5.data_ptr(), n5, d5_out.data_ptr(), cap5, mode=1, stream=sp)
                out5[0] = d5_out
        else:
            # every rank builds its own copy of the whole stream: its shard packed in place, the others' by broadcast
            full5 = torch.empty((lib.fb200_compress_bound(n5, 1) + 64 + 255) // 256 * 256, dtype=torch.uint8, device="cuda")

            def c5_step():
                out5[0], out5[1] = sharding.compress_simple_sharded(ctx, d5, lo5, hi5, n5, mode=1, container=0, out=full5)

            def c5_step_in_place():
                sharding.compress_simple_sharded(ctx, d5, lo5, hi5, n5, mode=1, container=0, out=full5, gather=False)

        c5_step()
        ctx.profile(True)
        ms5 = timed(c5_step, max(2, args.steps // 2), warmup=1)
        ph5 = ctx.profile_read()
        ctx.profile(False)
        total5 = out5[1]
        ms5_in_place = None
        if world > 1:
            c5_step_in_place()
            ms5_in_place = timed(c5_step_in_place, max(2, args.steps // 2), warmup=1)
            c5_step()   # (the parity check below reads the gathered stream)
        # e2e: N = 1 through fb200_compress with pinned host buffers; N > 1: host shard -> device, sharded compress
        # (all-gather included), this rank's share of the stream back to the host
        if world == 1:
            h5_out = torch.empty(cap5 + 64, dtype=torch.uint8).pin_memory()
            e5 = [0]

            def c5_host():
                e5[0] = e2e_compress(h5, n5, h5_out, cap5, 1)

            hms5 = timed_host(c5_host, 2, warmup=1)
            assert e5[0] == total5
            d2h5 = total5
        else:
            share = (total5 + world - 1) // world
            h5_part = torch.empty(share + 64, dtype=torch.uint8).pin_memory()

            def c5_host():
                d5.copy_(h5, non_blocking=True)
                fin, tot = sharding.compress_simple_sharded(ctx, d5, lo5, hi5, n5, mode=1, container=0, out=full5)
                a = min(tot, rank * share)
                b = min(tot, a + share)
                h5_part[: b - a].copy_(fin[a:b], non_blocking=True)
                torch.cuda.synchronize()

            hms5 = timed_host(c5_host, 2, warmup=1)
            d2h5 = share
        c5 = {"metric": "huffman-only MB/s in", "value": round(n5 / 1e6 / (ms5 / 1e3), 1), "unit": "MB/s", "ms_per_step": round(ms5, 3),
              "scaling": "strong", "higher_is_better": True,
              "workload": "huffman-only, ONE %d MiB random+zeros stream (runs of 4..256 KiB) over %d GPU(s) (BASELINE configs[4])%s"
                          % (n5 // MIB, world, "; sharded by 65535-byte block ranges, every shard packed at its bit offset into the rank's copy of "
                                               "the stream and broadcast from there (NCCL), so that every rank ends with the whole stream" if world > 1 else ""),
              "ratio": round(n5 / total5, 3), "compressed_bytes": total5,
              "e2e": {"value": round(n5 / 1e6 / (hms5 / 1e3), 1), "unit": "MB/s", "ms_per_step": round(hms5, 3),
                      "h2d_bytes_per_step": sh_bytes, "d2h_bytes_per_step": d2h5},
              "roofline": roofline_of(ph5, (n5 + total5) / world, "huffman:%d" % (sh_bytes // MIB))}
        if ms5_in_place is not None:
            c5["shards_in_place"] = {"value": round(n5 / 1e6 / (ms5_in_place / 1e3), 1), "unit": "MB/s", "ms_per_step": round(ms5_in_place, 3),
                                     "workload": "the same stream, left distributed: every rank keeps its own byte range of the stream (and the "
                                                 "bytes it shares with its neighbours, exchanged as values); no bulk collective"}
        # parity: zlib inflates the stream back to the generated bytes (rank 0 regenerates the other ranks' ranges), and at
        # N = 1 the first 256 MiB of blocks equal the oracle's bit for bit (huffman-only blocks do not depend on each other)
        if rank == 0 and not args.skip_cpu:
            import zlib
            comp = out5[0][:total5].cpu().numpy()
            dobj = zlib.decompressobj(-15)
            okz, p = True, 0
            t0 = time.perf_counter()
            for a in range(0, total5, 64 * MIB):
                got = dobj.decompress(comp[a:a + 64 * MIB].tobytes())
                q = p + len(got)
                want_piece = mine[p:q] if q <= sh_bytes else gen_range(p, q)
                okz = okz and got == want_piece.tobytes()
                p = q
            okz = okz and p == n5 and dobj.eof
            assert okz, "huffman-only stream does not inflate back to its input"
            o, olib = oracle_lib()
            samp = min(sh_bytes, 256 * MIB) // 65535 * 65535
            t0 = time.perf_counter()
            wanth = o.compress(mine[:samp], o.RAW, 1, _lib_override=olib)
            dt = time.perf_counter() - t0
            # the oracle's stream of the sample ends with an empty final stored block (3 header bits, padding, LEN, NLEN)
            # that the big stream does not have at that place: compare up to the byte before it can start
            same = comp[: len(wanth) - 7].tobytes() == wanth[:-7] if len(wanth) > 7 else True
            assert same, "huffman-only: GPU stream differs from the CPU oracle on the sample"
            c5["cpu_baseline"] = {"value": round(samp / 1e6 / dt, 1), "unit": "MB/s", "cores": 1, "kind": "port",
                                  "sample": "first %d MiB of the stream, one pass; those blocks are bit-identical in the GPU stream; "
                                            "the whole stream inflates (zlib) back to the input" % (samp // MIB)}
        del d5, h5

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {
        "metric": "deflate L6 MB/s in", "value": round(value, 1), "unit": "MB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "raw deflate level %d, %d MiB enwik-like synthetic bytes per GPU (BASELINE configs[1])"
                               % (level, n // MIB),
                   "ratio": round(n / out_len, 3), "compressed_bytes": out_len,
                   "l2": "inputs (256 MiB and more per leg) exceed the 126 MB L2; no explicit flush",
                   "multi_gpu": "independent chunk per rank + NCCL all-gather of per-shard outputs" if world > 1 else "n/a"},
        "e2e": {"value": round(e2e_value, 1), "unit": "MB/s", "ms_per_step": round(e2e_ms, 3),
                "h2d_bytes_per_step": n, "d2h_bytes_per_step": out_len},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "inflate": inflate,
        "c4_level9": c4,
        "c5_huffman": c5,
        "mixed": mixed,
        "streaming": streaming,
        "single_stream": single,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
