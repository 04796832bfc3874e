#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 DEFLATE engine (contract in the task statement).

Default workload = BASELINE.json configs[1]: raw deflate level 6 over 256 MiB of enwik-like synthetic
bytes on one B200.  A "step" is one pass of the hot path over that batch.  With --gpus N (launched
under torchrun) every rank compresses its own independent 256 MiB chunk (the path shards by chunk;
weak scaling) and the per-shard outputs are all-gathered over NCCL; for N > 1 it also compresses ONE stream
sharded by position over all ranks (single_stream).  The same run also measures the inflate side (config C3
shape: 1 MiB gzip members, 1 GiB of plain output split over the ranks, plus 1 GiB per rank as inflate.weak).

  value     : whole-job deflate L6 throughput, MB/s of INPUT, inputs resident in HBM (CUDA events)
  e2e       : same metric through the public host-buffer call (pinned host -> H2D -> kernels -> D2H)
  roofline  : dominant kernel (the phase with the largest live CUDA-event time: sparse_parse) algorithmic bytes /
              that time vs the measured HBM peak; traffic from profiles/traffic.json (ncu --set full)
  cpu_baseline : the CPU oracle (a port of the reference's algorithm; the reference is Zig and there
                 is no zig toolchain) timed on this box's host cores, single thread like the reference

--impl reference times the reference's CPU implementation of the path (the oracle port) instead.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.pop("NCCL_DEBUG", None)  # keep stdout to the one JSON line (NCCL prints its version banner there otherwise)

MIB = 1 << 20
WORKLOAD_BYTES = 256 * MIB
MEMBER_BYTES = 1 * MIB
INFLATE_TOTAL = 1024 * MIB
LEVEL = 6


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port) on the box's host cores.  One stream is one thread
    (the reference has no threads); the N-GPU workload is N independent chunks, which a CPU box can
    run side by side, so rank 0 compresses a bounded sample of every rank's chunk on min(N, host cores) threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as o
    try:
        lib = o.lib(o.build(native=True))
    except Exception:
        lib = o.lib()
    world = max(1, args.gpus)
    threads = max(1, min(world, host_threads()))
    sample = 32 * MIB
    chunks = [make_text(sample, r) for r in range(world)]
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
    value = world * sample / 1e6 / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "deflate L6 MB/s in", "value": round(value, 2), "unit": "MB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "raw deflate level %d, %d MiB enwik-like synthetic bytes per GPU (BASELINE configs[1])"
                               % (LEVEL, WORKLOAD_BYTES // MIB),
                   "sample": "first 32 MiB of every rank's chunk per step, %d chunk(s) on %d host thread(s)" % (world, threads),
                   "ratio": round(sample / out_len, 3)},
        "cpu_baseline": {"value": round(value, 2), "unit": "MB/s", "cores": threads, "kind": "port",
                         "sample": "32 MiB of each rank's synthetic text per step; oracle/flate_oracle.c -O3 -march=native"},
        "e2e": {"value": round(value, 2), "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bytes", type=int, default=WORKLOAD_BYTES, help="per-GPU chunk size (default 256 MiB)")
    ap.add_argument("--level", type=int, default=LEVEL)
    ap.add_argument("--skip-inflate", action="store_true")
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
    ctx = flate_b200.Context(local_rank)
    n = args.bytes
    level = args.level
    text = make_text(n, rank)                       # this rank's independent chunk
    h_in = torch.from_numpy(text).pin_memory()
    d_in = h_in.cuda(non_blocking=True)
    cap = ctx.lib.fb200_compress_bound(n, level) + 64
    d_out = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
    h_out = torch.empty(cap + 64, dtype=torch.uint8).pin_memory()
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        return ctx.compress_device(d_in.data_ptr(), n, d_out.data_ptr(), cap, mode=level, stream=sp)

    from flate_b200 import sharding
    d_outs = [d_out, torch.empty_like(d_out)] if world > 1 else [d_out]
    gather_bufs, pending = [], [None, None]
    out_len = device_step()
    if world > 1:
        # per-shard outputs are all-gathered (north star); pad to a common size agreed on once
        allsz = sharding.all_gather_sizes(out_len, d_out.device)
        pad = (int(max(allsz) * 1.02) + 4096) // 256 * 256
        gather_bufs = [torch.empty(world * pad, dtype=torch.uint8, device="cuda") for _ in range(2)]
    step_no = [0]

    def full_step():
        # double-buffered: the all-gather of step k runs on NCCL's stream while step k+1 compresses
        i = (step_no[0] & 1) if world > 1 else 0
        step_no[0] += 1
        if pending[i] is not None:
            pending[i].wait()
            pending[i] = None
        buf = d_outs[i]
        ln = ctx.compress_device(d_in.data_ptr(), n, buf.data_ptr(), cap, mode=level, stream=sp)
        if world > 1:
            pending[i] = dist.all_gather_into_tensor(gather_bufs[i], buf[:pad], async_op=True)
        return ln

    def drain_gathers():
        for i in range(2):
            if pending[i] is not None:
                pending[i].wait()
                pending[i] = None

    # ---- device-resident timing (value) ----
    for _ in range(args.warmup):
        full_step()
    drain_gathers()
    ctx.profile(True)
    launches0 = ctx.kernel_launches
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        out_len = full_step()
    drain_gathers()  # the stream now waits for the last all-gathers: they are inside the timed region
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    launches = ctx.kernel_launches - launches0
    phases = ctx.profile_read()
    ctx.profile(False)
    t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * n / 1e6 / (ms_step / 1e3)

    # ---- end-to-end through the public host-buffer call ----
    def e2e_step():
        ln = C.c_size_t(0)
        rc = ctx.lib.fb200_compress(ctx.h, flate_b200.RAW, level, h_in.data_ptr(), n, h_out.data_ptr(), cap, C.byref(ln))
        if rc:
            raise RuntimeError("fb200_compress failed: %d" % rc)
        return ln.value

    import ctypes as C
    for _ in range(2):
        e2e_len = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_len = e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = world * n / 1e6 / (e2e_ms / 1e3)
    assert e2e_len == out_len
    compressed = h_out[:out_len].numpy().tobytes()

    # ---- one stream over all GPUs: position-sharded search + all-gather of the lazy-step tables ----
    single = None
    if world > 1:
        d_stream = d_in.clone()
        dist.broadcast(d_stream, src=0)                      # every rank works on rank 0's stream
        for _ in range(2):
            m = sharding.compress_stream_sharded(ctx, d_stream, n, d_out, level=level)
        barrier()
        e0s, e1s = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0s.record(stream)
        for _ in range(args.steps):
            m = sharding.compress_stream_sharded(ctx, d_stream, n, d_out, level=level)
        e1s.record(stream)
        barrier()
        t = torch.tensor([e0s.elapsed_time(e1s) / args.steps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sms = float(t.item())
        if rank == 0:
            same = m == len(compressed) and d_out[:m].cpu().numpy().tobytes() == compressed
            single = {"metric": "deflate L6 MB/s in, ONE %d MiB stream sharded by position over %d GPUs" % (n // MIB, world),
                      "value": round(n / 1e6 / (sms / 1e3), 1), "unit": "MB/s", "ms_per_step": round(sms, 3),
                      "scaling": "strong", "identical_to_one_gpu_stream": bool(same),
                      "collective": "NCCL all-gather of the 4 B/position lazy-step tables"}
        del d_stream

    # ---- inflate side: 1 MiB gzip members (config C3 shape) ----
    inflate = None
    if not args.skip_inflate:
        total = INFLATE_TOTAL // world                 # strong scaling: 1 GiB of plain output over all ranks
        nmem = max(1, total // MEMBER_BYTES)
        uniq = min(nmem, n // MEMBER_BYTES)
        # Members are level-6 gzip streams of 1 MiB slices of the text.  The reference's inflate rejects
        # code-length runs that cross the literal/distance boundary (inflate.zig:161-170) although its own
        # block writer emits them (block_writer.zig:78-171); we reproduce that, so such members (about 1 in
        # 40) are re-cut from a shifted slice until the stream is one the reference itself accepts.
        members, plains = [], []
        for i in range(uniq):
            for shift in range(0, 64):
                lo = (i * MEMBER_BYTES + shift * 4099) % (n - MEMBER_BYTES + 1)
                sl = text[lo:lo + MEMBER_BYTES]
                m = ctx.compress(sl, flate_b200.GZIP, LEVEL)
                try:
                    ctx.decompress(m, flate_b200.GZIP, cap=MEMBER_BYTES + 64)
                    break
                except flate_b200.FlateError:
                    continue
            members.append(m)
            plains.append(sl)
        def time_inflate(nmem):
            blob = b"".join(members[i % uniq] for i in range(nmem))
            lens = np.array([len(members[i % uniq]) for i in range(nmem)], dtype=np.uint64)
            offs = np.zeros(nmem, dtype=np.uint64)
            offs[1:] = np.cumsum(lens)[:-1]
            d_blob = torch.from_numpy(np.frombuffer(blob, dtype=np.uint8).copy()).cuda()
            d_plain = torch.empty(nmem * MEMBER_BYTES + 64, dtype=torch.uint8, device="cuda")
            ooff = np.arange(nmem, dtype=np.uint64) * np.uint64(MEMBER_BYTES)
            ocap = np.full(nmem, MEMBER_BYTES, dtype=np.uint64)

            def inflate_step():
                rc, ol, used, st = ctx.decompress_members_device(d_blob.data_ptr(), offs, lens, d_plain.data_ptr(), ooff,
                                                                 ocap, flate_b200.GZIP, stream=sp)
                if rc:
                    raise RuntimeError("inflate failed: %d" % rc)
                return int(ol.sum())

            for _ in range(2):
                plain_bytes = inflate_step()
            ctx.profile(True)
            barrier()
            isteps = max(2, args.steps // 2)
            e0.record(stream)
            for _ in range(isteps):
                plain_bytes = inflate_step()
            e1.record(stream)
            barrier()
            ims = e0.elapsed_time(e1) / isteps
            iph = ctx.profile_read()
            ctx.profile(False)
            t = torch.tensor([ims], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            chk = d_plain[:MEMBER_BYTES].cpu().numpy()
            kms = iph["inflate_members"][0] / max(1, iph["inflate_members"][1])
            return float(t.item()), plain_bytes, len(blob), kms, chk

        ims, plain_bytes, blob_len, kms, chk = time_inflate(nmem)
        # parity: a member's plain bytes equal the text it was made from
        assert (chk == plains[0]).all(), "inflate output differs from the original text"
        peak, src = peaks()
        algo = float(blob_len + plain_bytes)
        inflate = {"metric": "inflate MB/s out", "value": round(world * plain_bytes / 1e6 / (ims / 1e3), 1), "unit": "MB/s",
                   "ms_per_step": round(ims, 3), "scaling": "strong",
                   "workload": "%d gzip members x 1 MiB plain (level 6 text) per GPU, %d MiB plain over %d GPU(s)"
                               % (nmem, world * plain_bytes // MIB, world),
                   "roofline": {"bound": "hbm", "achieved": round(algo / 1e9 / (kms / 1e3), 2), "peak": peak, "unit": "GB/s",
                                "frac": round(algo / 1e9 / (kms / 1e3) / peak, 5), "traffic": None,
                                "kernel": "inflate_members_kernel", "kernel_ms": round(kms, 3)}}
        # CPU baseline of the inflate side: the oracle (port of inflate.zig) on one host core, 64 members
        if rank == 0 and not args.skip_cpu:
            from oracle import oracle as o
            ksample = min(64, uniq)
            t0 = time.perf_counter()
            for i in range(ksample):
                pl, _ = o.decompress(members[i], o.GZIP, cap=MEMBER_BYTES + 64)
                assert len(pl) == MEMBER_BYTES
            dt = time.perf_counter() - t0
            inflate["cpu_baseline"] = {"value": round(ksample * MEMBER_BYTES / 1e6 / dt, 1), "unit": "MB/s", "cores": 1,
                                       "kind": "port", "sample": "%d of the same members, oracle inflate, 1 thread" % ksample}
        if world > 1:
            # the same members, 1 GiB of plain output PER GPU: one warp decodes one member, so a GPU needs
            # about a thousand members in flight; the strong-scaling leg above leaves 1024 / N per GPU
            wn = max(1, INFLATE_TOTAL // MEMBER_BYTES)
            wms, wplain, _, wk, wchk = time_inflate(wn)
            assert (wchk == plains[0]).all()
            inflate["weak"] = {"value": round(world * wplain / 1e6 / (wms / 1e3), 1), "unit": "MB/s",
                               "ms_per_step": round(wms, 3), "scaling": "weak", "kernel_ms": round(wk, 3),
                               "workload": "%d gzip members x 1 MiB plain per GPU" % wn}

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peak, peak_src = peaks()
    ratio = n / out_len
    algo_bytes = n + out_len                                  # N_in + N_out per launch
    dom = max((k for k in phases if phases[k][1]), key=lambda k: phases[k][0])
    dom_ms = phases[dom][0] / phases[dom][1]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(dom)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": round(algo_bytes / 1e9 / (dom_ms / 1e3), 2), "peak": peak, "unit": "GB/s",
                "frac": round(algo_bytes / 1e9 / (dom_ms / 1e3) / peak, 5), "traffic": traffic, "kernel": dom,
                "kernel_ms": round(dom_ms, 3), "peak_source": peak_src,
                "phases_ms": {k: round(v[0] / max(1, v[1]), 3) for k, v in phases.items() if v[1]}}

    # ---- CPU baseline: the oracle on this box's host cores, same bytes, and the parity check ----
    cpu = None
    if not args.skip_cpu:
        from oracle import oracle as o
        try:
            lib = o.lib(o.build(native=True))
        except Exception:
            lib = o.lib()
        sample = min(n, 256 * MIB)
        t0 = time.perf_counter()
        want = o.compress(text[:sample], o.RAW, level, _lib_override=lib)
        dt = time.perf_counter() - t0
        if sample == n:
            assert want == compressed, "GPU output differs from the CPU oracle"
        cpu = {"value": round(sample / 1e6 / dt, 2), "unit": "MB/s", "cores": 1, "kind": "port",
               "sample": "the whole %d MiB rank-0 chunk, one pass; oracle/flate_oracle.c (-O3 -march=native); "
                         "output compared byte for byte with the GPU's: %s; box has %d host threads, the reference is "
                         "single-threaded" % (sample // MIB, "identical" if sample == n else "n/a", host_threads())}

    line = {
        "metric": "deflate L6 MB/s in", "value": round(value, 1), "unit": "MB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "raw deflate level %d, %d MiB enwik-like synthetic bytes per GPU (BASELINE configs[1])"
                               % (level, n // MIB),
                   "ratio": round(ratio, 3), "compressed_bytes": out_len,
                   "l2": "inputs (256 MiB) exceed the 126 MB L2; no explicit flush",
                   "multi_gpu": "independent chunk per rank + NCCL all-gather of per-shard outputs" if world > 1 else "n/a"},
        "e2e": {"value": round(e2e_value, 1), "unit": "MB/s", "ms_per_step": round(e2e_ms, 3),
                "h2d_bytes_per_step": n, "d2h_bytes_per_step": out_len},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "inflate": inflate,
        "single_stream": single,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
