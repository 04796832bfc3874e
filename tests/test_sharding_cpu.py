"""World-size-2 gloo tests (CPU) of the multi-rank host logic: unit sharding and the ragged
all-gather of per-shard outputs.  The per-shard payloads here come from the CPU oracle (this is a
test of the plumbing, not of the kernels)."""
import os
import socket
import zlib

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flate_b200 import sharding


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 9, 1024, 65537):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flate_b200 import synth
        from oracle import oracle as o
        nmembers = 5
        plains = [synth.enwik_like(20000 + 3000 * i, seed=50 + i).tobytes() for i in range(nmembers)]
        lo, hi = sharding.shard_range(nmembers, rank, world)
        mine = b"".join(o.compress(p, o.GZIP, 6) for p in plains[lo:hi])
        shard = torch.zeros(1 << 17, dtype=torch.uint8)
        shard[: len(mine)] = torch.from_numpy(np.frombuffer(mine, dtype=np.uint8).copy())
        buf, sizes, pad = sharding.all_gather_ragged(shard, len(mine))
        whole = sharding.concat_ragged(buf, sizes, pad)
        # every rank ends up with the same multi-member gzip stream, in member order
        d = zlib.decompressobj(31)
        out, rest = b"", whole
        while rest:
            d = zlib.decompressobj(31)
            out += d.decompress(rest)
            rest = d.unused_data
        q.put((rank, out == b"".join(plains), sizes))
    finally:
        dist.destroy_process_group()


def test_ragged_all_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == res[1][2] and len(res[0][2]) == 2
