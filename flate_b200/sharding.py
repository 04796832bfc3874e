"""Host-side sharding of independent units (chunks, members, 65535-byte blocks) over ranks, and the
all-gather of ragged per-shard outputs.  One process per GPU; torch.distributed is plumbing only.
The deflate/inflate path has no data-path collective: units are independent."""
import torch
import torch.distributed as dist


def shard_range(n_units, rank, world):
    """Contiguous, balanced [lo, hi) of n_units for `rank` (first n_units % world ranks get one more)."""
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_sizes(size, device):
    """Every rank's byte count, as a python list."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [int(size)]
    t = torch.tensor([int(size)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(x.item()) for x in out]


def all_gather_ragged(shard, size, pad_to=None, out=None):
    """All-gather of per-shard outputs of different lengths.  `shard` is a uint8 tensor holding at
    least `pad_to` bytes of which the first `size` are payload.  Returns (buffer, sizes, pad) where
    rank r's payload is buffer[r*pad : r*pad + sizes[r]]."""
    sizes = all_gather_sizes(size, shard.device)
    world = len(sizes)
    pad = pad_to if pad_to is not None else (max(sizes) + 255) // 256 * 256
    if world == 1:
        return shard[:pad], sizes, pad
    if out is None:
        out = torch.empty(world * pad, dtype=torch.uint8, device=shard.device)
    dist.all_gather_into_tensor(out, shard[:pad].contiguous())
    return out, sizes, pad


def concat_ragged(buffer, sizes, pad):
    """Byte-exact concatenation of the gathered payloads (host bytes)."""
    return b"".join(buffer[r * pad: r * pad + sizes[r]].cpu().numpy().tobytes() for r in range(len(sizes)))


class _NoStream:  # CPU tensors (host-logic tests with a stub context)
    cuda_stream = 0

    def synchronize(self):
        pass


def shard_positions(n, world, align):
    """Position ranges [lo, hi) of a stream of n bytes for every rank: equal parts, `align`-aligned."""
    per = ((n + world - 1) // world + align - 1) // align * align if n else align
    return per, [(min(n, r * per), min(n, (r + 1) * per)) for r in range(world)]


def merge_overlap(own, tail):
    """Entries of a position range evaluated by its own rank (`own`) and, for its first positions, by the
    rank before it (`tail`, the overlap of that rank's last chunk): any evaluated entry (!= -1) is right."""
    k = min(own.numel(), tail.numel())
    if k:
        own[:k] = torch.where(own[:k] == -1, tail[:k], own[:k])
    return own


def compress_stream_sharded(ctx, d_in, n, d_out, level=6, container=0, root=0, keep=None):
    """One deflate stream of n bytes (resident at torch uint8 tensor d_in on every rank) compressed by all
    ranks together: every rank evaluates the lazy-parse steps of its own position range (sparse parse), the
    tables are all-gathered over NCCL, rank `root` runs the (cheap, sequential-in-nature) parse + block
    writer.  Returns the compressed size on `root` (0 elsewhere).  Output is byte-identical to
    Context.compress_device."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    ov = ctx.shard_overlap
    per, ranges = shard_positions(n, world, ctx.shard_align)
    lo, hi = ranges[rank]
    nx = torch.empty(world * per + ov, dtype=torch.int32, device=d_in.device)
    cur = torch.cuda.current_stream() if d_in.is_cuda else _NoStream()
    sp = cur.cuda_stream
    # the legacy default stream has handle 0, which the C ABI reads as "the context's own stream": make sure
    # torch's work on the inputs is complete before the library touches them (and again before stage 2)
    cur.synchronize()
    ok = ctx.shard_search(d_in.data_ptr(), n, lo, hi, nx.data_ptr(), level=level, stream=sp)
    if world > 1:
        flag = torch.tensor([0 if ok else 1], dtype=torch.int32, device=d_in.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        ok = int(flag.item()) == 0
    if not ok:  # periodic data somewhere: every rank falls back to the dense tables for its range
        ctx.set_parse_mode(1)
        try:
            cur.synchronize()
            ctx.shard_search(d_in.data_ptr(), n, lo, hi, nx.data_ptr(), level=level, stream=sp)
        finally:
            ctx.set_parse_mode(0)
    if world > 1:
        tail = nx[rank * per + per: rank * per + per + ov].clone()   # this rank's entries past its range
        tails = torch.empty(world * ov, dtype=torch.int32, device=d_in.device)
        dist.all_gather_into_tensor(tails, tail)
        dist.all_gather_into_tensor(nx[:world * per], nx[rank * per:(rank + 1) * per].clone())
        if rank == root and ok:
            for r in range(world - 1):
                merge_overlap(nx[(r + 1) * per:(r + 1) * per + ov], tails[r * ov:(r + 1) * ov])
    if keep is not None:
        keep["nx"], keep["per"] = nx, per   # development aid
    if rank != root:
        return 0
    cur.synchronize()
    cap = d_out.numel()
    return ctx.shard_finish(d_in.data_ptr(), n, nx.data_ptr(), d_out.data_ptr(), cap, level=level, container=container,
                            stream=sp)
