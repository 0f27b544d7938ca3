"""Batch sharding across ranks (one process per GPU) and the single exchange step of the eval path.

Every image is an independent unit: rank r evaluates the contiguous shard
[r*B/R, (r+1)*B/R) with no data-path collective; the only exchange is one all_gather of the per-image
integer IoU counts and fp64 HD / MSD values before rank 0 forms the means exactly as the reference does
(model/engine/inference.py:171-173).  Works with NCCL (CUDA tensors) and gloo (CPU tensors, tests)."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank=None, world_size=None):
    """Contiguous shard of `n` units for `rank`; the first n % world ranks get one extra unit."""
    if rank is None:
        rank, world_size = world()
    base, extra = divmod(n, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_rows(t):
    """all_gather of a [rows, cols] tensor whose row count may differ per rank (zero-padded to the max, then trimmed).
    Returns the concatenation in rank order on every rank."""
    rank, ws = world()
    if ws == 1:
        return t
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    m = max(counts)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad)
    return torch.cat([o[:c] for o, c in zip(out, counts)], 0)


def gather_names(names):
    """File names of every rank's shard, concatenated in rank order (same order as gather_rows)."""
    rank, ws = world()
    if ws == 1:
        return list(names)
    out = [None] * ws
    dist.all_gather_object(out, list(names))
    return [n for part in out for n in part]


def pack_metrics(inter, union, hd, msd):
    """[B,99] x4 -> one fp64 [B, 396] tensor (integer counts are exact in fp64)."""
    return torch.cat([inter.to(torch.float64), union.to(torch.float64), hd, msd], dim=1)


def unpack_metrics(packed):
    n = packed.shape[1] // 4
    p = packed.cpu().numpy()
    return (p[:, :n].astype("int64"), p[:, n:2 * n].astype("int64"), p[:, 2 * n:3 * n], p[:, 3 * n:])


def allreduce_flat(flat, bucket_elems=1 << 25, group=None):
    """Training exchange step: SUM all-reduce of the flat gradient buffer in large buckets (replaces nn.DataParallel's
    gradient reduction, reference train.py:117-121).  Returns the world size; the optimizer scales by 1 / world_size."""
    rank, ws = world()
    if ws == 1:
        return 1
    for s in range(0, flat.numel(), bucket_elems):
        dist.all_reduce(flat[s:s + bucket_elems], op=dist.ReduceOp.SUM, group=group)
    return ws


class OverlappedGradAllReduce:
    """SUM all-reduce of the flat gradient buffer in two parts: the segmentation net's slice [lo, hi) is launched (async, on
    NCCL's own stream) from an autograd hook the moment its backward is complete and runs under the backward of the SR net;
    the rest follows when backward returns.  Inside a CUDA-graph capture both become nodes of the graph (fork / join through
    the events ProcessGroupNCCL records), so a replay contains the exchange.  Falls back to one all-reduce of the whole
    buffer when the hook did not fire (segmentation net frozen / evaluated without a tape) or `span` is None."""

    def __init__(self, flat, span, bucket_elems=1 << 25, group=None):
        self.flat, self.span, self.bucket, self.group = flat, span, bucket_elems, group
        self.works = []

    @staticmethod
    def prefix_span(named_params, opt_params, slots, prefix="segmentation_model."):
        """(first element, one past the last element) of the `prefix` parameters in the flat buffer if their slots are
        contiguous and something else is left, else None."""
        index = {id(p): i for i, p in enumerate(opt_params)}
        idx = sorted(index[id(p)] for n, p in named_params if n.startswith(prefix) and id(p) in index)
        if not idx or idx != list(range(idx[0], idx[-1] + 1)) or len(idx) == len(opt_params):
            return None
        return slots[idx[0]][0], slots[idx[-1]][0] + slots[idx[-1]][1]

    def _active(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _reduce(self, lo, hi, async_op):
        out = []
        for s in range(lo, hi, self.bucket):
            w = dist.all_reduce(self.flat[s:min(hi, s + self.bucket)], op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
            if async_op:
                out.append(w)
        return out

    def seg_done(self):
        if self.span is None or self.works or not self._active():
            return
        self.works = self._reduce(self.span[0], self.span[1], True)

    def finish(self):
        """All-reduce what is still local, then join the async part.  Returns the world size."""
        if not self._active():
            self.works = []
            return 1
        n = self.flat.numel()
        if self.works:
            self._reduce(0, self.span[0], False)
            self._reduce(self.span[1], n, False)
        else:
            self._reduce(0, n, False)
        for w in self.works:
            w.wait()
        self.works = []
        return dist.get_world_size(self.group)
