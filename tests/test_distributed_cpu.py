"""world_size-2 gloo tests (CPU) of the N>1 host logic: contiguous batch shards + the one all_gather of metrics."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from csbsr_b200.engine import distributed as D
    from oracle import metrics_ref as M
    rng = np.random.default_rng(0)
    prob = rng.random((B, 1, 12, 16)).astype(np.float32)
    mask = (rng.random((B, 1, 12, 16)) < 0.4).astype(np.float32)
    lo, hi = D.shard_range(B)
    inter, union = M.iou_counts(prob[lo:hi], mask[lo:hi])          # the oracle stands in for the GPU kernel here
    hd, msd = M.distance_metrics(prob[lo:hi], mask[lo:hi], 50)
    packed = D.pack_metrics(torch.from_numpy(inter), torch.from_numpy(union), torch.from_numpy(hd), torch.from_numpy(msd))
    allm = D.gather_rows(packed)
    if rank == 0:
        q.put(allm.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [5, 8])
def test_sharded_metrics_equal_single_process(B):
    from csbsr_b200.engine import distributed as D
    from oracle import metrics_ref as M
    assert [D.shard_range(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    assert [D.shard_range(8, r, 4) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 8)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + B
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    prob = rng.random((B, 1, 12, 16)).astype(np.float32)
    mask = (rng.random((B, 1, 12, 16)) < 0.4).astype(np.float32)
    inter, union = M.iou_counts(prob, mask)
    hd, msd = M.distance_metrics(prob, mask, 50)
    i2, u2, h2, m2 = D.unpack_metrics(torch.from_numpy(got))
    assert np.array_equal(i2, inter) and np.array_equal(u2, union)
    assert np.array_equal(h2, hd) and np.array_equal(m2, msd)
    # AIU / AHD exactly as inference.py:171-173
    assert np.mean((i2 + 1e-5) / (u2 + 1e-5)) == np.mean((inter + 1e-5) / (union + 1e-5))


def _grad_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from csbsr_b200.engine import distributed as D
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(1000, generator=g)
    ws = D.allreduce_flat(flat, bucket_elems=256)                   # 4 buckets, the last one ragged
    if rank == 0:
        q.put((ws, (flat / ws).numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_is_mean_over_ranks():
    """The training exchange step: bucketed SUM all-reduce of the flat gradient, scaled by 1 / world in the optimizer."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + (os.getpid() % 500)
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ws, got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = sum(torch.randn(1000, generator=torch.Generator().manual_seed(100 + r)) for r in range(2)) / 2
    assert ws == 2 and np.allclose(got, ref.numpy(), rtol=0, atol=1e-7)


def _overlap_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from csbsr_b200.engine import distributed as D
    out = []
    for fire in (True, False):                                      # hook fired / segmentation net without a tape
        g = torch.Generator().manual_seed(200 + rank)
        flat = torch.randn(1000, generator=g)
        ar = D.OverlappedGradAllReduce(flat, (300, 700), bucket_elems=256)
        if fire:
            ar.seg_done()                                           # the slice [300, 700) leaves first (async) ...
            ar.seg_done()                                           # ... and only once
            assert len(ar.works) == 2
        ws = ar.finish()                                            # ... then [0, 300) and [700, 1000); join
        assert ar.works == []
        out.append((ws, flat.numpy().copy()))
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_gradient_allreduce_matches_plain_sum():
    """OverlappedGradAllReduce (segmentation slice from the autograd hook, SR slice after backward) == SUM over ranks of the
    whole flat gradient, with and without the hook firing; prefix_span finds the segmentation slice of the flat buffer."""
    from csbsr_b200.engine import distributed as D
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29300 + (os.getpid() % 500)
    procs = [ctx.Process(target=_overlap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = sum(torch.randn(1000, generator=torch.Generator().manual_seed(200 + r)) for r in range(2)).numpy()
    for ws, got in out:
        assert ws == 2 and np.allclose(got, ref, rtol=0, atol=1e-6)
    ps = [torch.nn.Parameter(torch.zeros(3)) for _ in range(5)]
    named = [("sr_model.a", ps[0]), ("sr_model.b", ps[1]), ("segmentation_model.c", ps[2]), ("segmentation_model.d", ps[3]),
             ("segmentation_model.e", ps[4])]
    slots = [(0, 4), (4, 4), (8, 4), (12, 4), (16, 4)]
    span = D.OverlappedGradAllReduce.prefix_span
    assert span(named, ps, slots) == (8, 20)
    assert span([("segmentation_model.x", ps[0]), ("segmentation_model.y", ps[1]), ("sr_model.z", ps[2])], ps[:3], slots[:3]) == (0, 8)
    assert span(named, [ps[2], ps[0], ps[3], ps[1], ps[4]], slots) is None      # not contiguous
    assert span(named[2:], ps[2:], slots[:3]) is None                          # nothing else in the buffer
