"""On-hardware N-rank parity ("1/2/4/8 ranks ... compare against the single-GPU result: counts
bit-exact, losses / gradients to tolerance").  Needs >= 2 visible GPUs (skipped otherwise): `gpurun --gpus 2 -- python -m
pytest tests/test_multigpu_gpu.py -m gpu`.

Eval: the same 8-image batch is evaluated (degrade -> KBPN -> PSPNet -> AIU / HD / MSD sweep) once by a single process and
once by two ranks over NCCL, each on its contiguous shard (engine/distributed.py::shard_range) with the one all_gather of
the packed [B, 4*99] rows; I, U, HD and MSD must be bit-equal (every image is an independent unit and every kernel on the
eval path is deterministic).

Training: each rank runs forward + loss + backward on its own 2-image shard (per-replica BatchNorm statistics and w^F
broadcast, exactly the reference's DataParallel semantics) and the flat gradients are SUM-all-reduced; rank 0's result / 2
must equal the mean of the two shard gradients computed one after the other by a single process, to 1e-6 of the gradient's
max-abs (the backward kernels accumulate with fp32 atomics, so bit-equality is not expected)."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B_EVAL, SIZE_EVAL, B_TRAIN, SIZE_TRAIN = 8, 192, 2, 96


def _eval_rows(lo, hi, dev):
    from csbsr_b200.config import cfg
    from csbsr_b200.data import degrade as G
    from csbsr_b200.engine import distributed as D, inference as E
    from csbsr_b200.modeling.build_model import JointModel
    from csbsr_b200.utils import synth
    c = cfg.clone()
    c.merge_from_file(os.path.join(ROOT, "config", "config_csbsr_pspnet.yaml"))
    m = JointModel(c)
    m.load_state_dict(synth.model_state_dict(), strict=True)
    m.chunk = 2
    hr, mask = synth.batch(0, B_EVAL, SIZE_EVAL)
    params = torch.as_tensor(synth.degradation_params(B_EVAL))
    lr, _ = G.degrade(hr[lo:hi].to(dev), params[lo:hi].to(dev))
    sr, seg, kp = m(lr, None)
    r = E.seg_metrics(seg, mask[lo:hi].to(dev), with_hd=True, to_host=False)
    return D.pack_metrics(r["inter"], r["union"], r["hd"], r["msd"])


def _train_grad(shard, dev):
    """flat gradient (on `dev`) of one forward + loss + backward on shard `shard` of the training batch; deterministic inputs."""
    from csbsr_b200.config import cfg
    from csbsr_b200.engine.losses import calc_loss
    from csbsr_b200.engine.optim import FusedAdam
    from csbsr_b200.modeling.build_model import JointModelWithLoss
    from csbsr_b200.data import degrade as G
    from csbsr_b200.utils import synth
    c = cfg.clone()
    c.merge_from_file(os.path.join(ROOT, "config", "config_csbsr_pspnet.yaml"))
    c.SOLVER.SEG_FAIL_ORIENTED_WEIGHT4SS_AMP = 1.0
    m = JointModelWithLoss(c, num_train_ds=100, resume_iter=40000)
    m.load_state_dict(synth.model_state_dict(), strict=True)
    m.to(dev).train()
    m.dropout = False                                  # Dropout2d masks are random per process; everything else is deterministic
    opt = FusedAdam(m.parameters(), lr=c.SOLVER.LR)
    hr, mask = synth.batch(500 + shard * B_TRAIN, B_TRAIN, SIZE_TRAIN)
    hr, mask = hr.to(dev), mask.to(dev)
    params = torch.as_tensor(synth.degradation_params(B_TRAIN, seed=60 + shard)).to(dev)
    lr_img, kernels = G.degrade(hr, params)
    seg_loss, sr_loss, seg, sr, kp = m(40001, lr_img, sr_targets=hr, segment_targets=mask, kernel_targets=kernels.unsqueeze(1))
    loss = calc_loss(sr_loss, seg_loss.mean(), c.SOLVER.TASK_LOSS_WEIGHT, 40001, c)
    loss.backward()
    return opt, loss.detach()


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RANK"], os.environ["WORLD_SIZE"], os.environ["LOCAL_RANK"] = str(rank), str(world), str(rank)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from csbsr_b200.engine import distributed as D
    lo, hi = D.shard_range(B_EVAL)
    rows = D.gather_rows(_eval_rows(lo, hi, dev))
    opt, loss = _train_grad(rank, dev)
    ws = opt.all_reduce_grads()
    torch.cuda.synchronize()
    if rank == 0:
        np.savez(out_path, rows=rows.cpu().numpy(), grad=(opt.flat_g / ws).cpu().numpy(), world=ws)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_ranks_equal_one_rank():
    import torch.multiprocessing as mp
    out_path = os.path.join(tempfile.mkdtemp(), "two_rank.npz")
    port = 29600 + (os.getpid() % 300)
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out_path)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=900)
        assert p.exitcode == 0
    got = np.load(out_path)
    assert int(got["world"]) == 2
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    # ---- eval: bit-equal counts / distances
    one = _eval_rows(0, B_EVAL, dev).cpu().numpy()
    assert one.shape == got["rows"].shape == (B_EVAL, 4 * 99)
    assert np.array_equal(one, got["rows"]), "2-rank I / U / HD / MSD differ from the single-process result"
    # ---- training: all-reduced mean gradient == mean of the shard gradients
    g = None
    for shard in range(2):
        opt, _ = _train_grad(shard, dev)
        torch.cuda.synchronize()
        g = opt.flat_g.clone() if g is None else g + opt.flat_g
    g = (g / 2).cpu().numpy()
    err = np.abs(g - got["grad"]).max()
    scale = np.abs(g).max()
    print("2-rank gradient: max-abs err %.3e of max |g| %.3e (%d elements)" % (err, scale, g.size))
    assert err <= 1e-6 * scale + 1e-9
