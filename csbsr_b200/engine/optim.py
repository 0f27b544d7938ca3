"""Optimizer side of the training step (reference train.py:91-96, trainer.py:57-72): Adam on flat buffers with one
fused kernel per step (csrc/train.cu), the LambdaLR(UpDownScheduler) factor, and the NCCL gradient all-reduce that
replaces nn.DataParallel's scatter / gather (train.py:117-121)."""
import torch

from .. import _lib


class UpDownScheduler:
    """model/utils/lr_scheduler.py:31-43: x10 between main-phase iterations 70000 and 95000 when enabled."""

    def __init__(self, pretrain_iter, resume_iter, scheduler_flag):
        self.pretrain_iter, self.resume_iter, self.scheduler_flag = pretrain_iter, resume_iter, scheduler_flag

    def __call__(self, _iter):
        main = _iter - (self.pretrain_iter - 1) + self.resume_iter
        return 10 if (70000 < main < 95000 and self.scheduler_flag) else 1


def active_runs(active, slots, steps):
    """Merge the flat-buffer slots of the active parameters into launch runs: adjacent slots whose parameters share the
    same (already incremented) step counter become one [offset, length, step] run.  Pure host logic (CPU-testable)."""
    runs, cur = [], None
    for on, (off, sz), step in zip(active, slots, steps):
        if not on:
            cur = None
            continue
        if cur is not None and cur[2] == step and cur[0] + cur[1] == off:
            cur[1] += sz
        else:
            cur = [off, sz, step]
            runs.append(cur)
    return runs


class FusedAdam:
    """torch.optim.Adam(params, lr, betas=(0.9, 0.999), eps=1e-8) semantics (no weight decay / amsgrad).

    All parameters are re-pointed at views of ONE flat fp32 buffer, their .grad at views of a second one, so that
    a step is a single elementwise kernel over the flat buffers (gradient zeroing for the next step fused in) and the
    data-parallel gradient exchange is a handful of large NCCL all-reduces on slices of the flat gradient."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8, lr_lambda=None, bucket_elems=1 << 25):
        self.params = [p for p in params if p.requires_grad]
        assert self.params and all(p.is_cuda and p.dtype == torch.float32 for p in self.params)
        dev = self.params[0].device
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]            # 16-byte aligned slots
        self.n = sum(sizes)
        self.flat_p = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.n, dtype=torch.float32, device=dev)
        off = 0
        self.slots = []                       # (offset, padded size) of every parameter in the flat buffers
        for p, sz in zip(self.params, sizes):
            view = self.flat_p[off:off + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_g[off:off + p.numel()].view(p.shape)
            self.slots.append((off, sz))
            off += sz
        from .. import kernels as K
        K.register_params(self.params)              # packed bf16 operands are cached per step; wgrad accumulates into .grad
        K.invalidate_packed()
        self.param_steps = [0] * len(self.params)   # torch.optim.Adam keeps one step counter per parameter
        self.base_lr, self.betas, self.eps = lr, betas, eps
        self.lr_lambda = lr_lambda
        self.step_count = 0                   # optimizer steps taken
        self.sched_count = 0                  # LambdaLR epoch counter (scheduler.step() calls)
        self.bucket_elems = bucket_elems

    @property
    def lr(self):
        return self.base_lr * (self.lr_lambda(self.sched_count) if self.lr_lambda else 1.0)

    def zero_grad(self):
        """Gradients are cleared inside step(); before the first step they are already zero."""
        return None

    def all_reduce_grads(self, group=None):
        """SUM all-reduce of the flat gradient in large buckets; step() then scales by 1 / world_size."""
        from .distributed import allreduce_flat
        return allreduce_flat(self.flat_g, self.bucket_elems, group)

    def step(self, world_size=1):
        """One Adam step on every parameter with requires_grad (frozen ones are skipped like torch's `grad is None`:
        neither their value nor their moments nor their step counter move).  Adjacent active parameters with the same step
        counter are updated by one launch -- a single launch over the whole buffer outside the pre-training phases."""
        self.step_count += 1
        active = [p.requires_grad for p in self.params]
        for i, on in enumerate(active):
            if on:
                self.param_steps[i] += 1
        for off, n, step in active_runs(active, self.slots, self.param_steps):
            rc = _lib.lib().csbsr_adam_step(self.flat_p.data_ptr() + 4 * off, self.flat_g.data_ptr() + 4 * off,
                                            self.exp_avg.data_ptr() + 4 * off, self.exp_avg_sq.data_ptr() + 4 * off, n, self.lr,
                                            self.betas[0], self.betas[1], self.eps, step, 1.0 / world_size, 1,
                                            _lib.stream_ptr())
            _lib.check(rc, "csbsr_adam_step")
            _lib.count_launch("csbsr_adam_step")
        from .. import kernels as K
        K.invalidate_packed()

    def scheduler_step(self):
        self.sched_count += 1

    def state_dict(self):
        """Moments, per-parameter step counters and schedule position (the reference saves optimizer.state_dict() next to
        every checkpoint, trainer.py:119-129); flat layout: slot i of `slots` belongs to parameter i."""
        return {"exp_avg": self.exp_avg.detach().cpu(), "exp_avg_sq": self.exp_avg_sq.detach().cpu(),
                "param_steps": list(self.param_steps), "slots": list(self.slots), "step_count": self.step_count,
                "sched_count": self.sched_count, "lr": self.base_lr, "betas": self.betas, "eps": self.eps}

    def load_state_dict(self, sd):
        assert list(sd["slots"]) == list(self.slots), "optimizer state belongs to a different parameter layout"
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.param_steps = list(sd["param_steps"])
        self.step_count, self.sched_count = sd["step_count"], sd["sched_count"]
