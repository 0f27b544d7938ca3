"""Host -> device prefetch of evaluation / training batches (the role of DataLoader(pin_memory=True) + `.to("cuda")` in the
reference's loops, model/engine/inference.py:88-90, trainer.py:57-60): the copy of batch i+1 runs on its own stream while
batch i is being processed, so the PCIe transfer (205 MB per 64 x 448^2 batch) is hidden behind the step."""
import torch


class DevicePrefetcher:
    """Iterates over `batches` (an iterable of tuples of PINNED host tensors) and yields tuples of device tensors, keeping one
    batch in flight.  Two sets of device buffers alternate; a yielded set stays valid until the batch after the next one is
    requested (events guard the reuse)."""

    def __init__(self, batches, device):
        self.it = iter(batches)
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.bufs = [None, None]
        self.done = [None, None]           # copy-finished events
        self.free = [None, None]           # consumer-finished events (recorded when the following batch is requested)
        self.slot = 0
        self.pending = None
        self._issue()

    def _issue(self):
        try:
            host = next(self.it)
        except StopIteration:
            self.pending = None
            return
        k = self.slot
        self.slot ^= 1
        if self.bufs[k] is None or any(b.shape != h.shape or b.dtype != h.dtype for b, h in zip(self.bufs[k], host)):
            self.bufs[k] = tuple(torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host)
        with torch.cuda.stream(self.stream):
            if self.free[k] is not None:
                self.stream.wait_event(self.free[k])            # the consumer of this buffer set has finished with it
            for b, h in zip(self.bufs[k], host):
                b.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.done[k] = ev
        self.pending = k

    def __iter__(self):
        return self

    def __next__(self):
        if self.pending is None:
            raise StopIteration
        k = self.pending
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.done[k])                             # this batch has arrived
        other = k ^ 1
        ev = torch.cuda.Event()
        ev.record(cur)                                           # everything queued so far used the other set at most
        self.free[other] = ev
        self._issue()                                            # start the next copy (into the other set) right away
        return self.bufs[k]
