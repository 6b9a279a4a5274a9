"""Double-buffered host -> device upload of the node-feature matrix.

The reference moves the feature matrix to the device inside every forward call (mrgcn.py:199-204) on the compute stream,
so the DMA and the step are strictly serial.  Here the copy of the NEXT step's matrix can be started ahead on a copy
stream (`MRGCN.prefetch(batch)`), into one of a few resident device buffers, while the current step computes; the forward
that consumes it only waits for the copy's event.  AM shape: 1.0 GB per step, 18 ms on PCIe against a 13 ms step.

Buffer life cycle: free -> pending (copy in flight or done) -> in use (consumed by the latest forward) -> released (a
later forward has begun, so the consumer's backward is already ahead of it on the compute stream; the event recorded
then is what the next copy into the buffer waits for) -> reused.  A buffer whose consumer is still the latest step is
never overwritten; when no buffer is available the prefetch is declined and the forward uploads in line as before.
"""
from __future__ import annotations

import torch


class _Slot:
    __slots__ = ("buf", "ready", "release", "state", "key", "seq")

    def __init__(self, buf):
        self.buf, self.ready, self.release, self.state, self.key, self.seq = buf, torch.cuda.Event(), None, "free", None, 0


def host_key(X):
    """Identity of a host matrix: the memory it lives in (the caller must not change it between prefetch and forward).
    The same matrix may be in flight more than once (one copy per step it will be used in)."""
    return (X.data_ptr(), tuple(X.shape), X.dtype)


class FeaturePrefetcher:
    def __init__(self, depth=3):
        self.depth, self.slots, self.stream, self.latest = depth, [], None, None
        self.copies = 0

    def _slot_for(self, X, dev):
        for s in self.slots:
            if s.state in ("free", "released") and s.buf.shape == X.shape and s.buf.dtype == X.dtype:
                return s
        if len(self.slots) < self.depth:
            s = _Slot(torch.empty(X.shape, dtype=X.dtype, device=dev))
            self.slots.append(s)
            return s
        return None

    def start(self, X, dev):
        """Begin copying host matrix X (pinned for a truly asynchronous DMA) into a resident buffer.  Returns False when
        every buffer is pending or in use."""
        slot = self._slot_for(X, dev)
        if slot is None:
            return False
        if self.stream is None:
            self.stream = torch.cuda.Stream(device=dev)
        if slot.release is not None:
            self.stream.wait_event(slot.release)       # the buffer's last consumer (forward and backward) is done
        else:
            self.stream.wait_stream(torch.cuda.current_stream(dev))     # fresh allocation: order after the allocator's work
        with torch.cuda.stream(self.stream):
            slot.buf.copy_(X, non_blocking=True)
            slot.ready.record(self.stream)
        self.copies += 1
        slot.state, slot.key, slot.release, slot.seq = "pending", host_key(X), None, self.copies
        return True

    def take(self, X, dev):
        """The device copy of X if one was prefetched (the compute stream is made to wait for it), else None."""
        key = host_key(X)
        for s in sorted(self.slots, key=lambda t: t.seq):          # oldest copy of this matrix first
            if s.state == "pending" and s.key == key:
                main = torch.cuda.current_stream(dev)
                if self.latest is not None and self.latest.state == "in use":
                    ev = torch.cuda.Event()
                    ev.record(main)                     # everything of the previous consumer's step is ahead of this point
                    self.latest.release, self.latest.state = ev, "released"
                main.wait_event(s.ready)
                s.state, self.latest = "in use", s
                return s.buf
        return None
