"""Graph-captured stepping: one training step recorded once into a CUDA graph and replayed.

Full-batch training as the reference runs it (node_classification.py:170-200, link_prediction.py:244-330) repeats the SAME
step every epoch - same graph, same shapes, same buffers - so the ~100 kernel launches (and, under a node partition, the
collectives between them) can be replayed with one host call.  That removes the host launch overhead which dominates
once the partitioned step drops to a few milliseconds.

    step = GraphedStep(lambda: loss_fn(model.rgcn(X, graph)), model.parameters())
    for epoch in range(n):
        loss = step()                 # replays zero-grad + forward + loss + backward; .grad tensors are updated in place
        optimizer.step()              # outside the graph (FusedClipAdam bakes the step count into its launch)

Rules of CUDA graph capture apply: the closure must launch the same work every time (no host-side data-dependent control
flow, no .item()), inputs are read from the tensors captured (update them in place with .copy_()), and the gradients
land in the graph's own tensors, which every call re-attaches as .grad (zeroing is part of the captured step).
"""
from __future__ import annotations

import torch


class GraphedStep:
    def __init__(self, loss_fn, params, after_backward=None, warmup=2, barrier=None):
        """loss_fn() -> scalar loss tensor (forward only; backward is called here).  params: the tensors whose .grad the
        step produces.  after_backward: called after loss.backward() inside the step (e.g. PartitionedRGCN.sync_grads).
        barrier: called between the eager warm-up and the capture (pass the process group's barrier under torchrun, so that
        no rank captures while another still runs eager collectives)."""
        self.params = [p for p in params if p.requires_grad]
        self.loss_fn, self.after_backward = loss_fn, after_backward
        self.graph, self.loss = None, None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):             # warm-up off the default stream: allocator pools, lazy work lists, NCCL
            for _ in range(max(1, warmup)):
                self.eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if barrier is not None:
            barrier()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self.eager()
        self.grads = [p.grad for p in self.params]      # the graph's own gradient tensors (fixed addresses)

    def eager(self):
        """The step as plain launches (what the graph records)."""
        for p in self.params:
            p.grad = None                         # inside the capture: the graph's private pool re-creates them at fixed addresses
        loss = self.loss_fn()
        loss.backward()
        if self.after_backward is not None:
            self.after_backward()
        return loss

    def __call__(self):
        self.graph.replay()
        for p, g in zip(self.params, self.grads):        # re-attach, in case the caller dropped or replaced .grad
            p.grad = g
        return self.loss
