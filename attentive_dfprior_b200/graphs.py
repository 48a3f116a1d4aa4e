"""CUDA-graph capture of a whole iteration around the hot path.

At the reference's own batch sizes (Mapper: 1000 pixels per iteration, configs/Replica/replica.yaml:23; Tracker: 200,
:12-13) an iteration is about 25 dependent kernel launches of 10-60 us each plus the host work of issuing them.  The
library makes no host synchronisation, allocates nothing itself and takes every launch parameter from its call's
arguments, so the whole iteration -- render_batch_ray forward, loss, backward, optionally the optimiser step -- can be
captured once into a CUDA graph and replayed with one launch:

    step = GraphedStep(iteration, static_inputs=(rays_o, rays_d, gt_depth, gt_color))
    for batch in batches:
        loss = step(*batch)            # copies the batch into the static inputs, replays, returns the static outputs

Rules (those of torch.cuda.graphs): shapes are fixed at capture; ``iteration`` must not synchronise with the host
(no ``.item()``, no boolean-mask indexing -- multiply by the mask instead, the sums are the same); ``iteration`` itself
sets the ``.grad`` of everything it differentiates to None at its start (the warm-up runs below would otherwise leave
gradients that the captured backward accumulates into, replay after replay); gradients are then read from the ``.grad``
tensors the capture produced.
"""
from __future__ import annotations

import torch


class GraphedStep:
    def __init__(self, fn, static_inputs=(), warmup: int = 3):
        self.fn = fn
        self.static_inputs = tuple(static_inputs)
        dev = self.static_inputs[0].device if self.static_inputs else torch.device("cuda", torch.cuda.current_device())
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                       # warm-up off the capture: lazy initialisations, caches, cuda module loads
            for _ in range(warmup):
                fn(*self.static_inputs)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_outputs = fn(*self.static_inputs)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_inputs, inputs):
            if dst is not src:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_outputs
