"""Multi-GPU use of the hot path: one process per GPU (torch.distributed, NCCL over NVLink).

The reference has no distributed code at all.  Rays and query points are
independent given replicated read-only state (grids, decoders, TSDF), so

* mapping (render_batch_ray fwd+bwd): rays are sharded; two collectives per step -- an all-reduce(max)
  of max(gt_depth) before sampling (the far clamp and the depth==0 sampling use the batch-global
  maximum, src/utils/Renderer.py:159,195) and an all-reduce(sum) of the grid + packed decoder
  gradients inside the backward;
* render_img and the Mesher dense query shard by chunk / voxel block with no data-path collective;
  results are gathered on request.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import functional as Fn


def shard_range(n: int, rank: int, world: int):
    """Contiguous [lo, hi) slice of n items owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def chunk_aligned_range(n: int, chunk: int, rank: int, world: int):
    """Like shard_range but on whole chunks, so per-chunk depth maxima stay exact (render_img)."""
    nchunks = (n + chunk - 1) // chunk
    lo, hi = shard_range(nchunks, rank, world)
    return min(lo * chunk, n), min(hi * chunk, n)


def interleaved_shard(n: int, rank: int, world: int) -> torch.Tensor:
    """Indices rank, rank + world, rank + 2*world, ... < n.  Dealing a mapping batch (rays of several keyframes, frame
    after frame) to the ranks this way gives every rank the same mix of frames; a contiguous split (``shard_range``)
    hands each rank whole frames, and the cost of a ray depends on its frame's pose (how much of it lies inside the
    bound / the TSDF band), which showed up as a 15 % load imbalance at 8 GPUs."""
    return torch.arange(rank, n, world, dtype=torch.int64)


def global_depth_max(gt_depth: torch.Tensor, group) -> torch.Tensor:
    import torch.distributed as dist
    m = gt_depth.detach().float().max().reshape(1)
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    return m


def render_batch_ray_sharded(renderer, c, decoders, rays_d, rays_o, tsdf_volume, tsdf_bnds, stage, gt_depth=None,
                             group=None, depth_max=None, overlap_allreduce=True):
    """``renderer.render_batch_ray`` on this rank's shard of a batch that is spread over ``group``.

    Outputs are the shard's; gradients of grids and decoders come back already summed over the ranks
    (identical on every rank, so identical optimiser steps keep the replicas in sync).

    ``depth_max``: max(gt_depth) of the WHOLE batch (the far clamp and the depth==0 sampling use it,
    src/utils/Renderer.py:159,195).  A Mapper whose ranks all hold the keyframes computes it locally before dealing the
    rays out and passes it here; without it the ranks agree on it with an all-reduce(max) -- one more collective and one
    more synchronisation point per step.  The gradient all-reduce runs in three buckets on a side stream, each bucket as
    soon as the backward has finished it (``overlap_allreduce=False``: one blocking all-reduce after the backward)."""
    if group is None:
        return renderer.render_batch_ray(c, decoders, rays_d, rays_o, rays_o.device, tsdf_volume, tsdf_bnds, stage,
                                         gt_depth=gt_depth)
    if gt_depth is None:
        dmax = None
    elif depth_max is not None:
        dmax = torch.as_tensor(depth_max, dtype=torch.float32, device=rays_o.device).reshape(1)
    else:
        dmax = global_depth_max(gt_depth, group)
    if rays_o.shape[0] == 0:
        raise ValueError("this rank's shard is empty: deal at least one ray to every rank (interleaved_shard / shard_range "
                         "of a batch with n >= world size), the other ranks would wait forever in the gradient all-reduce")
    return Fn.render_batch_ray(c, decoders, rays_d, rays_o, tsdf_volume, tsdf_bnds, renderer.bound, stage,
                               gt_depth=gt_depth, n_samples=renderer.N_samples, n_surface=renderer.N_surface,
                               n_importance=renderer.N_importance, depth_max=dmax, group=group,
                               overlap_allreduce=overlap_allreduce)


def gather_concat(x: torch.Tensor, sizes, group, dst: Optional[int] = None):
    """Concatenate per-rank pieces of different lengths (dim 0) on every rank (dst=None) or on ``dst`` only."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mx = max(sizes)
    pad = x.new_zeros((mx,) + tuple(x.shape[1:]))
    pad[:x.shape[0]] = x
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    if dst is not None and dist.get_rank(group) != dst:
        return None
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def render_img_sharded(renderer, c, decoders, c2w, device, tsdf_volume, tsdf_bnds, stage, gt_depth, group=None,
                       gather=True):
    """Full-frame render with the rays dealt evenly to the ranks, identical to the single-GPU frame bit for bit.

    The reference renders ``ray_batch_size`` chunks, each with its own max(gt_depth) (src/utils/Renderer.py:294-306, 159,
    195).  Whole chunks per rank would leave the frame's 9 chunks (816 000 rays / 100 000) on 8 ranks with one rank doing
    two of them; instead every rank takes n / world consecutive rays, cut at the chunk boundaries it straddles, and each
    piece is rendered with ITS chunk's maximum (computed once over the whole frame, on the device) passed in."""
    from .common import get_rays
    if group is None:
        return renderer.render_img(c, decoders, c2w, device, tsdf_volume, tsdf_bnds, stage, gt_depth=gt_depth)
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    H, W = renderer.H, renderer.W
    rb = renderer.ray_batch_size
    with torch.no_grad():
        rays_o, rays_d = get_rays(H, W, renderer.fx, renderer.fy, renderer.cx, renderer.cy, c2w, device)
        rays_o, rays_d, gt = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3), gt_depth.reshape(-1).float()
        n = rays_d.shape[0]
        nchunks = (n + rb - 1) // rb
        pad = nchunks * rb - n
        gtp = torch.cat([gt, gt.new_full((pad,), float("-inf"))]) if pad else gt
        chunk_max = gtp.reshape(nchunks, rb).max(dim=1)[0]                       # per-chunk max(gt_depth), device, no sync
        lo, hi = shard_range(n, rank, world)
        ds, us, cs = [], [], []
        a = lo
        while a < hi:
            ck = a // rb
            b = min(hi, (ck + 1) * rb)
            d, u, col, _ = Fn.render_batch_ray(c, decoders, rays_d[a:b], rays_o[a:b], tsdf_volume, tsdf_bnds, renderer.bound, stage,
                                               gt_depth=gt[a:b], n_samples=renderer.N_samples, n_surface=renderer.N_surface,
                                               depth_max=chunk_max[ck:ck + 1])
            ds.append(d); us.append(u); cs.append(col)
            a = b
        if ds:
            d, u, col = torch.cat(ds), torch.cat(us), torch.cat(cs)
        else:
            d = torch.empty(0, device=device, dtype=torch.float64)
            u = torch.empty(0, device=device, dtype=torch.float64)
            col = torch.empty(0, 3, device=device)
        if not gather:
            return (lo, hi), d, u, col
        sizes = [b_ - a_ for a_, b_ in (shard_range(n, r, world) for r in range(world))]
        d = gather_concat(d, sizes, group)
        u = gather_concat(u, sizes, group)
        col = gather_concat(col, sizes, group)
        return d.reshape(H, W), u.reshape(H, W), col.reshape(H, W, 3)
