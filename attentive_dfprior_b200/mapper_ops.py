"""Mapper-side operations around the render path, on the CUDA library.

* ``MaskedGridAdam``      -- the frustum-masked grid optimiser of ``Mapper.optimize_map``
                             (src/Mapper.py:351-361, 374-388, 472-482) as one in-place kernel per grid
* ``get_mask_from_c2w``   -- ``Mapper.get_mask_from_c2w`` (src/Mapper.py:90-158) without numpy / cv2.remap
* ``assemble_rays``       -- ``get_samples`` per frame + bounding-box pre-filter (src/Mapper.py:408-449,
                             src/common.py:76-136); the torch.randint draw stays with the caller
* ``mapper_loss``         -- the Mapper's loss (src/Mapper.py:457-469) with its gradient, differentiable

No CPU fallback: every function raises if the CUDA library is missing or a tensor is not on the GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .functional import _require_cuda, _stream


def _is_ndhwc(g: torch.Tensor) -> bool:
    return g.dim() == 5 and g.shape[0] == 1 and g.is_contiguous(memory_format=torch.channels_last_3d)


class MaskedGridAdam:
    """Adam over the voxels of a frustum mask, updating the dense grids in place.

    Upstream builds, for every grid, ``val_grad = val[mask]`` as the optimised leaf, re-scatters it into the dense grid
    before each iteration, lets autograd's index_put backward compact the dense gradient, steps one torch.optim.Adam
    over all groups and writes the result back (Mapper.py:351-361, 383-388, 472-482).  Here the grids themselves stay
    the leaves (channels-last ``[1,C,D,H,W]``, so the render kernels read them in place) and ``step`` runs Adam's
    arithmetic directly on the masked voxels; moments are dense buffers that only the masked voxels ever touch.
    The result equals the upstream flow: voxels outside the mask never change, their gradients are discarded.

    grids: {'grid_low': tensor, ...};  masks: {'grid_low': bool/uint8 [D,H,W] (device) or None = all voxels}.
    A fresh instance per ``optimize_map`` call, like upstream's fresh optimizer (Mapper.py:374).
    """

    def __init__(self, grids: dict, masks: dict | None = None, betas=(0.9, 0.999), eps: float = 1e-8):
        self.grids = grids
        self.betas = betas
        self.eps = eps
        self.state = {}
        masks = masks or {}
        for key, g in grids.items():
            _require_cuda(g)
            if not _is_ndhwc(g) or g.dtype != torch.float32:
                raise L.DfpError(f"{key}: MaskedGridAdam updates in place and needs a float32 [1,C,D,H,W] grid in "
                                 "torch.channels_last_3d memory format")
            m = masks.get(key)
            if m is not None:
                if m.dim() == 5:                                   # upstream's repeated [1,C,D,H,W] mask
                    m = m[0, 0]
                if tuple(m.shape) != tuple(g.shape[2:]):
                    raise L.DfpError(f"{key}: mask shape {tuple(m.shape)} does not match the grid {tuple(g.shape[2:])}")
                m = m.to(device=g.device, dtype=torch.uint8).contiguous()
            self.state[key] = {"step": 0, "mask": m,
                               "exp_avg": torch.zeros_like(g, memory_format=torch.preserve_format),
                               "exp_avg_sq": torch.zeros_like(g, memory_format=torch.preserve_format)}

    @torch.no_grad()
    def step(self, lrs: dict, grads: dict | None = None, set_to_none: bool = True):
        """One optimiser step.  lrs: {'grid_low': lr, ...} (cfg['mapping']['stage'][stage][...]*lr_factor,
        Mapper.py:392-398).  grads default to ``grid.grad``; grids without a gradient are skipped, as torch's Adam
        skips parameters whose .grad is None (their step count does not advance)."""
        lib = L.lib()
        for key, g in self.grids.items():
            grad = grads.get(key) if grads is not None else g.grad
            if grad is None:
                continue
            if not _is_ndhwc(grad):
                grad = grad.contiguous(memory_format=torch.channels_last_3d)
            st = self.state[key]
            st["step"] += 1
            a = L.DfpAdam()
            a.param, a.grad = g.data_ptr(), grad.data_ptr()
            a.exp_avg, a.exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            a.voxel_mask = None if st["mask"] is None else st["mask"].data_ptr()
            a.n_vox = g.shape[2] * g.shape[3] * g.shape[4]
            a.C = g.shape[1]
            a.step = st["step"]
            a.lr, a.beta1, a.beta2, a.eps = float(lrs[key]), self.betas[0], self.betas[1], self.eps
            L.check(lib.dfp_masked_adam_step(C.byref(a), _stream()))
            if grads is None and set_to_none:
                g.grad = None                                      # optimizer.zero_grad() (Mapper.py:473)


def frustum_mask(c2w, bound, val_shape, depth: torch.Tensor, H, W, fx, fy, cx, cy) -> torch.Tensor:
    """Frustum mask as a uint8 device tensor in the grid's voxel order [D,H,W]."""
    _require_cuda(depth)
    dev = depth.device
    D, Hg, Wg = (int(v) for v in val_shape)
    depth = depth.to(torch.float32).contiguous()
    pose = c2w.detach().cpu().numpy() if torch.is_tensor(c2w) else np.asarray(c2w)
    w2c = np.linalg.inv(pose)                                      # f32 in, f32 out: what Mapper.py:113 computes
    f = L.DfpFrustum()
    f.depth, f.H, f.W = depth.data_ptr(), int(depth.shape[0]), int(depth.shape[1])
    f.fx, f.fy, f.cx, f.cy = float(fx), float(fy), float(cx), float(cy)
    f.w2c[:] = [float(v) for v in np.asarray(w2c, dtype=np.float32).reshape(-1)]
    f.cam_o[:] = [float(v) for v in np.asarray(pose, dtype=np.float32)[:3, 3]]
    # vertex coordinates exactly as upstream builds them (torch.linspace in f32, Mapper.py:106-108)
    axes = [torch.linspace(bound[k][0], bound[k][1], n).to(dev) for k, n in ((0, Wg), (1, Hg), (2, D))]
    f.xs, f.ys, f.zs = (a.data_ptr() for a in axes)
    f.Dg, f.Hg, f.Wg = D, Hg, Wg
    lib = L.lib()
    nbytes = lib.dfp_frustum_scratch_bytes(D, Hg, Wg)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    mask = torch.empty((D, Hg, Wg), dtype=torch.uint8, device=dev)
    L.check(lib.dfp_frustum_mask(C.byref(f), mask.data_ptr(), scratch.data_ptr(), nbytes, _stream()))
    return mask


def get_mask_from_c2w(c2w, key, val_shape, depth_np, bound, H, W, fx, fy, cx, cy, device="cuda:0"):
    """Drop-in for ``Mapper.get_mask_from_c2w(c2w, key, val_shape, depth_np)`` (src/Mapper.py:90): returns the numpy
    bool mask shaped (val_shape[2], val_shape[1], val_shape[0]) that Mapper.py:347-348 permutes and repeats."""
    depth = torch.as_tensor(depth_np, dtype=torch.float32).to(device)
    m = frustum_mask(c2w, bound, val_shape, depth, H, W, fx, fy, cx, cy)
    return m.permute(2, 1, 0).bool().cpu().numpy()


def assemble_rays(frames, indices: torch.Tensor, H, W, fx, fy, cx, cy, bound, window=None, filter: bool = True):
    """``get_samples`` for every frame + cat + bounding-box pre-filter (src/Mapper.py:408-449).

    frames: [(depth [H,W] f32, color [H,W,3] f32, c2w [4,4])] on the GPU; indices int64 [n_frames, n] = the
    ``torch.randint(window_h * window_w, (n,))`` draws of select_uv (common.py:101) in frame order.
    Returns (rays_o, rays_d, gt_depth, gt_color) of the surviving rays, in order.  One scalar device->host read
    (the survivor count) sizes the outputs.
    """
    H0, H1, W0, W1 = window or (0, H, 0, W)
    indices = indices.to(torch.int64).contiguous()
    _require_cuda(indices)
    dev = indices.device
    nf, n = indices.shape
    keep = []
    arr = (L.DfpFrame * nf)()
    for k, (depth, color, c2w) in enumerate(frames):
        _require_cuda(depth, color)
        depth = depth.to(torch.float32).contiguous()
        color = color.to(torch.float32).contiguous()
        keep += [depth, color]
        arr[k].depth, arr[k].color = depth.data_ptr(), color.data_ptr()
        pose = c2w.detach().to("cpu", torch.float32)
        arr[k].c2w[:] = [float(v) for v in pose[:3, :4].reshape(-1)]
    b = L.DfpRayBatch()
    b.frames, b.n_frames, b.indices, b.n_per_frame = arr, nf, indices.data_ptr(), n
    b.H0, b.H1, b.W0, b.W1, b.W = H0, H1, W0, W1, W
    b.fx, b.fy, b.cx, b.cy = float(fx), float(fy), float(cx), float(cy)
    L.fill_bounds(b.bound, bound)
    b.filter = 1 if filter else 0
    total = nf * n
    out = L.DfpRayBatchOut()
    rays_o = torch.empty((total, 3), dtype=torch.float32, device=dev)
    rays_d = torch.empty((total, 3), dtype=torch.float32, device=dev)
    gt_depth = torch.empty((total,), dtype=torch.float32, device=dev)
    gt_color = torch.empty((total, 3), dtype=torch.float32, device=dev)
    n_kept = torch.zeros(1, dtype=torch.int32, device=dev)
    out.rays_o, out.rays_d, out.gt_depth, out.gt_color, out.n_kept = (t.data_ptr() for t in (rays_o, rays_d, gt_depth, gt_color, n_kept))
    lib = L.lib()
    nbytes = lib.dfp_assemble_scratch_bytes(total)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    L.check(lib.dfp_assemble_rays(C.byref(b), C.byref(out), scratch.data_ptr(), nbytes, _stream()))
    k = int(n_kept.item())
    return rays_o[:k], rays_d[:k], gt_depth[:k], gt_color[:k]


class _MapperLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, color, weight, gt_depth, gt_color, w_color, use_weight, stage_color):
        _require_cuda(depth, gt_depth)
        n = depth.shape[0]
        dev = depth.device
        depth = depth.contiguous()
        l = L.DfpMapLoss()
        l.depth, l.gt_depth = depth.data_ptr(), gt_depth.contiguous().data_ptr()
        l.n = n
        loss = torch.zeros((), dtype=torch.float64, device=dev)
        g_depth = torch.empty_like(depth)
        g_color = g_weight = None
        hold = [depth]
        if stage_color:
            color, gt_color = color.contiguous(), gt_color.contiguous()
            g_color = torch.empty_like(color)
            l.color, l.gt_color, l.g_color, l.w_color = color.data_ptr(), gt_color.data_ptr(), g_color.data_ptr(), float(w_color)
            hold += [color, gt_color]
        if use_weight:
            weight = weight.contiguous()
            g_weight = torch.empty_like(weight)
            l.weight, l.g_weight, l.use_weight, l.S = weight.data_ptr(), g_weight.data_ptr(), 1, weight.numel() // n
            hold.append(weight)
        l.g_depth, l.loss = g_depth.data_ptr(), loss.data_ptr()
        L.check(L.lib().dfp_mapper_loss(C.byref(l), _stream()))
        ctx.save_for_backward(g_depth, g_color, g_weight)
        return loss

    @staticmethod
    def backward(ctx, g):
        g_depth, g_color, g_weight = ctx.saved_tensors
        return (g_depth * g, None if g_color is None else g_color * g.to(g_color.dtype),
                None if g_weight is None else g_weight * g.to(g_weight.dtype), None, None, None, None, None)


def mapper_loss(depth, color, weight, gt_depth, gt_color, w_color: float, use_weight: bool, stage_color: bool):
    """Mapper loss (src/Mapper.py:457-469) as one kernel that also produces the gradient; returns an f64 scalar."""
    return _MapperLoss.apply(depth, color, weight, gt_depth, gt_color, w_color, use_weight, stage_color)
