"""GPU parity of the Mapper-side kernels (csrc/dfp_mapper.cu, through attentive_dfprior_b200.mapper_ops -> ctypes ->
C ABI) against the reference-generated fixtures and the CPU oracle.  Masks / compaction order / survivor sets are
bit-exact; float results within 1e-6 (Adam, rays) -- tighter than the path's 1e-4 forward tolerance."""
import os

import numpy as np
import pytest
import torch

from oracle import mapper_oracle as MO

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = torch.load(os.path.join(ROOT, "tests", "golden", "mapper_rows.pt"))


def cam():
    return tuple(GOLD[k] for k in ("H", "W", "fx", "fy", "cx", "cy"))


@pytest.mark.parametrize("k", range(len(GOLD["masks"])))
def test_frustum_mask_fixture(k):
    from attentive_dfprior_b200 import mapper_ops as M
    m = GOLD["masks"][k]
    H, W, fx, fy, cx, cy = cam()
    got = M.get_mask_from_c2w(m["c2w"], "grid_low", m["val_shape"], m["depth"].numpy(), GOLD["bound"], H, W, fx, fy, cx, cy, device=DEV)
    ref = m["mask"].numpy()
    assert got.shape == ref.shape and got.dtype == np.bool_
    assert np.array_equal(got, ref)


def test_frustum_mask_room0_vs_oracle():
    """full-size grids of the Replica room0 configuration, 680x1200 depth image"""
    from attentive_dfprior_b200 import mapper_ops as M
    bound = torch.tensor([[-2.9, 8.94], [-3.2, 5.76], [-3.5, 3.54]], dtype=torch.float64)
    H, W, fx, fy, cx, cy = 680, 1200, 600.0, 600.0, 599.5, 339.5
    g = torch.Generator().manual_seed(5)
    depth = (0.5 + 4.0 * torch.rand(H, W, generator=g)).float()
    depth[torch.rand(H, W, generator=g) < 0.1] = 0
    c2w = torch.eye(4)
    c2w[:3, 3] = torch.tensor([3.0, 1.2, 0.1])
    ang = 0.7
    c2w[0, 0], c2w[0, 2], c2w[2, 0], c2w[2, 2] = np.cos(ang), np.sin(ang), -np.sin(ang), np.cos(ang)
    for val_shape in ((21, 28, 37), (43, 56, 74)):
        ref = MO.get_mask_from_c2w(c2w, bound, val_shape, depth.numpy(), H, W, fx, fy, cx, cy)
        got = M.get_mask_from_c2w(c2w, "grid_high", val_shape, depth.numpy(), bound, H, W, fx, fy, cx, cy, device=DEV)
        assert 0.02 < ref.mean() < 0.98
        assert np.array_equal(got, ref)


@pytest.mark.parametrize("filt", [False, True])
def test_assemble_rays_fixture(filt):
    from attentive_dfprior_b200 import mapper_ops as M
    H, W, fx, fy, cx, cy = cam()
    for b in GOLD["batches"]:
        frames = [(b["depth"].to(DEV), b["color"].to(DEV), b["c2w"])]
        got = M.assemble_rays(frames, b["indices"][None].to(DEV), H, W, fx, fy, cx, cy, GOLD["bound"], window=b["window"], filter=filt)
        ref = [b["rays_o"], b["rays_d"], b["gt_depth"], b["gt_color"]]
        if filt:
            keep = MO.box_filter(b["rays_o"], b["rays_d"], b["gt_depth"], GOLD["bound"])
            assert 0 < keep.sum() < keep.numel()
            ref = [r[keep] for r in ref]
        for x, r in zip(got, ref):
            assert x.shape == r.shape                                 # same survivors ...
            assert torch.allclose(x.cpu(), r, rtol=1e-6, atol=1e-7)   # ... in the same order
        assert torch.equal(got[2].cpu(), ref[2]) and torch.equal(got[3].cpu(), ref[3])      # gathered values: exact


def test_assemble_rays_multi_frame_large():
    """5 frames x 20 000 pixels (multi-block compaction) against the oracle"""
    from attentive_dfprior_b200 import mapper_ops as M
    H, W, fx, fy, cx, cy = cam()
    g = torch.Generator().manual_seed(9)
    frames, idx = [], []
    for f in range(5):
        b = GOLD["batches"][f % 2]
        c2w = b["c2w"].clone()
        c2w[:3, 3] += 0.05 * f
        frames.append(((0.3 + 3 * torch.rand(H, W, generator=g)).float(), torch.rand(H, W, 3, generator=g).float(), c2w))
        idx.append(torch.randint(H * W, (20000,), generator=g))
    idx = torch.stack(idx)
    ref = MO.assemble_batch(frames, idx, H, W, fx, fy, cx, cy, GOLD["bound"])
    got = M.assemble_rays([(d.to(DEV), c.to(DEV), p) for d, c, p in frames], idx.to(DEV), H, W, fx, fy, cx, cy, GOLD["bound"])
    assert 0 < ref[0].shape[0] < 100000
    for x, r in zip(got, ref):
        assert x.shape == r.shape
        assert torch.allclose(x.cpu(), r, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("masked", [True, False])
def test_masked_adam_fixture(masked):
    from attentive_dfprior_b200 import mapper_ops as M
    a = GOLD["adam"]
    grid = a["val"].to(DEV).contiguous(memory_format=torch.channels_last_3d)
    opt = M.MaskedGridAdam({"grid_low": grid}, {"grid_low": a["mask"].to(DEV) if masked else None})
    for it, lr in enumerate(a["lrs"]):
        g = a["grads"][it]
        opt.step({"grid_low": lr}, grads={"grid_low": None if g is None else g.to(DEV)})
    ref = a["final_masked" if masked else "final_all"]
    assert torch.allclose(grid.cpu(), ref, rtol=1e-6, atol=1e-8)
    if masked:
        outside = ~a["mask"][None, None].expand_as(a["val"])
        assert torch.equal(grid.cpu()[outside], a["val"][outside])   # untouched voxels: bit-exact


def test_masked_adam_through_autograd_room0_size():
    """grid.grad path, room0 high-grid size, 3 steps vs the oracle flow"""
    from attentive_dfprior_b200 import mapper_ops as M
    g = torch.Generator().manual_seed(3)
    shape = (1, 32, 43, 56, 74)
    val = (0.01 * torch.randn(shape, generator=g)).float()
    vmask = torch.rand(shape[2:], generator=g) < 0.3
    grads = [(0.1 * torch.randn(shape, generator=g)).float() for _ in range(3)]
    lrs = [0.005, 0.0, 0.01]
    ref = MO.masked_adam_run(val, vmask, lambda v, it: grads[it], lrs)
    grid = val.to(DEV).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    opt = M.MaskedGridAdam({"grid_high": grid}, {"grid_high": vmask.to(DEV)})
    for it, lr in enumerate(lrs):
        (grid * grads[it].to(DEV)).sum().backward()
        opt.step({"grid_high": lr})
        assert grid.grad is None
    assert torch.allclose(grid.detach().cpu(), ref, rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("use_weight,stage_color", [(False, False), (True, False), (False, True), (True, True)])
def test_mapper_loss_and_grad(use_weight, stage_color):
    from attentive_dfprior_b200 import mapper_ops as M
    g = torch.Generator().manual_seed(21)
    n, S = 3000, 48
    gt_depth = (3 * torch.rand(n, generator=g)).float()
    gt_depth[torch.rand(n, generator=g) < 0.1] = 0
    gt_color = torch.rand(n, 3, generator=g)
    depth = (gt_depth.double() + 0.3 * torch.randn(n, generator=g).double()).requires_grad_(True)
    color = (gt_color + 0.2 * torch.randn(n, 3, generator=g)).requires_grad_(True)
    weight = torch.rand(n, S, 1, generator=g).requires_grad_(True)
    ref = MO.mapper_loss(depth, color, weight, gt_depth, gt_color, use_weight, 0.2, stage_color)
    ref.backward()
    d2, c2, w2 = (t.detach().to(DEV).requires_grad_(True) for t in (depth, color, weight))
    got = M.mapper_loss(d2, c2, w2, gt_depth.to(DEV), gt_color.to(DEV), 0.2, use_weight, stage_color)
    (got * 1.0).backward()
    assert got.dtype == torch.float64
    assert abs(got.item() - ref.item()) <= 1e-6 * abs(ref.item())
    assert torch.equal(d2.grad.cpu(), depth.grad)
    if stage_color:
        assert torch.allclose(c2.grad.cpu(), color.grad, rtol=1e-6)
    else:
        assert c2.grad is None
    if use_weight:
        assert torch.equal(w2.grad.cpu(), weight.grad)
    else:
        assert w2.grad is None
