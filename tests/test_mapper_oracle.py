"""CPU: the Mapper-side oracle (oracle/mapper_oracle.py) against fixtures produced by the reference's own functions
(tests/golden/mapper_rows.pt, oracle/make_golden_mapper.py)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mapper_oracle as MO  # noqa: E402

GOLD = torch.load(os.path.join(ROOT, "tests", "golden", "mapper_rows.pt"))


@pytest.mark.parametrize("k", range(len(GOLD["masks"])))
def test_frustum_mask_matches_reference(k):
    g, m = GOLD, GOLD["masks"][k]
    mine = MO.get_mask_from_c2w(m["c2w"], g["bound"], m["val_shape"], m["depth"].numpy(), g["H"], g["W"], g["fx"], g["fy"], g["cx"], g["cy"])
    ref = m["mask"].numpy()
    assert mine.shape == ref.shape and mine.dtype == np.bool_
    assert 0 < ref.sum() < ref.size
    assert np.array_equal(mine, ref)                                  # bit-exact mask


def test_remap_restatement_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    rs = np.random.RandomState(0)
    img = rs.rand(37, 53).astype(np.float32)
    u = rs.uniform(-8, 60, 5000).astype(np.float32)
    v = rs.uniform(-8, 45, 5000).astype(np.float32)
    u[:6] = [0.0, 52.0, 51.984375, 1e9, -1e9, np.float32(52.015625)]
    v[:6] = [0.0, 36.0, 35.5, 3.0, 3.0, np.float32(1e12)]
    ref = cv2.remap(img, u, v, interpolation=cv2.INTER_LINEAR)[:, 0]
    assert np.array_equal(MO.cv_remap_linear(img, u, v), ref)


@pytest.mark.parametrize("k", range(len(GOLD["batches"])))
def test_get_samples_matches_reference(k):
    g, b = GOLD, GOLD["batches"][k]
    H0, H1, W0, W1 = b["window"]
    o, d, gd, gc = MO.get_samples(H0, H1, W0, W1, b["indices"], g["fx"], g["fy"], g["cx"], g["cy"], b["c2w"], b["depth"], b["color"])
    for mine, ref in ((o, b["rays_o"]), (d, b["rays_d"]), (gd, b["gt_depth"]), (gc, b["gt_color"])):
        assert torch.equal(mine, ref)


def test_box_filter_and_loss_basics():
    g, b = GOLD, GOLD["batches"][0]
    keep = MO.box_filter(b["rays_o"], b["rays_d"], b["gt_depth"], g["bound"])
    assert 0 < keep.sum() < keep.numel()
    # a ray is kept iff its depth point lies inside the box along every axis it moves on
    p = b["rays_o"].double() + b["rays_d"].double() * b["gt_depth"].double()[:, None]
    inside = ((p >= g["bound"][:, 0] - 1e-9) & (p <= g["bound"][:, 1] + 1e-9)).all(1)
    assert (inside == keep).float().mean() > 0.99
    n = keep.numel()
    depth = b["gt_depth"].double() + 0.1
    loss = MO.mapper_loss(depth, b["gt_color"] * 0.5, torch.full((n, 4, 1), 0.75), b["gt_depth"], b["gt_color"], True, 0.2, True)
    want = 0.1 * n + 0.25 * 4 * n + 0.2 * (0.5 * b["gt_color"]).sum().item()
    assert abs(loss.item() - want) < 1e-3 * want


def test_masked_adam_fixture():
    a = GOLD["adam"]
    for mask, key in ((a["mask"], "final_masked"), (None, "final_all")):
        out = MO.masked_adam_run(a["val"], mask, lambda v, it: a["grads"][it], a["lrs"])
        assert torch.allclose(out, a[key], rtol=1e-6, atol=1e-9)
    # voxels outside the mask never move
    outside = ~a["mask"][None, None].expand_as(a["val"])
    assert torch.equal(a["final_masked"][outside], a["val"][outside])
    assert not torch.equal(a["final_masked"], a["val"])
