"""CPU: the offline-row oracles (oracle/fusion_oracle.py) against the reference's own outputs (tests/golden/fusion_small.pt,
oracle/make_golden_fusion.py): TSDF integrate (src/fusion.py) and Mesher.point_masks (src/utils/Mesher.py:58-204)."""
import os

import numpy as np
import torch

from conftest import GOLDEN_DIR
from oracle import fusion_oracle as FO


def load():
    return torch.load(os.path.join(GOLDEN_DIR, "fusion_small.pt"), weights_only=False)


def run_oracle(fx):
    i = fx["inputs"]
    bnds = i["vol_bnds"].numpy().copy()
    voxel = i["voxel_size"]
    dim = np.ceil((bnds[:, 1] - bnds[:, 0]) / voxel).astype(int)
    bnds[:, 1] = bnds[:, 0] + dim * voxel
    origin = bnds[:, 0].astype(np.float32)
    tsdf = -np.ones(dim, dtype=np.float32)
    weight = np.zeros(dim, dtype=np.float32)
    color = np.zeros(dim, dtype=np.float32)
    for c, d, p in i["frames"]:
        FO.integrate(tsdf, weight, color, origin, voxel, 5 * voxel, d.numpy(), FO.fold_color(c.numpy()), i["K"].numpy(), p.numpy(), 1.0)
    return tsdf, weight, color, bnds, dim


def test_integrate_oracle_matches_reference_cpu_mode():
    """The oracle restates upstream's CUDA kernel; the fixture is upstream's CPU mode.  The two differ in how they round
    (inverse-matrix transform in f64 vs transposed rotation in f32, np.round vs roundf), which moves a voxel across a
    pixel or truncation boundary now and then: the volumes must agree on all but a small fraction of voxels."""
    fx = load()
    o = fx["outputs"]
    tsdf, weight, color, bnds, dim = run_oracle(fx)
    assert np.array_equal(dim, o["vol_dim"].numpy()) and np.allclose(bnds, o["vol_bnds"].numpy())
    n = tsdf.size
    same_w = weight == o["weight"].numpy()
    assert same_w.mean() >= 0.995, same_w.mean()
    close = np.abs(tsdf - o["tsdf"].numpy()) <= 1e-4
    assert (close | ~same_w).all(), "where the observation counts agree the TSDF values must agree"
    assert (close.sum() / n) >= 0.995
    # colour channels: upstream's kernel rounds half away from zero (roundf), its CPU mode half to even (np.round): a
    # running average that lands on .5 differs by one level between the two modes, never more
    def channels(c):
        b = np.floor(c / 65536)
        g = np.floor((c - b * 65536) / 256)
        return np.stack([b, g, c - b * 65536 - g * 256], -1)
    dc = np.abs(channels(color.astype(np.float64)) - channels(o["color"].numpy().astype(np.float64)))
    assert dc[same_w].max() <= 1
    assert (dc[same_w].max(-1) == 0).mean() >= 0.98


def test_point_masks_oracle_matches_reference():
    fx = load()["point_masks"]
    cam = fx["cam"]
    w2c = [np.linalg.inv(k["est_c2w"].numpy()) for k in fx["keyframes"]]
    md = [float(torch.max(k["depth"])) * 1.1 for k in fx["keyframes"]]
    seen, fore, unseen = FO.point_masks(fx["points"], w2c, md, cam["H"], cam["W"], cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    assert np.array_equal(seen, fx["seen"].numpy()) and np.array_equal(fore, fx["forecast"].numpy()) and np.array_equal(unseen, fx["unseen"].numpy())
    w2c = [np.linalg.inv(p.numpy()) for p in fx["est"]]
    seen, fore, unseen = FO.point_masks(fx["points"], w2c, [float("inf")] * 4, cam["H"], cam["W"], cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    assert np.array_equal(seen, fx["seen_all"].numpy()) and np.array_equal(fore, fx["forecast_all"].numpy())
