"""Generate tests/golden/fusion_small.pt: the reference's own TSDF fusion (CPU mode of src/fusion.py, imported with numba /
skimage / pycuda stubbed: @njit becomes the identity, prange = range) and the reference's own Mesher.point_masks on
seeded synthetic RGB-D frames.  TEST INFRASTRUCTURE ONLY (authoring container; /root/reference mounted).

Usage:  python oracle/make_golden_fusion.py
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))
import make_golden as MG  # noqa: E402
import make_golden_extra as MGE  # noqa: E402


def load_reference_fusion():
    nb = types.ModuleType("numba")
    nb.njit = lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda f: f))
    nb.prange = range
    sys.modules.setdefault("numba", nb)
    sk = types.ModuleType("skimage"); sk.measure = types.ModuleType("skimage.measure")
    sys.modules.setdefault("skimage", sk); sys.modules.setdefault("skimage.measure", sk.measure)
    sys.path.insert(0, MG.REF)
    cwd = os.getcwd()
    os.chdir(MG.REF)
    try:
        import src.fusion as F
    finally:
        os.chdir(cwd)
    return F


def synth_frames(n, H, W, seed):
    """Camera poses looking into a box room from its centre region + analytic depth + random colour."""
    rs = np.random.RandomState(seed)
    lo, hi = np.array([-1.0, -0.9, -0.8]), np.array([1.1, 1.2, 0.7])
    fx = fy = 40.0; cx, cy = (W - 1) / 2, (H - 1) / 2
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    frames = []
    for k in range(n):
        ang = rs.uniform(-0.4, 0.4, 3) + np.array([0, rs.uniform(0, 2 * np.pi), 0])
        cxa, cya, cza = np.cos(ang); sxa, sya, sza = np.sin(ang)
        R = (np.array([[cza, -sza, 0], [sza, cza, 0], [0, 0, 1]]) @ np.array([[cya, 0, sya], [0, 1, 0], [-sya, 0, cya]])
             @ np.array([[1, 0, 0], [0, cxa, -sxa], [0, sxa, cxa]]))
        pose = np.eye(4); pose[:3, :3] = R; pose[:3, 3] = (lo + hi) / 2 + rs.uniform(-0.2, 0.2, 3)
        # z-depth of the box walls along every pixel ray (camera looks along +z, as fusion.py's pinhole model)
        j, i = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
        d = np.stack([(i - cx) / fx, (j - cy) / fy, np.ones_like(i, dtype=float)], -1) @ R.T
        o = pose[:3, 3]
        with np.errstate(divide="ignore"):
            t = np.maximum((lo - o) / d, (hi - o) / d).min(-1)
        depth = t.astype(np.float32)
        depth[rs.rand(H, W) < 0.1] = 0
        color = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
        frames.append((color, depth, pose))
    return K, frames, (lo, hi)


def main():
    F = load_reference_fusion()
    H, W = 36, 48
    K, frames, (lo, hi) = synth_frames(3, H, W, 0)
    bnds = np.stack([lo - 0.05, hi + 0.07], 1)
    vol = F.TSDFVolume(bnds.copy(), voxel_size=0.07, use_gpu=False)
    for color, depth, pose in frames:
        vol.integrate(color, depth, K, pose, obs_weight=1.)
    tsdf, col, out_bnds = vol.get_volume()
    fx = {"name": "fusion_small",
          "inputs": {"vol_bnds": torch.from_numpy(bnds), "voxel_size": 0.07, "K": torch.from_numpy(K), "H": H, "W": W,
                     "frames": [(torch.from_numpy(c), torch.from_numpy(d), torch.from_numpy(p)) for c, d, p in frames]},
          "outputs": {"tsdf": torch.from_numpy(tsdf.copy()), "color": torch.from_numpy(col.copy()),
                      "weight": torch.from_numpy(vol._weight_vol_cpu.copy()), "vol_bnds": torch.from_numpy(np.asarray(out_bnds).copy()),
                      "vol_dim": torch.from_numpy(vol._vol_dim.copy())}}
    # ---- Mesher.point_masks (the reference's method, unbound, on a namespace)
    RefMesher = MGE.load_reference_mesher()
    ref = MG.load_reference()
    cam = dict(H=60, W=80, fx=70.0, fy=70.0, cx=39.5, cy=29.5)
    me = SimpleNamespace(points_batch_size=3000, depth_test=False, **cam)
    rs = np.random.RandomState(5)
    kfs, est = [], []
    for k in range(4):
        c2w = np.eye(4); ang = rs.uniform(0, 2 * np.pi)
        c2w[:3, :3] = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
        c2w[:3, 3] = rs.uniform(-0.3, 0.3, 3)
        c2w_t = torch.from_numpy(c2w).float()
        kfs.append({"est_c2w": c2w_t, "depth": torch.from_numpy(rs.uniform(0.5, 2.5, (cam["H"], cam["W"])).astype(np.float32))})
        est.append(c2w_t)
    pts = torch.from_numpy(rs.uniform(-3, 3, (7000, 3)).astype(np.float32))
    seen, fore, unseen = RefMesher.point_masks(me, pts.clone(), kfs, torch.stack(est), 3, "cpu", get_mask_use_all_frames=False)
    seen_a, fore_a, unseen_a = RefMesher.point_masks(me, pts.clone(), kfs, torch.stack(est), 3, "cpu", get_mask_use_all_frames=True)
    fx["point_masks"] = {"cam": cam, "points": pts, "keyframes": kfs, "est": torch.stack(est),
                         "seen": torch.from_numpy(seen), "forecast": torch.from_numpy(fore), "unseen": torch.from_numpy(unseen),
                         "seen_all": torch.from_numpy(seen_a), "forecast_all": torch.from_numpy(fore_a), "unseen_all": torch.from_numpy(unseen_a)}
    path = os.path.join(REPO, "tests", "golden", "fusion_small.pt")
    torch.save(fx, path)
    print("fusion_small: dims", vol._vol_dim, "observed voxels", int((vol._weight_vol_cpu > 0).sum()), "band voxels",
          int(((tsdf > -1) & (tsdf < 1)).sum()), "| point masks: seen", int(seen.sum()), "forecast", int(fore.sum()), "unseen", int(unseen.sum()),
          "->", f"{os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
