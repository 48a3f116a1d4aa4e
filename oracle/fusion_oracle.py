"""CPU oracle for the offline rows next to the hot path.  TEST INFRASTRUCTURE ONLY.

* ``integrate``: numpy f32 restatement of the reference's TSDF `integrate` CUDA kernel (src/fusion.py:69-142), operation
  by operation (no fused multiply-adds).  Pinned against the reference's own CPU-mode ``TSDFVolume.integrate``
  (src/fusion.py:236-290) by oracle/make_golden_fusion.py -> tests/golden/fusion_small.pt; the two upstream modes differ
  in rounding (matrix inverse vs transposed rotation, np.round vs roundf), so that pin is up to boundary voxels.
* ``point_masks``: the torch ops of Mesher.point_masks (src/utils/Mesher.py:58-204, depth_test False), pinned against the
  reference's method itself (same generator script).
"""
from __future__ import annotations

import numpy as np
import torch

f32 = np.float32


def _roundf(x):
    """C roundf (half away from zero) of an f32 array, exactly."""
    x64 = x.astype(np.float64)
    return np.where(x64 >= 0, np.floor(x64 + 0.5), np.ceil(x64 - 0.5)).astype(np.float32)


def integrate(tsdf, weight, color, origin, voxel, trunc, depth, color_folded, K, pose, obs_weight):
    """In place on the [X,Y,Z] f32 volumes.  color_folded [H,W] f32 = floor(b*65536 + g*256 + r)."""
    dx, dy, dz = tsdf.shape
    H, W = depth.shape
    voxel, trunc, obs = f32(voxel), f32(trunc), f32(obs_weight)
    K = np.asarray(K, dtype=np.float64).reshape(-1).astype(np.float32)
    P = np.asarray(pose, dtype=np.float64).reshape(-1).astype(np.float32)
    origin = np.asarray(origin, dtype=np.float32)
    idx = np.arange(dx * dy * dz, dtype=np.int64)
    idf = idx.astype(np.int32).astype(np.float32)
    vx = np.floor(idf / f32(dy * dz))                                                    # fusion.py:89
    rem = (idx - vx.astype(np.int64) * dy * dz)
    vy = np.floor(rem.astype(np.int32).astype(np.float32) / f32(dz))                     # :90
    vz = (rem - vy.astype(np.int64) * dz).astype(np.int32).astype(np.float32)            # :91
    px, py, pz = origin[0] + vx * voxel, origin[1] + vy * voxel, origin[2] + vz * voxel
    tx, ty, tz = px - P[3], py - P[7], pz - P[11]
    cx = (P[0] * tx + P[4] * ty) + P[8] * tz
    cy = (P[1] * tx + P[5] * ty) + P[9] * tz
    cz = (P[2] * tx + P[6] * ty) + P[10] * tz
    with np.errstate(divide="ignore", invalid="ignore"):
        fx_ = _roundf(K[0] * (cx / cz) + K[2])
        fy_ = _roundf(K[4] * (cy / cz) + K[5])
    ok = np.isfinite(fx_) & np.isfinite(fy_)
    pixx = np.where(ok, fx_, -1).astype(np.int64)
    pixy = np.where(ok, fy_, -1).astype(np.int64)
    valid = ok & (pixx >= 0) & (pixx < W) & (pixy >= 0) & (pixy < H) & ~(cz < 0)
    dv = np.zeros_like(cz)
    dv[valid] = depth[pixy[valid], pixx[valid]]
    valid &= dv != 0
    diff = dv - cz
    valid &= ~(diff < -trunc)
    dist = np.minimum(f32(1.0), diff / trunc)
    v = np.nonzero(valid)[0]
    t, w, c = tsdf.reshape(-1), weight.reshape(-1), color.reshape(-1)
    w_old = w[v].copy()
    w_new = w_old + obs
    w[v] = w_new
    t[v] = (t[v] * w_old + obs * dist[v]) / w_new
    oc = c[v]
    ob = np.floor(oc / f32(65536))
    og = np.floor((oc - ob * f32(65536)) / f32(256))
    orr = oc - ob * f32(65536) - og * f32(256)
    nc = color_folded[pixy[v], pixx[v]].astype(np.float32)
    nb = np.floor(nc / f32(65536))
    ng = np.floor((nc - nb * f32(65536)) / f32(256))
    nr = nc - nb * f32(65536) - ng * f32(256)
    nb = np.minimum(_roundf((ob * w_old + obs * nb) / w_new), f32(255))
    ng = np.minimum(_roundf((og * w_old + obs * ng) / w_new), f32(255))
    nr = np.minimum(_roundf((orr * w_old + obs * nr) / w_new), f32(255))
    c[v] = nb * f32(65536) + ng * f32(256) + nr


def fold_color(color_im):
    c = np.asarray(color_im).astype(np.float32)
    return np.floor(c[..., 2] * f32(65536) + c[..., 1] * f32(256) + c[..., 0])           # fusion.py:207-208


def point_masks(points, w2c_list, max_depths, H, W, fx, fy, cx, cy):
    """seen / forecast / unseen bool arrays -- Mesher.py:58-204 with depth_test False (torch ops as upstream)."""
    points = torch.as_tensor(points).float()
    seen = torch.zeros(points.shape[0], dtype=torch.bool)
    fore = torch.zeros(points.shape[0], dtype=torch.bool)
    for w2c, md in zip(w2c_list, max_depths):
        w2c = torch.as_tensor(w2c).float()
        homo = torch.cat([points, torch.ones_like(points[:, :1])], dim=1).reshape(-1, 4, 1)
        cam = (w2c @ homo)[:, :3]
        K = torch.from_numpy(np.array([[fx, .0, cx], [.0, fy, cy], [.0, .0, 1.0]]).reshape(3, 3))
        cam[:, 0] *= -1
        uv = K.float() @ cam.float()
        z = uv[:, -1:] + 1e-8
        uv = (uv[:, :2] / z).float()
        def inside(edge):
            return ((uv[:, 0] < W - edge) & (uv[:, 0] > edge) & (uv[:, 1] < H - edge) & (uv[:, 1] > edge) & (z[:, :, 0] < 0)).reshape(-1)
        s, f = inside(0), inside(-1000)
        proj = -cam[:, 2].reshape(-1)
        s = s & (proj < md)
        f = f & (proj < md)
        seen |= s
        fore |= f
    fore &= ~seen
    return seen.numpy(), fore.numpy(), (~(seen | fore)).numpy()
