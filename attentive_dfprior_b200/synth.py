"""Seeded synthetic Replica-shaped inputs for the hot path.

Shared by ``bench.py`` and the tests so that the CUDA path and the oracle see
identical inputs.  Nothing here depends on the oracle or on the reference
checkout.  Shapes and value rules follow the reference's host glue:

* scene bound rounding          -- src/DF_Prior.py:185-190
* feature-grid shapes and init  -- src/DF_Prior.py:221-266 (normal_(0, 0.01) low/colour, 1e-4 high)
* TSDF layout                   -- get_tsdf.py:95-97 (stored [X,Y,Z], viewed as non-contiguous
                                   [1,1,Z,Y,X]), voxel 4/256 m and bounds src/fusion.py:38-52
* camera intrinsics             -- configs/df_prior.yaml:84-90
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np
import torch

ROOM0_BOUND = [[-2.9, 8.9], [-3.2, 5.5], [-3.5, 3.3]]          # configs/Replica/room0.yaml:3
SCANNET0000_BOUND = [[-2.0, 11.0], [-2.0, 11.5], [-2.0, 5.5]]  # configs/ScanNet/scene0000.yaml:3
SMALL_BOUND = [[-1.0, 1.1], [-0.9, 1.2], [-0.8, 0.7]]          # test-sized scene (fixtures stay small)

GRID_LEN = {"low": 0.32, "high": 0.16, "color": 0.16}          # configs/df_prior.yaml:6-10
CAM = dict(H=680, W=1200, fx=600.0, fy=600.0, cx=599.5, cy=339.5)   # configs/df_prior.yaml:84-90


def round_bound(bound, divisible: float = 0.32) -> torch.Tensor:
    """src/DF_Prior.py:185-190 -- enlarge the upper edge to a multiple of ``divisible``."""
    b = torch.from_numpy(np.array(bound, dtype=np.float64))
    b[:, 1] = (((b[:, 1] - b[:, 0]) / divisible).int() + 1) * divisible + b[:, 0]
    return b


def grid_shape(bound: torch.Tensor, grid_len: float):
    """[D,H,W] = vertex counts along z,y,x -- src/DF_Prior.py:241-246."""
    s = list(map(int, ((bound[:, 1] - bound[:, 0]) / grid_len).tolist()))
    s[0], s[2] = s[2], s[0]
    return s


@dataclass
class Scene:
    bound: torch.Tensor                      # [3,2] f64, CPU (as the reference keeps it)
    grids: Dict[str, torch.Tensor]           # grid_low/high/color, logical [1,32,D,H,W] f32
    tsdf_volume: torch.Tensor                # logical [1,1,Z,Y,X] f32, non-contiguous (x-major memory)
    tsdf_bnds: torch.Tensor                  # [3,2] f64, on the compute device
    room_lo: torch.Tensor                    # [3] f64 analytic room box (for depth synthesis)
    room_hi: torch.Tensor
    voxel: float
    cam: dict = field(default_factory=lambda: dict(CAM))

    def to(self, device) -> "Scene":
        return Scene(self.bound, {k: v.to(device) for k, v in self.grids.items()},
                     self.tsdf_volume.to(device), self.tsdf_bnds.to(device),
                     self.room_lo, self.room_hi, self.voxel, dict(self.cam))


def make_scene(kind: str = "room0", seed: int = 0, device="cpu", c_dim: int = 32,
               high_std: float = 1e-4, voxel: Optional[float] = None,
               channels_last: bool = False) -> Scene:
    """Random-init grids + analytic box-room TSDF for one of the named bounds."""
    raw = {"room0": ROOM0_BOUND, "scannet0000": SCANNET0000_BOUND, "small": SMALL_BOUND}[kind]
    bound = round_bound(raw)
    if voxel is None:
        voxel = 4.0 / 256 if kind != "small" else 0.05
    g = torch.Generator().manual_seed(seed)
    grids = {}
    for name, std in (("low", 0.01), ("high", high_std), ("color", 0.01)):
        D, H, W = grid_shape(bound, GRID_LEN[name])
        v = torch.randn(1, c_dim, D, H, W, generator=g) * std
        v = v.to(device)
        if channels_last:
            v = v.contiguous(memory_format=torch.channels_last_3d)
        grids["grid_" + name] = v

    # TSDF over the (rounded) scene box: dims = ceil(L / voxel) per axis, upper bound = lower + dims * voxel
    # (src/fusion.py:38-43); room0 at 4/256 m gives [X,Y,Z] = [758,574,451]
    b = bound.numpy().astype(np.float64).copy()
    dims = np.ceil(np.round((b[:, 1] - b[:, 0]) / voxel, 9)).astype(int)
    tb = b.copy()
    tb[:, 1] = tb[:, 0] + dims * voxel
    braw = np.array(raw, dtype=np.float64)                            # the analytic room sits in the unrounded box
    centre = (braw[:, 0] + braw[:, 1]) / 2
    half = (braw[:, 1] - braw[:, 0]) / 2 * 0.8
    room_lo, room_hi = centre - half, centre + half
    ax = [torch.arange(int(dims[k]), dtype=torch.float32, device=device) * voxel + float(tb[k, 0])
          for k in range(3)]
    # signed distance to the nearest wall: > 0 inside the room (free space), < 0 behind a wall
    sx = torch.minimum(ax[0] - float(room_lo[0]), float(room_hi[0]) - ax[0])
    sy = torch.minimum(ax[1] - float(room_lo[1]), float(room_hi[1]) - ax[1])
    sz = torch.minimum(ax[2] - float(room_lo[2]), float(room_hi[2]) - ax[2])
    sdf = torch.minimum(torch.minimum(sx[:, None, None], sy[None, :, None]), sz[None, None, :])   # [X,Y,Z]
    trunc = 5 * voxel                                                   # src/fusion.py:38
    tsdf = torch.clamp(sdf / trunc, max=1.0)
    tsdf = torch.where(sdf < -trunc, torch.full_like(tsdf, -1.0), tsdf)
    X, Y, Z = tsdf.shape
    tsdf_volume = tsdf.reshape(1, 1, X, Y, Z).permute(0, 1, 4, 3, 2)    # get_tsdf.py:95-97
    return Scene(bound, grids, tsdf_volume, torch.from_numpy(tb).to(device),
                 torch.from_numpy(room_lo), torch.from_numpy(room_hi), float(voxel))


def make_pose(scene: Scene, seed: int = 0, jitter: float = 0.15) -> torch.Tensor:
    """Camera-to-world [4,4] f32 near the room centre with a small random rotation."""
    rs = np.random.RandomState(seed)
    ang = rs.uniform(-jitter, jitter, 3) + np.array([0.0, rs.uniform(0, 2 * math.pi), 0.0])
    cx, cy, cz = np.cos(ang)
    sx, sy, sz = np.sin(ang)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    c2w = np.eye(4)
    c2w[:3, :3] = Rz @ Ry @ Rx
    lo, hi = scene.room_lo.numpy(), scene.room_hi.numpy()
    c2w[:3, 3] = (lo + hi) / 2 + rs.uniform(-0.1, 0.1, 3) * (hi - lo)
    return torch.from_numpy(c2w).float()


def depth_along(scene: Scene, rays_o: torch.Tensor, rays_d: torch.Tensor) -> torch.Tensor:
    """Analytic z-depth to the room box (rays_d has camera-z = -1, so t is z-depth)."""
    o, d = rays_o.double(), rays_d.double()
    lo = scene.room_lo.to(o.device)
    hi = scene.room_hi.to(o.device)
    t = torch.maximum((lo - o) / d, (hi - o) / d).min(dim=-1)[0]
    return t.float()


def make_rays(scene: Scene, n: int, seed: int = 0, device="cpu", zero_frac: float = 0.1,
              c2w: Optional[torch.Tensor] = None, edge: int = 0):
    """n random pixels of one synthetic RGB-D frame -> rays_o, rays_d, gt_depth, gt_color (all f32).

    Ray construction follows src/common.py:76-91; ``zero_frac`` of the depths are zeroed to
    exercise the depth==0 sampling branch (src/utils/Renderer.py:194-201)."""
    cam = scene.cam
    g = torch.Generator().manual_seed(seed + 7919)
    if c2w is None:
        c2w = make_pose(scene, seed)
    c2w = c2w.to(device)
    i = torch.randint(edge, cam["W"] - edge, (n,), generator=g).float().to(device)
    j = torch.randint(edge, cam["H"] - edge, (n,), generator=g).float().to(device)
    dirs = torch.stack([(i - cam["cx"]) / cam["fx"], -(j - cam["cy"]) / cam["fy"], -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs.reshape(-1, 1, 3) * c2w[:3, :3], -1)
    rays_o = c2w[:3, 3].expand(rays_d.shape).contiguous()
    depth = depth_along(scene, rays_o, rays_d)
    zero = torch.rand(n, generator=g) < zero_frac
    depth = torch.where(zero.to(device), torch.zeros_like(depth), depth)
    color = torch.rand(n, 3, generator=g).to(device)
    return rays_o.float().contiguous(), rays_d.float().contiguous(), depth.contiguous(), color


def make_depth_image(scene: Scene, c2w: torch.Tensor, device="cpu", zero_frac: float = 0.1, seed: int = 0):
    """Full-frame synthetic depth [H,W] f32 for render_img (config 4)."""
    cam = scene.cam
    H, W = cam["H"], cam["W"]
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="ij")
    i, j = i.t().to(device), j.t().to(device)
    dirs = torch.stack([(i - cam["cx"]) / cam["fx"], -(j - cam["cy"]) / cam["fy"], -torch.ones_like(i)], -1)
    c2w = c2w.to(device)
    rays_d = torch.sum(dirs.reshape(H, W, 1, 3) * c2w[:3, :3], -1).reshape(-1, 3)
    rays_o = c2w[:3, 3].expand(rays_d.shape)
    depth = depth_along(scene, rays_o, rays_d)
    g = torch.Generator().manual_seed(seed + 104729)
    zero = (torch.rand(H * W, generator=g) < zero_frac).to(device)
    return torch.where(zero, torch.zeros_like(depth), depth).reshape(H, W)
