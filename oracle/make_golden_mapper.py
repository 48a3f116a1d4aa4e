"""Generate tests/golden/mapper_rows.pt from the REFERENCE's own functions (run in the authoring container only).

  python oracle/make_golden_mapper.py

* frustum masks: `Mapper.get_mask_from_c2w` (src/Mapper.py:90-158) called unbound on a stand-in `self`, with the
  container's cv2 (4.13) doing the remap.  src/Mapper.py imports colorama, the dataset readers and the Visualizer
  (matplotlib/open3d), none of which the function needs: they are stubbed in sys.modules for the import only.
* ray batches: `get_samples` (src/common.py:127-136) after torch.manual_seed, so the torch.randint draw of
  select_uv is reproducible; the fixture stores the drawn indices.
The optimiser flow / filter / loss are inline in Mapper.optimize_map; their fixtures come from the restatement in
oracle/mapper_oracle.py (torch.optim.Adam underneath) and pin the arithmetic across torch versions.
"""
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

for name in ("colorama", "src.utils.datasets", "src.utils.Visualizer"):
    m = types.ModuleType(name)
    m.Fore = m.Style = SimpleNamespace()
    m.get_dataset = None
    m.Visualizer = None
    sys.modules[name] = m
from src.Mapper import Mapper as RefMapper          # noqa: E402
from src.common import get_samples as ref_get_samples          # noqa: E402

from oracle import mapper_oracle as MO          # noqa: E402


def make_pose(bound, seed):
    """camera-to-world [4,4] f32: random rotation about y plus a small tilt, position near the box centre"""
    rs = np.random.RandomState(seed)
    a, b = rs.uniform(0, 2 * np.pi), rs.uniform(-0.2, 0.2)
    Ry = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
    c2w = np.eye(4)
    c2w[:3, :3] = Ry @ Rx
    lo, hi = bound[:, 0].numpy(), bound[:, 1].numpy()
    c2w[:3, 3] = (lo + hi) / 2 + rs.uniform(-0.15, 0.15, 3) * (hi - lo)
    return torch.from_numpy(c2w).float()


def main():
    torch.manual_seed(7)
    np.random.seed(7)
    H, W, fx, fy, cx, cy = 68, 120, 60.0, 60.0, 59.5, 33.5
    bound = torch.tensor([[-1.5, 2.34], [-1.2, 1.68], [-1.0, 1.24]], dtype=torch.float64)
    out = {"H": H, "W": W, "fx": fx, "fy": fy, "cx": cx, "cy": cy, "bound": bound, "masks": [], "batches": []}

    # ---- frustum masks: two poses x two grid shapes
    for seed, val_shape in ((0, (8, 10, 13)), (1, (15, 19, 25)), (2, (8, 10, 13))):
        g = torch.Generator().manual_seed(seed)
        c2w = make_pose(bound, seed)
        depth = (1.0 + 2.0 * torch.rand(H, W, generator=g)).float()
        depth[torch.rand(H, W, generator=g) < 0.1] = 0.0          # holes -> filled with the maximum (Mapper.py:137-138)
        if seed == 2:
            depth[:] = 0.3                                        # most vertices fail the depth test
        stand_in = SimpleNamespace(H=H, W=W, fx=fx, fy=fy, cx=cx, cy=cy, bound=bound)
        mask = RefMapper.get_mask_from_c2w(stand_in, c2w, "grid_low", val_shape, depth.numpy())
        out["masks"].append({"c2w": c2w, "depth": depth, "val_shape": val_shape, "mask": torch.from_numpy(mask.copy())})

    # ---- ray batches: whole image (Mapper) and edge-cropped window (Tracker)
    for seed, (H0, H1, W0, W1), n in ((3, (0, H, 0, W), 300), (4, (6, H - 6, 10, W - 10), 200)):
        g = torch.Generator().manual_seed(seed)
        c2w = make_pose(bound, seed)
        depth = (0.5 + 3.0 * torch.rand(H, W, generator=g)).float()
        color = torch.rand(H, W, 3, generator=g).float()
        torch.manual_seed(100 + seed)
        idx = torch.randint((H1 - H0) * (W1 - W0), (n,))          # the draw select_uv will repeat
        torch.manual_seed(100 + seed)
        ro, rd, gd, gc = ref_get_samples(H0, H1, W0, W1, n, H, W, fx, fy, cx, cy, c2w, depth, color, "cpu")
        out["batches"].append({"c2w": c2w, "depth": depth, "color": color, "window": (H0, H1, W0, W1), "indices": idx,
                               "rays_o": ro.clone(), "rays_d": rd.clone(), "gt_depth": gd.clone(), "gt_color": gc.clone()})

    # ---- masked Adam: restated flow (torch.optim.Adam), 6 iterations with the stage-dependent lr pattern incl. lr = 0
    g = torch.Generator().manual_seed(11)
    val = (0.01 * torch.randn(1, 32, 6, 7, 9, generator=g)).float()
    vmask = torch.rand(6, 7, 9, generator=g) < 0.4
    grads = [(0.05 * torch.randn(1, 32, 6, 7, 9, generator=g)).float() for _ in range(6)]
    grads[2] = None                                               # the grid took no part in that iteration
    lrs = [0.1, 0.1, 0.005, 0.0, 0.005, 0.02]
    final = MO.masked_adam_run(val, vmask, lambda v, it: grads[it], lrs)
    final_all = MO.masked_adam_run(val, None, lambda v, it: grads[it], lrs)
    out["adam"] = {"val": val, "mask": vmask, "grads": grads, "lrs": lrs, "final_masked": final, "final_all": final_all}

    path = os.path.join(ROOT, "tests", "golden", "mapper_rows.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
