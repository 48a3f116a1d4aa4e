"""Diagnostic (GPU box): map_color_s2 fixture, tensor-core vs SIMT vs reference gradients; which samples differ?"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch
from conftest import load_golden
from test_gpu_parity import DEV, build, rel_err
from attentive_dfprior_b200 import functional as Fn

fx = load_golden("map_color_s2")
inp, ref = fx["inputs"], fx["outputs"]
gt = inp["gt_depth"].to(DEV); gc = inp["gt_color"].to(DEV)
out = {}
for simt in (True, False):
    model, renderer, grids, tsdf, tb = build(inp, channels_last=True)
    g1 = {k: v.clone().requires_grad_(True) for k, v in grids.items()}
    (d, u, c, w), ax = Fn.render_batch_ray(g1, model, inp["rays_d"].to(DEV), inp["rays_o"].to(DEV), tsdf, tb, inp["bound"], "color", gt_depth=gt, simt=simt, want_aux=True)
    m = gt > 0
    loss = torch.abs(gt[m] - d[m]).sum() + torch.abs(w - torch.ones_like(w)).sum() + 0.2 * torch.abs(gc - c).sum()
    loss.backward()
    print("simt" if simt else "tc  ", {k: "%.1e" % rel_err(g1[k].grad, fx["grads"]["grids"][k]) for k in g1}, flush=True)
    out[simt] = (w.detach().cpu().reshape(-1), ax["raw"].cpu().reshape(-1, 4), {k: v.grad.cpu() for k, v in g1.items()}, ax["z"].cpu())
gt_, gs_ = out[False][2]["grid_high"], out[True][2]["grid_high"]
dd = (gt_ - gs_).abs(); mx = float(gs_.abs().max())
bad = (dd > 2e-4 * mx).nonzero()
vox = sorted(set((int(b[2]), int(b[3]), int(b[4])) for b in bad))
print("grid_high tc vs simt max rel", float(dd.max()) / mx, "bad elements", len(bad), "voxels", vox[:12])
# samples whose high-grid cell touches the bad voxels
z = out[False][3]
pts = (inp["rays_o"][:, None, :].double() + inp["rays_d"][:, None, :].double() * z[:, :, None]).reshape(-1, 3)
b = inp["bound"]; D, H, W = gs_.shape[2:]
xn = (pts - b[:, 0]) / (b[:, 1] - b[:, 0])
ix, iy, iz = xn[:, 0] * (W - 1), xn[:, 1] * (H - 1), xn[:, 2] * (D - 1)
wt, ws = out[False][0], out[True][0]
for (vz, vy, vx) in vox[:3]:
    near = ((ix.floor() - vx).abs() <= 1) & ((iy.floor() - vy).abs() <= 1) & ((iz.floor() - vz).abs() <= 1) & (ix.floor() <= vx) & (iy.floor() <= vy) & (iz.floor() <= vz)
    idx = near.nonzero().flatten().tolist()
    print("voxel", (vz, vy, vx), "samples:", idx[:16])
    for i in idx[:16]:
        print("    s", i, "w tc %.9f simt %.9f ref %.9f | occ tc %.7f simt %.7f" % (float(wt[i]), float(ws[i]), float(ref["weight"].reshape(-1)[i]), float(out[False][1][i, 3]), float(out[True][1][i, 3])))
