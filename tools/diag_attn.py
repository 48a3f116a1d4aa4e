"""Diagnostic (GPU box): tensor-core vs SIMT path on the non-leaf grid flow case; where do the gradients differ?"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch
from oracle import dfprior_oracle as O
from test_gpu_parity import DEV, build, rel_err
from test_oracle_extra import load_extra
from attentive_dfprior_b200 import functional as Fn

fx = load_extra("extra_nodepth_color")
inp = fx["inputs"]
gt = inp["gt_depth"]
sd = {k: v.clone() for k, v in inp["state_dict"].items()}
gcpu = {k: v.clone().requires_grad_(True) for k, v in inp["grids"].items()}
pcpu = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
(d_ref, u_ref, c_ref, w_ref), aux = O.render_batch_ray(pcpu, gcpu, inp["rays_d"], inp["rays_o"], inp["tsdf_volume"], inp["tsdf_bnds"], inp["bound"], "color", gt, return_aux=True)
O.mapper_loss(d_ref, c_ref, w_ref, gt, inp["gt_color"], "color", attn_term=True).backward()
out = {}
for simt in (True, False):
    model, renderer, grids, tsdf, tb = build(inp, channels_last=True)
    g1 = {k: v.clone().requires_grad_(True) for k, v in grids.items()}
    (d, u, c, w), ax = Fn.render_batch_ray(g1, model, inp["rays_d"].to(DEV), inp["rays_o"].to(DEV), tsdf, tb, inp["bound"], "color", gt_depth=gt.to(DEV), simt=simt, want_aux=True)
    g = gt.to(DEV); m = g > 0
    (torch.abs(g[m] - d[m]).sum() + torch.abs(w - 1).sum() + 0.2 * torch.abs(inp["gt_color"].to(DEV) - c).sum()).backward()
    named = dict(model.named_parameters())
    print("simt" if simt else "tc  ", "fwd: depth %.1e w %.1e raw %.1e" % (rel_err(d, d_ref), rel_err(w, w_ref), rel_err(ax["raw"], aux["raw"])),
          "| grads:", {k: "%.1e" % rel_err(g1[k].grad, gcpu[k].grad) for k in g1},
          "| worst mlp param:", max(("%.1e" % rel_err(named[k].grad, v.grad), k) for k, v in pcpu.items() if k.startswith("mlp.") and v.grad is not None and float(v.grad.abs().max()) > 0), flush=True)
    out[simt] = (w.detach().cpu(), ax["raw"].cpu(), {k: v.grad.cpu() for k, v in g1.items()}, {k: named[k].grad.cpu() for k in named})
wt, ws = out[False][0], out[True][0]
print("w: tc vs simt max abs diff", float((wt - ws).abs().max()), " raw occ tc vs simt", float((out[False][1][..., 3] - out[True][1][..., 3]).abs().max()))
gl_t, gl_s = out[False][2]["grid_low"], out[True][2]["grid_low"]
dd = (gl_t - gl_s).abs(); mx = float(gl_s.abs().max())
bad = (dd > 2e-4 * mx).nonzero()
print("grid_low tc vs simt: max rel", float(dd.max()) / mx, "elements > 2e-4:", len(bad), "voxels:", sorted(set((int(b[2]), int(b[3]), int(b[4])) for b in bad))[:10])
for k in ("mlp.pts_linears.0.weight", "mlp.pts_linears.0.bias", "mlp.pts_linears.1.weight", "mlp.pts_linears.2.weight", "mlp.pts_linears.3.weight", "mlp.pts_linears.3.bias", "mlp.output_linear.weight", "mlp.output_linear.bias"):
    print(k, "tc vs simt %.1e" % rel_err(out[False][3][k], out[True][3][k]))
