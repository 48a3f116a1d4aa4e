"""smoke()'s comparison with per-tensor errors (diagnostics)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import attentive_dfprior_b200 as A
from attentive_dfprior_b200 import synth
from attentive_dfprior_b200.config import default_cfg, slam_namespace
from oracle import dfprior_oracle as O
dev = "cuda:0"
nrays = int(os.environ.get("NRAYS", "64"))
torch.manual_seed(0)
scene = synth.make_scene("small", seed=0, high_std=1e-2)
model = A.DF()
with torch.no_grad():
    for dec in (model.low_decoder, model.high_decoder):
        dec.output_linear.weight.mul_(0.05)
sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
rays_o, rays_d, gt, gt_color = synth.make_rays(scene, nrays, seed=1)
params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
grids_cpu = {k: v.clone().requires_grad_(True) for k, v in scene.grids.items()}
d_ref, u_ref, c_ref, w_ref = O.render_batch_ray(params, grids_cpu, rays_d, rays_o, scene.tsdf_volume, scene.tsdf_bnds, scene.bound, "color", gt)
O.mapper_loss(d_ref, c_ref, w_ref, gt, gt_color, "color").backward()
model = model.to(dev); sc = scene.to(dev)
grids = {k: v.contiguous(memory_format=torch.channels_last_3d).requires_grad_(True) for k, v in sc.grids.items()}
renderer = A.Renderer(default_cfg(), None, slam_namespace(scene.bound, sc.tsdf_bnds))
d, u, c, w = renderer.render_batch_ray(grids, model, rays_d.to(dev), rays_o.to(dev), dev, sc.tsdf_volume, sc.tsdf_bnds, "color", gt_depth=gt.to(dev))
gtd = gt.to(dev); m = gtd > 0
(torch.abs(gtd[m] - d[m]).sum() + 0.2 * torch.abs(gt_color.to(dev) - c).sum()).backward()
torch.cuda.synchronize()
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)
named = dict(model.named_parameters())
errs = sorted(((rel(named[k].grad, v.grad), k, float(v.grad.abs().max())) for k, v in params.items() if v.grad is not None and float(v.grad.abs().max()) > 0), reverse=True)
print(os.environ.get("TAG", ""), "nrays", nrays, "grids", {k: f"{rel(grids[k].grad, grids_cpu[k].grad):.1e}" for k in grids})
for e, k, mx in errs[:6]:
    print(f"   {e:.2e}  {k}  (max |grad| {mx:.2e})")
