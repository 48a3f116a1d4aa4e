"""How many samples of the bench batch receive an exactly-zero / negligible upstream gradient (transmittance underflow)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import attentive_dfprior_b200 as A
from attentive_dfprior_b200 import functional as F, synth
from attentive_dfprior_b200.config import default_cfg, slam_namespace
dev = torch.device("cuda:0")
scene = synth.make_scene("room0", seed=0, device=dev, high_std=1e-2, channels_last=True)
model = A.DF().to(dev)
ro, rd, g, gc = synth.make_rays(scene, 65536, seed=100, device=dev)
out, aux = F.render_batch_ray(scene.grids, model, rd, ro, scene.tsdf_volume, scene.tsdf_bnds, scene.bound, "color", gt_depth=g, want_aux=True)
raw = aux["raw"]
alpha = torch.sigmoid(10 * raw[..., 3])
T = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1 - alpha + 1e-10], 1), 1)[:, :-1]
w = alpha * T
for thr in (0.0, 1e-30, 1e-12, 1e-8, 1e-6, 1e-4):
    print(f"T <= {thr:g}: {(T <= thr).float().mean().item():.3f}   w <= {thr:g}: {(w <= thr).float().mean().item():.3f}")
print("alpha mean", alpha.mean().item(), "occ min/max", raw[..., 3].min().item(), raw[..., 3].max().item())
