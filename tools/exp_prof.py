"""Experiment build only (tools/exp_run.sh): print the phase timelines recorded by the instrumented fwd / K1 kernels."""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
import bench  # noqa
from attentive_dfprior_b200 import _lib as L
lib = L.lib()
import subprocess, json, os
# one bench process fills the counters; run the step loop in-process instead
import attentive_dfprior_b200 as A
from attentive_dfprior_b200 import synth, mapper_ops as M
from attentive_dfprior_b200.config import default_cfg, slam_namespace
dev = torch.device("cuda", 0)
scene = synth.make_scene("room0", seed=0, device=dev, high_std=1e-2, channels_last=True)
model = A.DF().to(dev)
renderer = A.Renderer(default_cfg(), None, slam_namespace(scene.bound, scene.tsdf_bnds))
grids = {k: v.requires_grad_(True) for k, v in scene.grids.items()}
ro, rd, gt, gc = synth.make_rays(scene, 65536, seed=100, device=dev)
def step():
    d, u, c, w = renderer.render_batch_ray(grids, model, rd, ro, dev, scene.tsdf_volume, scene.tsdf_bnds, "color", gt_depth=gt)
    loss = M.mapper_loss(d, c, w, gt, gc, 0.2, False, True)
    loss.backward()
for _ in range(3): step()
torch.cuda.synchronize()
for name in ("fwd", "dh"):
    getattr(lib, f"dfp_exp_clear_prof_{name}")()
step(); torch.cuda.synchronize()
for name in ("fwd", "dh"):
    buf = (C.c_ulonglong * 32)()
    getattr(lib, f"dfp_exp_read_prof_{name}")(buf)
    n = max(1, buf[31])
    print(name, "tiles", buf[31], "cycles/tile by phase:", {i: round(buf[i] / n) for i in range(31) if buf[i]}, "total", round(sum(buf[:31]) / n))
