"""Where does the host time of a 1000-ray mapping step go?  (run on the GPU box)"""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import attentive_dfprior_b200 as A
from attentive_dfprior_b200 import functional as F, synth, _lib as L
from attentive_dfprior_b200.config import default_cfg, slam_namespace

acc = collections.defaultdict(float)
def wrap(owner, name, label=None):
    fn = getattr(owner, name)
    raw = fn.__func__ if isinstance(fn, staticmethod) else fn
    def w(*a, **k):
        t = time.perf_counter()
        try:
            return raw(*a, **k)
        finally:
            acc[label or name] += time.perf_counter() - t
    setattr(owner, name, staticmethod(w) if isinstance(owner, type) and name in ("forward", "backward") else w)

wrap(F._RenderBatchRay, "forward", "fn.forward")
wrap(F._RenderBatchRay, "backward", "fn.backward")
for n in ("_param_grads", "_grads_struct", "_scene_struct", "_cfg_struct", "_grid_ndhwc"):
    wrap(F, n)

dev = torch.device("cuda:0")
scene = synth.make_scene("room0", seed=0, device=dev, high_std=1e-2, channels_last=True)
model = A.DF().to(dev)
renderer = A.Renderer(default_cfg(), None, slam_namespace(scene.bound, scene.tsdf_bnds))
grids = {k: v.requires_grad_(True) for k, v in scene.grids.items()}
params = list(model.parameters())
ro, rd, g, gc = synth.make_rays(scene, 1000, seed=7, device=dev)

def step(sync_loss=True):
    for t in params: t.grad = None
    for t in grids.values(): t.grad = None
    A.decoder.param_pack_of(model).key = None
    t0 = time.perf_counter()
    d, u, c, w = renderer.render_batch_ray(grids, model, rd, ro, dev, scene.tsdf_volume, scene.tsdf_bnds, "color", gt_depth=g)
    acc["render call"] += time.perf_counter() - t0
    t0 = time.perf_counter()
    if sync_loss:
        m = g > 0
        loss = torch.abs(g[m] - d[m]).sum() + 0.2 * torch.abs(gc - c).sum()
    else:
        loss = A.mapper_ops.mapper_loss(d, c, w, g, gc, 0.2, False, True)
    acc["loss"] += time.perf_counter() - t0
    t0 = time.perf_counter()
    loss.backward()
    acc["backward call"] += time.perf_counter() - t0

for mode in (True, False):
    for _ in range(10): step(mode)
    torch.cuda.synchronize(); acc.clear()
    n = 100
    t0 = time.perf_counter()
    for _ in range(n): step(mode)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    print(f"torch loss={mode}: wall {wall:.3f} ms/step")
    for k, v in sorted(acc.items(), key=lambda x: -x[1]):
        print(f"   {k:16s} {v / n * 1e3:.3f} ms")
