"""Generate tests/golden/*.pt by running the UNMODIFIED reference on seeded inputs.

TEST INFRASTRUCTURE ONLY.  Runs in the authoring container, where the reference
checkout is mounted read-only at /root/reference (it does not exist on the GPU
box, so nothing at test/bench time calls this).  The reference's hot-path
modules are imported as they are; the only accommodation is the CPU shim from
SURVEY.md section 8c (the reference hard-codes 'cuda:N' device strings at
src/conv_onet/models/decoder.py:241,312) which maps those strings to 'cpu'.

Usage:  python oracle/make_golden.py            # writes tests/golden/*.pt

Each fixture holds the inputs (bound, grids, TSDF, decoder state_dict, rays)
and the reference's outputs (z_vals, pts, raw, depth/uncertainty/colour/weight,
and gradients for the mapper / tracker losses).
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DFPRIOR_REF", "/root/reference")
sys.path.insert(0, REPO)


def _install_cpu_shim():
    orig_to = torch.Tensor.to

    def _to(self, *a, **k):
        a = tuple("cpu" if isinstance(x, str) and x.startswith("cuda") else x for x in a)
        if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
            k["device"] = "cpu"
        return orig_to(self, *a, **k)

    if not torch.cuda.is_available():
        torch.Tensor.to = _to


def load_reference():
    """Import the reference's Renderer/config modules (cwd-relative config inheritance,
    src/config.py:26-32, so we chdir for the duration of the load)."""
    _install_cpu_shim()
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        from src import config as ref_config
        from src.utils.Renderer import Renderer as RefRenderer
        import src.utils.Renderer as ref_renderer_mod
        import src.common as ref_common
        cfg = ref_config.load_config("configs/Replica/room0.yaml", "configs/df_prior.yaml")
    finally:
        os.chdir(cwd)
    return SimpleNamespace(config=ref_config, Renderer=RefRenderer, renderer_mod=ref_renderer_mod,
                           common=ref_common, cfg=cfg)


def build_reference_model(ref, bound, seed):
    torch.manual_seed(seed)
    model = ref.config.get_model(ref.cfg)
    for m in (model, model.low_decoder, model.high_decoder, model.color_decoder):
        m.bound = bound                                   # src/DF_Prior.py:191-194
    return model


def reference_renderer(ref, scene):
    slam = SimpleNamespace(bound=scene.bound, vol_bnds=scene.tsdf_bnds, **scene.cam)
    return ref.Renderer(ref.cfg, None, slam)


def run_case(ref, name, stage, n_rays, seed, mode, high_std, out_scale=1.0):
    """mode: 'mapper' (grads to grids + decoders), 'tracker' (grads to the 7-vector and rays)."""
    from attentive_dfprior_b200 import synth
    torch.set_num_threads(1)                               # deterministic reductions
    scene = synth.make_scene("small", seed=seed, high_std=high_std)
    model = build_reference_model(ref, scene.bound, seed)
    if out_scale != 1.0:
        # un-saturate sigmoid(10*occ): default xavier init gives |occ| ~ 1, i.e. alpha in {0,1} and
        # vanishing gradients; a smaller output layer makes the gradient fixtures informative
        with torch.no_grad():
            for dec in (model.low_decoder, model.high_decoder):
                dec.output_linear.weight.mul_(out_scale)
    renderer = reference_renderer(ref, scene)
    grids = {k: v.clone().requires_grad_(mode == "mapper") for k, v in scene.grids.items()}
    cam = scene.cam
    c2w0 = synth.make_pose(scene, seed)
    captured = {}

    # record z_vals / pts / raw by wrapping the two symbols render_batch_ray calls
    orig_eval = renderer.eval_points
    orig_r2o = ref.renderer_mod.raw2outputs_nerf_color

    def eval_rec(p, *a, **k):
        captured["pts"] = p.detach().clone()
        raw, w = orig_eval(p, *a, **k)
        captured["raw"] = raw.detach().clone()
        return raw, w

    def r2o_rec(raw, z_vals, *a, **k):
        captured["z"] = z_vals.detach().clone()
        return orig_r2o(raw, z_vals, *a, **k)

    renderer.eval_points = eval_rec
    ref.renderer_mod.raw2outputs_nerf_color = r2o_rec
    try:
        if mode == "tracker":
            # pixels -> rays through the differentiable pose chain (src/Tracker.py:93-97)
            g = torch.Generator().manual_seed(seed + 7919)
            i = torch.randint(100, cam["W"] - 100, (n_rays,), generator=g).float()
            j = torch.randint(100, cam["H"] - 100, (n_rays,), generator=g).float()
            quat = torch.tensor([0.9, 0.1, -0.2, 0.15])
            quat = quat / quat.norm()
            cam_tensor = torch.cat([quat, c2w0[:3, 3]]).clone().requires_grad_(True)
            _get_device = torch.Tensor.get_device
            # quad2rotation does .to(quad.get_device()) (src/common.py:152): -1 on CPU
            torch.Tensor.get_device = lambda self: "cpu" if not self.is_cuda else _get_device(self)
            try:
                c2w = ref.common.get_camera_from_tensor(cam_tensor)
            finally:
                torch.Tensor.get_device = _get_device
            rays_o, rays_d = ref.common.get_rays_from_uv(i, j, c2w, cam["H"], cam["W"], cam["fx"], cam["fy"],
                                                         cam["cx"], cam["cy"], "cpu")
            rays_o.retain_grad(); rays_d.retain_grad()
            with torch.no_grad():
                gt_depth = synth.depth_along(scene, rays_o.detach(), rays_d.detach())
                zero = torch.rand(n_rays, generator=g) < 0.1
                gt_depth = torch.where(zero, torch.zeros_like(gt_depth), gt_depth)
                gt_color = torch.rand(n_rays, 3, generator=g)
            extra_in = {"pix_i": i, "pix_j": j, "cam_tensor": cam_tensor.detach().clone()}
        else:
            rays_o, rays_d, gt_depth, gt_color = synth.make_rays(scene, n_rays, seed, c2w=c2w0)
            extra_in = {}
        depth, unc, color, weight = renderer.render_batch_ray(
            grids, model, rays_d, rays_o, "cpu", scene.tsdf_volume, scene.tsdf_bnds, stage, gt_depth=gt_depth)
        out = {"depth": depth.detach().clone(), "uncertainty": unc.detach().clone(),
               "color": color.detach().clone(), "weight": weight.detach().clone(), **captured}
        if mode == "mapper":
            # src/Mapper.py:457-469 (with the attention term so every output carries gradient)
            m = gt_depth > 0
            loss = torch.abs(gt_depth[m] - depth[m]).sum() + torch.abs(weight - torch.ones(weight.shape)).sum()
            if stage == "color":
                loss = loss + 0.2 * torch.abs(gt_color - color).sum()
            loss.backward()
            grads = {"grids": {k: v.grad.clone() if v.grad is not None else None for k, v in grids.items()},
                     "params": {k: (v.grad.clone() if v.grad is not None else None)
                                for k, v in model.named_parameters()}}
        else:
            # src/Tracker.py:116-129
            u = unc.detach()
            tmp = torch.abs(gt_depth - depth) / torch.sqrt(u + 1e-10)
            m = (tmp < 10 * tmp.median()) & (gt_depth > 0)
            loss = (torch.abs(gt_depth - depth) / torch.sqrt(u + 1e-10))[m].sum() \
                + 0.5 * torch.abs(gt_color - color)[m].sum()
            loss.backward()
            grads = {"cam_tensor": cam_tensor.grad.clone(), "rays_o": rays_o.grad.clone(),
                     "rays_d": rays_d.grad.clone()}
        out["loss"] = loss.detach().clone()
    finally:
        renderer.eval_points = orig_eval
        ref.renderer_mod.raw2outputs_nerf_color = orig_r2o

    fixture = {
        "name": name, "stage": stage, "mode": mode, "seed": seed, "high_std": high_std, "out_scale": out_scale,
        "torch_version": torch.__version__,
        "inputs": {"bound": scene.bound.clone(), "grids": {k: v.detach().clone() for k, v in scene.grids.items()},
                   "tsdf_xyz": scene.tsdf_volume.permute(0, 1, 4, 3, 2).contiguous()[0, 0].clone(),   # [X,Y,Z]
                   "tsdf_bnds": scene.tsdf_bnds.clone(),
                   "state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()},
                   "rays_o": rays_o.detach().clone(), "rays_d": rays_d.detach().clone(),
                   "gt_depth": gt_depth.clone(), "gt_color": gt_color.clone(), "cam": dict(cam), **extra_in},
        "outputs": out, "grads": grads,
    }
    return fixture


CASES = [
    # name,                stage,   n_rays, seed, mode,      high_std
    ("map_low_s0",         "low",   24,     0,    "mapper",  1e-4, 0.05),
    ("map_high_s1",        "high",  24,     1,    "mapper",  1e-2, 0.05),
    ("map_color_s2",       "color", 24,     2,    "mapper",  1e-2, 0.05),
    ("map_color_s3_default", "color", 16,   3,    "mapper",  1e-4, 1.0),
    ("track_color_s4",     "color", 16,     4,    "tracker", 1e-2, 0.05),
]


def main():
    ref = load_reference()
    out_dir = os.path.join(REPO, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, stage, n, seed, mode, hs, osc in CASES:
        fx = run_case(ref, name, stage, n, seed, mode, hs, osc)
        path = os.path.join(out_dir, name + ".pt")
        torch.save(fx, path)
        o = fx["outputs"]
        print(f"{name}: depth[0:3]={o['depth'][:3].tolist()} loss={float(o['loss']):.6f} "
              f"band={(o['weight'] != 1).float().mean():.2f} -> {os.path.getsize(path)/1e3:.0f} kB")


if __name__ == "__main__":
    main()
