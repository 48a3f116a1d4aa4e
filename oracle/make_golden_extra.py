"""Generate tests/golden/extra_*.pt: reference outputs for the entry points and branches round 1 left unpinned.

TEST INFRASTRUCTURE ONLY (runs in the authoring container where /root/reference is mounted; see make_golden.py).

  extra_nodepth_color   Renderer.render_batch_ray with gt_depth=None (src/utils/Renderer.py:140-142,161-162: N_surface=0,
                        near=0.01, far=far_bb), loss with an uncertainty term so d(uncertainty) flows
  extra_unc_high        stage 'high', loss |gt-d|/sqrt(unc) with the uncertainty NOT detached (gradient through var)
  extra_tsdf_points     Renderer.eval_points_tsdf / sample_grid_tsdf (Renderer.py:73-107) on f32 and f64 points
  extra_mesher          the reference's own Mesher.get_grid_uniform and Mesher.eval_points (src/utils/Mesher.py:286-326,
                        365-393), called unbound on a namespace (the module's open3d / skimage / trimesh / matplotlib /
                        dataset imports are stubbed: none of them is touched by these two methods); also the
                        scannet0000 256^3 lattice as a digest (sha256 of the f32 bytes + sampled rows)

Usage:  python oracle/make_golden_extra.py
"""
from __future__ import annotations

import hashlib
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))
import make_golden as MG  # noqa: E402


def load_reference_mesher():
    """Import src.utils.Mesher with its geometry / plotting / dataset dependencies stubbed out."""
    for name in ("open3d", "skimage", "skimage.measure", "trimesh", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    ds = types.ModuleType("src.utils.datasets")
    ds.get_dataset = lambda *a, **k: None
    sys.modules.setdefault("src.utils.datasets", ds)
    cwd = os.getcwd()
    os.chdir(MG.REF)
    try:
        import src.utils.Mesher as M
    finally:
        os.chdir(cwd)
    return M.Mesher


def render_case(ref, name, stage, n_rays, seed, with_depth, loss_kind):
    from attentive_dfprior_b200 import synth
    torch.set_num_threads(1)
    scene = synth.make_scene("small", seed=seed, high_std=1e-2)
    model = MG.build_reference_model(ref, scene.bound, seed)
    with torch.no_grad():
        for dec in (model.low_decoder, model.high_decoder):
            dec.output_linear.weight.mul_(0.05)
    renderer = MG.reference_renderer(ref, scene)
    grids = {k: v.clone().requires_grad_(True) for k, v in scene.grids.items()}
    rays_o, rays_d, gt_depth, gt_color = synth.make_rays(scene, n_rays, seed)
    captured = {}
    orig_r2o = ref.renderer_mod.raw2outputs_nerf_color

    def r2o_rec(raw, z_vals, *a, **k):
        captured["z"] = z_vals.detach().clone()
        captured["raw"] = raw.detach().clone()
        return orig_r2o(raw, z_vals, *a, **k)

    ref.renderer_mod.raw2outputs_nerf_color = r2o_rec
    try:
        depth, unc, color, weight = renderer.render_batch_ray(
            grids, model, rays_d, rays_o, "cpu", scene.tsdf_volume, scene.tsdf_bnds, stage,
            gt_depth=gt_depth if with_depth else None)
    finally:
        ref.renderer_mod.raw2outputs_nerf_color = orig_r2o
    if loss_kind == "nodepth":
        loss = 0.05 * depth.sum() + unc.sum() + 0.2 * torch.abs(gt_color - color).sum() \
            + torch.abs(weight - torch.ones(weight.shape)).sum()
    elif loss_kind == "unc":
        m = gt_depth > 0
        loss = (torch.abs(gt_depth - depth) / torch.sqrt(unc + 1e-10))[m].sum()
    else:
        raise ValueError(loss_kind)
    loss.backward()
    return {
        "name": name, "stage": stage, "seed": seed, "with_depth": with_depth, "loss_kind": loss_kind,
        "torch_version": torch.__version__,
        "inputs": {"bound": scene.bound.clone(), "grids": {k: v.detach().clone() for k, v in scene.grids.items()},
                   "tsdf_xyz": scene.tsdf_volume.permute(0, 1, 4, 3, 2).contiguous()[0, 0].clone(),
                   "tsdf_bnds": scene.tsdf_bnds.clone(),
                   "state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()},
                   "rays_o": rays_o.clone(), "rays_d": rays_d.clone(), "gt_depth": gt_depth.clone(),
                   "gt_color": gt_color.clone(), "cam": dict(scene.cam)},
        "outputs": {"depth": depth.detach().clone(), "uncertainty": unc.detach().clone(), "color": color.detach().clone(),
                    "weight": weight.detach().clone(), "loss": loss.detach().clone(), **captured},
        "grads": {"grids": {k: (v.grad.clone() if v.grad is not None else None) for k, v in grids.items()},
                  "params": {k: (v.grad.clone() if v.grad is not None else None) for k, v in model.named_parameters()}},
    }


def tsdf_case(ref, seed=7):
    from attentive_dfprior_b200 import synth
    scene = synth.make_scene("small", seed=seed, high_std=1e-2)
    renderer = MG.reference_renderer(ref, scene)
    renderer.points_batch_size = 300                      # several batches (Renderer.py:95-106)
    g = torch.Generator().manual_seed(seed)
    tb = scene.tsdf_bnds
    lo, hi = tb[:, 0], tb[:, 1]
    p64 = torch.rand(700, 3, generator=g, dtype=torch.float64) * (hi - lo) * 1.2 + lo - 0.1 * (hi - lo)
    p32 = p64.float()
    out = {}
    for tag, p in (("f64", p64), ("f32", p32)):
        out["vals_" + tag] = renderer.eval_points_tsdf(p.clone(), scene.tsdf_volume, "cpu").detach().clone()
        out["grid_" + tag] = renderer.sample_grid_tsdf(p.clone(), scene.tsdf_volume, "cpu").detach().clone()
    return {"name": "extra_tsdf_points", "seed": seed,
            "inputs": {"tsdf_xyz": scene.tsdf_volume.permute(0, 1, 4, 3, 2).contiguous()[0, 0].clone(),
                       "tsdf_bnds": tb.clone(), "bound": scene.bound.clone(), "p64": p64, "p32": p32},
            "outputs": out}


def mesher_case(ref, seed=8):
    from attentive_dfprior_b200 import synth
    RefMesher = load_reference_mesher()
    scene = synth.make_scene("small", seed=seed, high_std=1e-2)
    model = MG.build_reference_model(ref, scene.bound, seed)
    with torch.no_grad():
        for dec in (model.low_decoder, model.high_decoder):
            dec.output_linear.weight.mul_(0.05)
    mc_bound = np.array(synth.SMALL_BOUND) * 1.0
    me = SimpleNamespace(marching_cubes_bound=torch.from_numpy(mc_bound), bound=scene.bound, points_batch_size=700)
    res = 12
    grid = RefMesher.get_grid_uniform(me, res)
    pts = grid["grid_points"]
    with torch.no_grad():
        rets = {st: RefMesher.eval_points(me, pts.clone(), model, scene.tsdf_volume, scene.tsdf_bnds, scene.grids, st, "cpu")
                .detach().clone() for st in ("low", "high", "color")}
    # scannet0000 lattice at the reference's meshing resolution (configs/df_prior.yaml meshing.resolution = 256)
    import yaml
    cfg = yaml.safe_load(open(os.path.join(MG.REF, "configs/ScanNet/scene0000.yaml")))
    mcb = np.array(cfg["mapping"]["marching_cubes_bound"])
    me2 = SimpleNamespace(marching_cubes_bound=torch.from_numpy(mcb * 1.0))
    big = RefMesher.get_grid_uniform(me2, 256)["grid_points"]
    idx = torch.arange(0, big.shape[0], 65521)
    digest = hashlib.sha256(big.numpy().tobytes()).hexdigest()
    return {"name": "extra_mesher", "seed": seed,
            "inputs": {"bound": scene.bound.clone(), "grids": {k: v.detach().clone() for k, v in scene.grids.items()},
                       "tsdf_xyz": scene.tsdf_volume.permute(0, 1, 4, 3, 2).contiguous()[0, 0].clone(),
                       "tsdf_bnds": scene.tsdf_bnds.clone(),
                       "state_dict": {k: v.detach().clone() for k, v in model.state_dict().items()},
                       "marching_cubes_bound": torch.from_numpy(mc_bound), "resolution": res,
                       "scannet0000_marching_cubes_bound": torch.from_numpy(mcb * 1.0)},
            "outputs": {"grid_points": pts.clone(), "xyz": [torch.as_tensor(np.asarray(a, dtype=np.float64)).clone() for a in grid["xyz"]],
                        "ret": rets, "scannet0000_sha256": digest, "scannet0000_rows": idx,
                        "scannet0000_sample": big[idx].clone(), "scannet0000_n": big.shape[0]}}


def main():
    ref = MG.load_reference()
    out_dir = os.path.join(REPO, "tests", "golden")
    cases = [
        render_case(ref, "extra_nodepth_color", "color", 20, 5, False, "nodepth"),
        render_case(ref, "extra_unc_high", "high", 20, 6, True, "unc"),
        tsdf_case(ref),
        mesher_case(ref),
    ]
    for fx in cases:
        path = os.path.join(out_dir, fx["name"] + ".pt")
        torch.save(fx, path)
        print(fx["name"], "->", f"{os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
