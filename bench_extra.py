#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (configs 2-5) on one GPU (or sharded with torchrun): tracker step, the Mapper
iteration with its surrounding rows, full-frame render_img, Mesher dense occupancy query.  One JSON line per workload (rank 0).
Not the driver's contract bench (that is bench.py); results are kept under profiles/."""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)


def timeit(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--resolution", type=int, default=256)
    args = ap.parse_args()
    import attentive_dfprior_b200 as A
    from attentive_dfprior_b200 import dist as D
    from attentive_dfprior_b200 import synth
    from attentive_dfprior_b200.config import default_cfg, slam_namespace
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    torch.manual_seed(0)
    out = []

    # ---- config 3: tracker step (200 / 1000 rays, stage colour, gradient to the 7-vector only)
    scene = synth.make_scene("room0", seed=0, device=dev, high_std=1e-2, channels_last=True)
    model = A.DF().to(dev)
    for p in model.parameters():
        p.requires_grad_(False)                       # the Tracker's decoders are frozen copies
    renderer = A.Renderer(default_cfg(), None, slam_namespace(scene.bound, scene.tsdf_bnds))
    c2w = synth.make_pose(scene, 0).to(dev)
    cam = scene.cam
    for n in (200, 1000):
        g = torch.Generator().manual_seed(5)
        i = torch.randint(100, cam["W"] - 100, (n,), generator=g).float().to(dev)
        j = torch.randint(100, cam["H"] - 100, (n,), generator=g).float().to(dev)
        quat = torch.tensor([0.9, 0.1, -0.2, 0.15]); quat = quat / quat.norm()
        cam_t = torch.cat([quat, c2w[:3, 3].cpu()]).to(dev).requires_grad_(True)
        with torch.no_grad():
            ro0, rd0 = A.common.get_rays_from_uv(i, j, A.common.get_camera_from_tensor(cam_t), cam["H"], cam["W"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], dev)
            gt = synth.depth_along(scene, ro0, rd0)
        gc = torch.rand(n, 3, device=dev)

        def tracker_step():
            cam_t.grad = None
            ro, rd = A.common.get_rays_from_uv(i, j, A.common.get_camera_from_tensor(cam_t), cam["H"], cam["W"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], dev)
            d, u, c, _ = renderer.render_batch_ray(scene.grids, model, rd, ro, dev, scene.tsdf_volume, scene.tsdf_bnds, "color", gt_depth=gt)
            u = u.detach()
            tmp = torch.abs(gt - d) / torch.sqrt(u + 1e-10)
            m = (tmp < 10 * tmp.median()) & (gt > 0)
            loss = (torch.abs(gt - d) / torch.sqrt(u + 1e-10))[m].sum() + 0.5 * torch.abs(gc - c)[m].sum()
            loss.backward()
        ms = timeit(tracker_step, 3, 9)
        out.append({"workload": "tracker_step", "rays": n, "stage": "color", "ms_per_step": ms, "rays_per_s": n / ms * 1e3})

    # ---- config 2 as the Mapper runs it (rank 0): frustum masks, then per iteration ray assembly -> render -> loss ->
    #      backward -> masked-grid Adam, against the same iteration with the upstream torch flow
    #      (index_put of val_grad, torch get_samples, torch.optim.Adam) around the same renderer
    if rank == 0:
        from attentive_dfprior_b200 import mapper_ops as M
        import time
        mdl = A.DF().to(dev)
        H, W = cam["H"], cam["W"]
        intr = (cam["fx"], cam["fy"], cam["cx"], cam["cy"])
        frames = []
        for f in range(5):
            pose = synth.make_pose(scene, f).to(dev)
            frames.append((synth.make_depth_image(scene, pose, device=dev, seed=f), torch.rand(H, W, 3, device=dev), pose))
        n_pix = 200
        keys = ("grid_low", "grid_high", "grid_color")
        t0 = time.perf_counter()
        masks = {k: M.frustum_mask(frames[0][2], scene.bound, scene.grids[k].shape[2:], frames[0][0], H, W, *intr) for k in keys}
        torch.cuda.synchronize()
        ms_mask = timeit(lambda: [M.frustum_mask(frames[0][2], scene.bound, scene.grids[k].shape[2:], frames[0][0], H, W, *intr) for k in keys], 2, 5)
        frac = {k: float(masks[k].float().mean()) for k in keys}
        lrs = {"grid_low": 0.1, "grid_high": 0.005, "grid_color": 0.0}     # upstream never sets the colour-grid group's lr: it stays 0 (Mapper.py:372-376, 392-398)

        grids_f = {k: scene.grids[k].detach().clone(memory_format=torch.preserve_format).requires_grad_(True) for k in keys}
        opt_g = M.MaskedGridAdam(grids_f, masks)
        opt_d = torch.optim.Adam(mdl.parameters(), lr=0.005, fused=True)

        def fused_iter():
            idx = torch.randint(H * W, (5, n_pix), device=dev)
            ro, rd, gd, gc = M.assemble_rays(frames, idx, H, W, *intr, scene.bound)
            d, u, c, w = renderer.render_batch_ray(grids_f, mdl, rd, ro, dev, scene.tsdf_volume, scene.tsdf_bnds, "color", gt_depth=gd)
            loss = M.mapper_loss(d, c, w, gd, gc, 0.2, False, True)
            loss.backward()
            opt_g.step(lrs)
            opt_d.step(); opt_d.zero_grad()
        ms_fused = timeit(fused_iter, 5, 15)

        # upstream flow of the rows (Mapper.py:351-361, 383-388, 414-449, 457-482) in torch on the same device
        vals = {k: scene.grids[k].detach().clone(memory_format=torch.preserve_format) for k in keys}
        bmask = {k: masks[k].bool()[None, None].expand(vals[k].shape) for k in keys}
        vgrad = {k: vals[k][bmask[k]].clone().requires_grad_(True) for k in keys}
        opt_t = torch.optim.Adam([{"params": list(mdl.parameters()), "lr": 0.005}] + [{"params": [vgrad[k]], "lr": lrs[k]} for k in keys])
        bound_d = scene.bound.to(dev)

        def torch_iter():
            c = {}
            for k in keys:
                v = vals[k]
                v[bmask[k]] = vgrad[k]
                c[k] = v
            ros, rds, gds, gcs = [], [], [], []
            for depth, color, pose in frames:
                o, dd, gdep, gcol = A.common.get_samples(0, H, 0, W, n_pix, H, W, *intr, pose, depth, color, dev)
                ros.append(o.float()); rds.append(dd.float()); gds.append(gdep.float()); gcs.append(gcol.float())
            ro, rd, gd, gc = torch.cat(ros), torch.cat(rds), torch.cat(gds), torch.cat(gcs)
            with torch.no_grad():
                t = (bound_d.unsqueeze(0) - ro.unsqueeze(-1)) / rd.unsqueeze(-1)
                t, _ = torch.min(torch.max(t, dim=2)[0], dim=1)
                inside = t >= gd
            ro, rd, gd, gc = ro[inside], rd[inside], gd[inside], gc[inside]
            d, u, col, w = renderer.render_batch_ray(c, mdl, rd, ro, dev, scene.tsdf_volume, scene.tsdf_bnds, "color", gt_depth=gd)
            dm = gd > 0
            loss = torch.abs(gd[dm] - d[dm]).sum() + 0.2 * torch.abs(gc - col).sum()
            loss.backward()
            opt_t.step(); opt_t.zero_grad()
            for k in keys:
                v = vals[k].detach()
                v[bmask[k]] = vgrad[k].clone().detach()
                vals[k] = v
        ms_torch = timeit(torch_iter, 5, 15)

        # the Adam kernel alone on the colour grid: 7 float streams over the masked voxels + the mask bytes
        gk = grids_f["grid_color"]
        gbuf = torch.randn_like(gk)
        ms_adam = timeit(lambda: opt_g.step({"grid_color": 0.005}, grads={"grid_color": gbuf}), 3, 9)
        nvox = gk.shape[2] * gk.shape[3] * gk.shape[4]
        adam_bytes = frac["grid_color"] * nvox * 32 * 4 * 7 + nvox
        out.append({"workload": "mapper_iteration_5x200px_color", "ms_fused_rows": ms_fused, "ms_torch_rows": ms_torch,
                    "frustum_masks_3grids_ms": ms_mask, "mask_fraction": frac,
                    "masked_adam_color_grid_ms": ms_adam, "masked_adam_GBps": adam_bytes / ms_adam / 1e6})
        del grids_f, vals, vgrad, opt_g, opt_t, bmask

    # ---- config 4: full-frame render_img 1200x680, chunk-sharded across ranks
    gt_img = synth.make_depth_image(scene, c2w, device=dev, seed=1)
    def render():
        return D.render_img_sharded(renderer, scene.grids, model, c2w, dev, scene.tsdf_volume, scene.tsdf_bnds, "color", gt_img, group=group)
    ms = timeit(render, 1, 3)
    nr = cam["H"] * cam["W"]
    out.append({"workload": "render_img_1200x680", "rays": nr, "stage": "color", "n_gpus": world, "ms_per_frame": ms,
                "rays_per_s": nr / ms * 1e3, "algorithmic_GBps": nr / ms * 1e3 * 149240 / 1e9})
    del scene, gt_img
    torch.cuda.empty_cache()

    # ---- config 5: Mesher dense occupancy query over the ScanNet scene0000 bound, voxel-block sharded
    scene = synth.make_scene("scannet0000", seed=0, device=dev, high_std=1e-2, channels_last=True)
    cfg = default_cfg(); cfg["meshing"]["resolution"] = args.resolution
    cfg["mapping"]["marching_cubes_bound"] = synth.SCANNET0000_BOUND
    mesher = A.Mesher(cfg, None, slam_namespace(scene.bound, scene.tsdf_bnds))
    pts = mesher.get_grid_uniform(args.resolution)["grid_points"].to(dev)
    def query():
        return mesher.query_occupancy(scene.grids, model, scene.tsdf_volume, dev, points=pts, group=group, gather=world > 1)
    ms = timeit(query, 1, 3)
    npts = pts.shape[0]
    out.append({"workload": f"mesher_query_{args.resolution}^3_scannet0000", "points": npts, "stage": "high", "n_gpus": world,
                "ms": ms, "points_per_s": npts / ms * 1e3, "algorithmic_GBps": npts / ms * 1e3 * 2096 / 1e9})
    if world == 1:
        # marching cubes over the queried lattice (replaces skimage.measure.marching_cubes, Mesher.py:463-480)
        from attentive_dfprior_b200 import mesher as MS
        z = query()
        r = args.resolution
        vol = z.reshape(r, r, r).permute(1, 0, 2).contiguous()
        level = float(vol[vol < 50].median())                     # random-init decoders: take a level the lattice actually crosses
        def mc():
            return MS.marching_cubes(vol, level, (0.03, 0.03, 0.03), (0.0, 0.0, 0.0))
        ms_mc = timeit(mc, 1, 3)
        v, f = mc()
        out.append({"workload": f"marching_cubes_{r}^3", "cells": (r - 1) ** 3, "triangles": int(f.shape[0]), "vertices": int(v.shape[0]),
                    "ms": ms_mc, "note": "count + prefix sum + emit + weld (torch.unique); one D2H for the triangle count"})
    if rank == 0:
        for o in out:
            print(json.dumps(o))
    if group is not None:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
