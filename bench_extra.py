#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (configs 3-5) on one GPU (or sharded with torchrun): tracker step,
full-frame render_img, Mesher dense occupancy query.  One JSON line per workload (rank 0).
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
    if rank == 0:
        for o in out:
            print(json.dumps(o))
    if group is not None:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
