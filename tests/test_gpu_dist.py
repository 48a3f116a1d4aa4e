"""GPU, world_size 2 (NCCL): the ray-sharded mapping step gives the same outputs and the same (all-reduced)
gradients as the single-GPU step on the whole batch.  Skipped when fewer than 2 GPUs are visible."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import attentive_dfprior_b200 as A
        from attentive_dfprior_b200 import dist as D
        from attentive_dfprior_b200 import synth
        from attentive_dfprior_b200.config import default_cfg, slam_namespace
        torch.manual_seed(0)
        scene = synth.make_scene("small", seed=3, device=dev, high_std=1e-2, channels_last=True)
        model = A.DF().to(dev)
        with torch.no_grad():
            for dec in (model.low_decoder, model.high_decoder):
                dec.output_linear.weight.mul_(0.05)
        renderer = A.Renderer(default_cfg(), None, slam_namespace(scene.bound, scene.tsdf_bnds))
        n = 512
        rays_o, rays_d, gt, gc = synth.make_rays(scene, n, seed=5, device=dev)      # identical on every rank

        def run(lo, hi, group, **kw):
            grids = {k: v.clone().requires_grad_(True) for k, v in scene.grids.items()}
            for p in model.parameters():
                p.grad = None
            d, u, c, w = D.render_batch_ray_sharded(renderer, grids, model, rays_d[lo:hi], rays_o[lo:hi], scene.tsdf_volume,
                                                    scene.tsdf_bnds, "color", gt[lo:hi], group=group, **kw)
            g = gt[lo:hi]
            m = g > 0
            (torch.abs(g[m] - d[m]).sum() + 0.2 * torch.abs(gc[lo:hi] - c).sum()).backward()
            return d.detach(), {k: v.grad.clone() for k, v in grids.items()}, [p.grad.clone() for p in model.parameters()]

        lo, hi = D.shard_range(n, rank, world)
        d_full, gg_full, gp_full = run(0, n, None)
        def rel(a, b):
            return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)
        ok, worst = True, 0.0
        # bucketed all-reduce on the side stream (default), one blocking all-reduce, and the caller-supplied global depth
        # maximum (no all-reduce(max)): all three must reproduce the single-GPU step
        for kw in ({}, {"overlap_allreduce": False}, {"depth_max": gt.max()}):
            d_sh, gg_sh, gp_sh = run(lo, hi, dist.group.WORLD, **kw)
            ok = ok and bool(torch.equal(d_sh, d_full[lo:hi]))     # per-ray outputs do not depend on the sharding (global depth max)
            worst = max([worst] + [rel(gg_sh[k], gg_full[k]) for k in gg_full] +
                        [rel(a, b) for a, b in zip(gp_sh, gp_full) if float(b.abs().max()) > 0])

        # BASELINE config 4: chunk-sharded render_img against the single-GPU frame (bit-exact: per-chunk maxima are
        # chunk-local, no collective on the data path)
        cam = dict(H=24, W=40, fx=30.0, fy=30.0, cx=19.5, cy=11.5)
        r2 = A.Renderer(default_cfg(), None, slam_namespace(scene.bound, scene.tsdf_bnds, cam), ray_batch_size=96)
        scene.cam = cam
        c2w = synth.make_pose(scene, 2).to(dev)
        gt_img = synth.make_depth_image(scene, c2w, device=dev, seed=3)
        full = r2.render_img(scene.grids, model, c2w, dev, scene.tsdf_volume, scene.tsdf_bnds, "color", gt_depth=gt_img)
        shard = D.render_img_sharded(r2, scene.grids, model, c2w, dev, scene.tsdf_volume, scene.tsdf_bnds, "color", gt_img,
                                     group=dist.group.WORLD)
        ok_img = all(bool(torch.equal(a, b)) for a, b in zip(full, shard))

        # BASELINE config 5: voxel-block sharded Mesher query against the single-GPU query (bit-exact)
        cfg = default_cfg()
        cfg["meshing"]["resolution"] = 20
        cfg["mapping"]["marching_cubes_bound"] = synth.SMALL_BOUND
        mesher = A.Mesher(cfg, None, slam_namespace(scene.bound, scene.tsdf_bnds), points_batch_size=1500)
        z_full = mesher.query_occupancy(scene.grids, model, scene.tsdf_volume, dev)
        z_sh = mesher.query_occupancy(scene.grids, model, scene.tsdf_volume, dev, group=dist.group.WORLD)
        ok_mesh = bool(torch.equal(z_full, z_sh)) and z_full.shape[0] == 20 ** 3
        ret[rank] = (ok, worst, ok_img, ok_mesh)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_step_matches_single_gpu():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29611, ret), nprocs=2, join=True)
    for r in range(2):
        ok, worst, ok_img, ok_mesh = ret[r]
        assert ok, "sharded outputs differ from the single-GPU outputs"
        assert worst <= 1e-4, worst
        assert ok_img, "chunk-sharded render_img differs from the single-GPU frame"
        assert ok_mesh, "voxel-block sharded Mesher query differs from the single-GPU query"
