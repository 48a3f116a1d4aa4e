"""GPU parity tests: the CUDA path (through the Python drop-in classes -> ctypes -> C ABI)
against (a) the reference-generated golden fixtures and (b) the CPU oracle on seeded inputs.

Tolerances are the north-star's: bit-exact sample depths / masks, 1e-4 relative forward,
1e-3 relative gradients (relative to the tensor's max magnitude; an element-wise relative gate
on raw occupancies is not meaningful).
"""
import copy

import pytest
import torch

from conftest import GOLDEN_CASES, load_golden
from oracle import dfprior_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


MARGINS = {}


def rel_err(a, b, tag=None):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    e = float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)
    if tag is not None:
        MARGINS[tag] = max(MARGINS.get(tag, 0.0), e)
    return e


def within(mine, ref, truth, tol, tag=None):
    """``mine`` within ``tol`` of the reference value ``ref`` (relative to the tensor's max).  Where the fp32 reference is
    itself further than that from the f64 evaluation of the same formulas on the same input (``truth``; an
    ill-conditioned input, e.g. one sample sitting on a ReLU edge: SURVEY.md appendix A.5 asks for exactly this
    comparison), the gate is widened by the reference's own distance to the truth and ``mine`` must be within ``tol``
    of the truth."""
    e = rel_err(mine, ref, tag)
    if e <= tol:
        return True
    e_ref = rel_err(ref, truth)
    return e <= tol + e_ref and rel_err(mine, truth) <= tol


@pytest.fixture(scope="module", autouse=True)
def _dump_margins():
    """With DFP_PARITY_LOG=<file> the measured errors of the tagged comparisons are written out (profiles/ evidence)."""
    yield
    import json
    import os
    path = os.environ.get("DFP_PARITY_LOG")
    if path and MARGINS:
        with open(path, "w") as f:
            json.dump(MARGINS, f, indent=1, sort_keys=True)


def build(inp, channels_last=False):
    import attentive_dfprior_b200 as A
    from attentive_dfprior_b200.config import default_cfg, slam_namespace
    model = A.DF().to(DEV)
    model.load_state_dict(inp["state_dict"])
    bound = inp["bound"]
    for m in (model, model.low_decoder, model.high_decoder, model.color_decoder):
        m.bound = bound
    tsdf = inp["tsdf_volume"].to(DEV)
    tb = inp["tsdf_bnds"].to(DEV)
    grids = {}
    for k, v in inp["grids"].items():
        g = v.to(DEV)
        if channels_last:
            g = g.contiguous(memory_format=torch.channels_last_3d)
        grids[k] = g
    renderer = A.Renderer(default_cfg(), None, slam_namespace(bound, tb, inp.get("cam")))
    return model, renderer, grids, tsdf, tb


@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_render_matches_reference_fixture(name, channels_last):
    import attentive_dfprior_b200 as A
    from attentive_dfprior_b200 import functional as Fn
    fx = load_golden(name)
    inp, ref, stage = fx["inputs"], fx["outputs"], fx["stage"]
    model, renderer, grids, tsdf, tb = build(inp, channels_last)
    grids = {k: v.requires_grad_(fx["mode"] == "mapper") for k, v in grids.items()}
    gt = inp["gt_depth"].to(DEV)
    gt_color = inp["gt_color"].to(DEV)
    if fx["mode"] == "tracker":
        for prm in model.parameters():          # the Tracker's decoders are frozen copies: skip their gradients
            prm.requires_grad_(False)
        cam = inp["cam"]
        ct = inp["cam_tensor"].to(DEV).requires_grad_(True)
        c2w = A.common.get_camera_from_tensor(ct)
        rays_o, rays_d = A.common.get_rays_from_uv(inp["pix_i"].to(DEV), inp["pix_j"].to(DEV), c2w, cam["H"], cam["W"],
                                                   cam["fx"], cam["fy"], cam["cx"], cam["cy"], DEV)
        assert rel_err(rays_d, inp["rays_d"]) < 1e-6
    else:
        rays_o, rays_d = inp["rays_o"].to(DEV), inp["rays_d"].to(DEV)
    N, S = rays_o.shape[0], 48

    # aux outputs for the bit-exact checks (on exactly the fixture's rays)
    (d0, u0, c0, w0), aux = Fn.render_batch_ray(grids, model, inp["rays_d"].to(DEV), inp["rays_o"].to(DEV), tsdf, tb,
                                                inp["bound"], stage, gt_depth=gt, want_aux=True)
    assert torch.equal(aux["z"].cpu(), ref["z"]), "sorted sample depths must be bit-exact"
    raw_ref = ref["raw"].reshape(N, S, 4)
    raw = aux["raw"].cpu()
    assert torch.equal(raw[..., 3] == 100, raw_ref[..., 3] == 100), "in-bound mask must be exact"
    assert torch.equal(w0.cpu().reshape(N, S) != 1, ref["weight"].reshape(N, S) != 1), "band mask must be exact"
    rms = float(raw_ref[..., 3][raw_ref[..., 3] != 100].pow(2).mean().sqrt())
    assert float((raw - raw_ref).abs().max()) <= 1e-4 * max(rms, 1.0) + 2e-4 * rms

    depth, unc, color, weight = renderer.render_batch_ray(grids, model, rays_d, rays_o, DEV, tsdf, tb, stage, gt_depth=gt)
    assert depth.dtype == torch.float64 and unc.dtype == torch.float64 and color.dtype == torch.float32
    assert weight.shape == (N, S, 1)
    assert rel_err(depth, ref["depth"]) <= 1e-4
    assert rel_err(unc, ref["uncertainty"]) <= 1e-4
    if stage == "color":
        assert rel_err(color, ref["color"]) <= 1e-4
    else:
        assert float(color.abs().max()) == 0
    assert rel_err(weight, ref["weight"]) <= 1e-4

    if fx["mode"] == "mapper":
        m = gt > 0
        loss = torch.abs(gt[m] - depth[m]).sum() + torch.abs(weight - torch.ones_like(weight)).sum()
        if stage == "color":
            loss = loss + 0.2 * torch.abs(gt_color - color).sum()
        loss.backward()
        assert abs(float(loss) - float(ref["loss"])) <= 1e-4 * abs(float(ref["loss"]))
        for k, g in fx["grads"]["grids"].items():
            if g is None:
                assert grids[k].grad is None or float(grids[k].grad.abs().max()) == 0
            else:
                assert grids[k].grad.shape == g.shape
                assert rel_err(grids[k].grad, g) <= 1e-3, k
        worst = ("", 0.0)
        named = dict(model.named_parameters())
        for k, g in fx["grads"]["params"].items():
            mine = named[k].grad
            if g is None or float(g.abs().max()) == 0:
                assert mine is None or float(mine.abs().max()) == 0, k
                continue
            e = rel_err(mine, g)
            if e > worst[1]:
                worst = (k, e)
        assert worst[1] <= 1e-3, worst
    else:
        u = unc.detach()
        tmp = torch.abs(gt - depth) / torch.sqrt(u + 1e-10)
        m = (tmp < 10 * tmp.median()) & (gt > 0)
        loss = (torch.abs(gt - depth) / torch.sqrt(u + 1e-10))[m].sum() + 0.5 * torch.abs(gt_color - color)[m].sum()
        loss.backward()
        assert abs(float(loss) - float(ref["loss"])) <= 1e-4 * abs(float(ref["loss"]))
        assert rel_err(ct.grad, fx["grads"]["cam_tensor"]) <= 1e-3
        assert all(p.grad is None for p in model.parameters())


def _random_points(bound, n, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    lo, hi = bound[:, 0], bound[:, 1]
    p = torch.rand(n, 3, generator=g, dtype=torch.float64) * (hi - lo) * 1.1 + lo - 0.05 * (hi - lo)
    return p.to(dtype)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("stage", ["low", "high", "color"])
def test_points_query_matches_oracle(stage, dtype):
    """DF.forward and Renderer.eval_points on f64 (ray path) and f32 (Mesher) points, fwd + bwd."""
    fx = load_golden("map_color_s2")
    inp = fx["inputs"]
    model, renderer, grids, tsdf, tb = build(inp)
    p_cpu = _random_points(inp["bound"], 3000, 11, dtype)
    torch.set_num_threads(8)
    gr = torch.randn(3000, 4, generator=torch.Generator().manual_seed(5))
    gw = torch.randn(3000, generator=torch.Generator().manual_seed(6))

    def oracle(prec):              # the reference's arithmetic (prec = f32) or the same formulas in f64 on the same points
        params = {k: v.clone().to(prec).requires_grad_(True) for k, v in inp["state_dict"].items()}
        grids_cpu = {k: v.clone().to(prec).requires_grad_(True) for k, v in inp["grids"].items()}
        p_ref = (p_cpu.clone() if prec == torch.float32 else p_cpu.double().clone()).requires_grad_(True)
        raw_ref, w_ref = O.eval_points(params, p_ref, grids_cpu, inp["tsdf_volume"].to(prec), inp["tsdf_bnds"], inp["bound"], stage)
        ((raw_ref * gr.to(prec)).sum() + (w_ref * gw.to(prec)).sum()).backward()
        return raw_ref.detach(), w_ref.detach(), params, grids_cpu, p_ref

    raw_ref, w_ref, params, grids_cpu, p_ref = oracle(torch.float32)
    _, _, params_t, grids_t, p_t = oracle(torch.float64)

    grids = {k: v.requires_grad_(True) for k, v in grids.items()}
    p = p_cpu.detach().to(DEV).requires_grad_(True)
    raw, w = renderer.eval_points(p, model, tsdf, tb, grids, stage, DEV)
    assert torch.equal(raw[:, 3].cpu() == 100, raw_ref[:, 3] == 100)
    assert torch.equal(w.cpu() != 1, w_ref != 1)
    scale = float(raw_ref[:, 3][raw_ref[:, 3] != 100].abs().max())
    assert float((raw.cpu() - raw_ref).abs().max()) <= 1e-4 * max(scale, 1.0)
    assert float((w.cpu() - w_ref).abs().max()) <= 1e-4
    ((raw * gr.to(DEV)).sum() + (w * gw.to(DEV)).sum()).backward()
    for k, v in grids_cpu.items():
        if v.grad is not None and float(v.grad.abs().max()) > 0:
            assert within(grids[k].grad, v.grad, grids_t[k].grad, 1e-3), k
    named = dict(model.named_parameters())
    for k, v in params.items():
        if v.grad is not None and float(v.grad.abs().max()) > 0:
            assert within(named[k].grad, v.grad, params_t[k].grad, 1e-3), k
    assert within(p.grad, p_ref.grad, p_t.grad, 1e-3)

    # bare DF.forward: no out-of-bound override
    raw2, w2 = model(p.detach().unsqueeze(0), c_grid=grids, tsdf_volume=tsdf, tsdf_bnds=tb, stage=stage)
    raw2_ref, _ = O.df_forward(inp["state_dict"], p_cpu, inp["grids"], inp["tsdf_volume"], inp["tsdf_bnds"], inp["bound"], stage)
    assert float((raw2.cpu() - raw2_ref).abs().max()) <= 1e-4 * max(scale, 1.0)


def test_render_img_chunked_matches_oracle():
    """Chunk-exact render_img: every ray_batch_size chunk uses its own max(gt_depth) (Renderer.py:294-306)."""
    import attentive_dfprior_b200 as A
    from attentive_dfprior_b200 import synth
    from attentive_dfprior_b200.config import default_cfg, slam_namespace
    fx = load_golden("map_color_s2")
    inp = fx["inputs"]
    model, _, grids, tsdf, tb = build(inp)
    cam = dict(H=20, W=30, fx=25.0, fy=25.0, cx=14.5, cy=9.5)
    renderer = A.Renderer(default_cfg(), None, slam_namespace(inp["bound"], tb, cam), ray_batch_size=100)
    scene = synth.make_scene("small", seed=2, high_std=1e-2)
    scene.cam = cam
    c2w = synth.make_pose(scene, 2)
    gt = synth.make_depth_image(scene, c2w, seed=3)
    d, u, c = renderer.render_img(grids, model, c2w.to(DEV), DEV, tsdf, tb, "color", gt_depth=gt.to(DEV))
    d_ref, u_ref, c_ref = O.render_img(inp["state_dict"], inp["grids"], c2w, inp["tsdf_volume"], inp["tsdf_bnds"],
                                       inp["bound"], "color", gt, ray_batch_size=100, **cam)
    assert d.shape == (20, 30) and c.shape == (20, 30, 3)
    assert rel_err(d, d_ref) <= 1e-4 and rel_err(u, u_ref) <= 1e-4 and rel_err(c, c_ref) <= 1e-4


def test_room0_mapping_batch_matches_oracle():
    """Config 1: Replica-room0-shaped mapping batch (1000 rays), forward + a few gradient tensors."""
    import attentive_dfprior_b200 as A
    from attentive_dfprior_b200 import synth
    from attentive_dfprior_b200.config import default_cfg, slam_namespace
    torch.manual_seed(0)
    scene = synth.make_scene("room0", seed=0, high_std=1e-2, voxel=8.0 / 256)   # coarser TSDF keeps the CPU oracle quick
    model = A.DF()
    with torch.no_grad():
        for dec in (model.low_decoder, model.high_decoder):
            dec.output_linear.weight.mul_(0.05)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    rays_o, rays_d, gt, gt_color = synth.make_rays(scene, 1000, seed=1)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    grids_cpu = {k: v.clone().requires_grad_(True) for k, v in scene.grids.items()}
    torch.set_num_threads(8)
    d_ref, u_ref, c_ref, w_ref = O.render_batch_ray(params, grids_cpu, rays_d, rays_o, scene.tsdf_volume, scene.tsdf_bnds,
                                                    scene.bound, "color", gt)
    O.mapper_loss(d_ref, c_ref, w_ref, gt, gt_color, "color").backward()

    model = model.to(DEV)
    sc = scene.to(DEV)
    grids = {k: v.contiguous(memory_format=torch.channels_last_3d).requires_grad_(True) for k, v in sc.grids.items()}
    renderer = A.Renderer(default_cfg(), None, slam_namespace(scene.bound, sc.tsdf_bnds))
    d, u, c, w = renderer.render_batch_ray(grids, model, rays_d.to(DEV), rays_o.to(DEV), DEV, sc.tsdf_volume, sc.tsdf_bnds,
                                           "color", gt_depth=gt.to(DEV))
    assert rel_err(d, d_ref, "room0_1000.depth") <= 1e-4 and rel_err(u, u_ref, "room0_1000.uncertainty") <= 1e-4
    assert rel_err(c, c_ref, "room0_1000.color") <= 1e-4 and rel_err(w, w_ref, "room0_1000.weight") <= 1e-4
    gtd = gt.to(DEV)
    m = gtd > 0
    (torch.abs(gtd[m] - d[m]).sum() + 0.2 * torch.abs(gt_color.to(DEV) - c).sum()).backward()
    for k in grids:
        assert rel_err(grids[k].grad, grids_cpu[k].grad, "room0_1000.grad." + k) <= 1e-3, k
    named = dict(model.named_parameters())
    for k, v in params.items():
        if v.grad is not None and float(v.grad.abs().max()) > 0:
            assert rel_err(named[k].grad, v.grad, "room0_1000.grad.decoders(max over tensors)") <= 1e-3, k


def test_properties_at_full_size():
    """Size-independent properties on a 65 536-ray batch (too large for the oracle to be quick):
    permutation equivariance, attention weight in [0,1], non-negative variance, depth within [0, far]."""
    import attentive_dfprior_b200 as A
    from attentive_dfprior_b200 import synth
    from attentive_dfprior_b200.config import default_cfg, slam_namespace
    torch.manual_seed(0)
    scene = synth.make_scene("room0", seed=0, device=DEV, high_std=1e-2, channels_last=True)
    model = A.DF().to(DEV)
    renderer = A.Renderer(default_cfg(), None, slam_namespace(scene.bound, scene.tsdf_bnds))
    N = 65536
    rays_o, rays_d, gt, _ = synth.make_rays(scene, N, seed=3, device=DEV)
    with torch.no_grad():
        d, u, c, w = renderer.render_batch_ray(scene.grids, model, rays_d, rays_o, DEV, scene.tsdf_volume, scene.tsdf_bnds,
                                               "color", gt_depth=gt)
        perm = torch.randperm(N, device=DEV)
        d2, u2, c2, w2 = renderer.render_batch_ray(scene.grids, model, rays_d[perm], rays_o[perm], DEV, scene.tsdf_volume,
                                                   scene.tsdf_bnds, "color", gt_depth=gt[perm])
    assert torch.equal(d[perm], d2) and torch.equal(u[perm], u2) and torch.equal(c[perm], c2) and torch.equal(w[perm], w2)
    assert float(w.min()) >= 0 and float(w.max()) <= 1
    assert float(u.min()) >= 0
    assert float(d.min()) >= 0 and float(d.max()) <= float(gt.max()) * 1.2 + 1e-6
    assert torch.isfinite(d).all() and torch.isfinite(c).all()


def test_decoder_module_semantics():
    """deepcopy / state_dict round trip / in-place parameter updates (src/Tracker.py:144, Logger.py:26), each result
    compared with the oracle evaluated on the corresponding state_dict."""
    fx = load_golden("map_high_s1")
    inp = fx["inputs"]
    model, renderer, grids, tsdf, tb = build(inp)
    p_cpu = _random_points(inp["bound"], 500, 3, torch.float64)
    p = p_cpu.to(DEV)

    def oracle(sd):
        raw, _ = O.eval_points({k: v.detach().cpu() for k, v in sd.items()}, p_cpu, inp["grids"], inp["tsdf_volume"],
                               inp["tsdf_bnds"], inp["bound"], "high")
        return raw

    def close(x, ref):
        scale = float(ref[:, 3][ref[:, 3] != 100].abs().max())
        return float((x.cpu() - ref).abs().max()) <= 1e-4 * max(scale, 1.0)

    with torch.no_grad():
        ref_a = oracle(model.state_dict())
        a, _ = renderer.eval_points(p, model, tsdf, tb, grids, "high", DEV)
        assert close(a, ref_a)
        m2 = copy.deepcopy(model)
        b, _ = renderer.eval_points(p, m2, tsdf, tb, grids, "high", DEV)
        assert torch.equal(a, b)
        model.low_decoder.output_linear.bias.add_(0.25)          # an in-place update is picked up by the next call
        c, _ = renderer.eval_points(p, model, tsdf, tb, grids, "high", DEV)
        assert not torch.equal(a, c)
        assert close(c, oracle(model.state_dict()))
        assert not close(c, ref_a)
        m3 = type(model)().to(DEV)
        m3.load_state_dict(m2.state_dict())
        d, _ = renderer.eval_points(p, m3, tsdf, tb, grids, "high", DEV)
        assert torch.equal(a, d)


def test_errors_are_loud():
    import attentive_dfprior_b200 as A
    from attentive_dfprior_b200 import _lib as L
    fx = load_golden("map_low_s0")
    inp = fx["inputs"]
    model, renderer, grids, tsdf, tb = build(inp)
    with pytest.raises(L.DfpError):
        renderer.render_batch_ray({k: v.cpu() for k, v in grids.items()}, model, inp["rays_d"], inp["rays_o"], "cpu",
                                  inp["tsdf_volume"], inp["tsdf_bnds"], "low", gt_depth=inp["gt_depth"])
    with pytest.raises(ValueError):
        renderer.render_batch_ray(grids, model, inp["rays_d"].to(DEV)[:0], inp["rays_o"].to(DEV)[:0], DEV, tsdf, tb, "low",
                                  gt_depth=inp["gt_depth"].to(DEV)[:0])


def test_smoke_case_64_rays():
    """The driver's smoke() case as a test: 64 rays of the small scene, run here AFTER other scenes have been rendered in
    the same process (round 1 failed exactly this: an address-keyed bounds cache handed the small scene room0's box)."""
    import __graft_entry__ as G
    G.smoke()
