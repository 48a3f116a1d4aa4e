"""Diagnostic (GPU box): TC backward of the point query with grid gradients but no point gradient, various n / stages."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch
from oracle import dfprior_oracle as O
from test_gpu_parity import DEV, build, rel_err
from test_oracle_extra import load_extra
from attentive_dfprior_b200 import functional as Fn

fx = load_extra("extra_unc_high")
inp = fx["inputs"]
sd = {k: v.clone() for k, v in inp["state_dict"].items()}
for stage, n, want_p in (("high", 400, False), ("high", 384, False), ("low", 400, False), ("low", 130, False), ("low", 128, False), ("high", 400, True), ("low", 3000, False)):
    p = torch.rand(n, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(2))
    p = (p * (inp["bound"][:, 1] - inp["bound"][:, 0]) + inp["bound"][:, 0])
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    gcpu = {k: v.clone().requires_grad_(True) for k, v in inp["grids"].items()}
    r2, _ = O.eval_points(params, p, gcpu, inp["tsdf_volume"], inp["tsdf_bnds"], inp["bound"], stage)
    r2[:, 3].sum().backward()
    model, renderer, grids, tsdf, tb = build(inp)
    g1 = {k: v.clone().requires_grad_(True) for k, v in grids.items()}
    pg = p.to(DEV).requires_grad_(want_p)
    raw, _ = Fn.eval_points(pg, model, g1, tsdf, tb, inp["bound"], stage, apply_bound_mask=True)
    raw[:, 3].sum().backward()
    named = dict(model.named_parameters())
    worst = max((rel_err(named[k].grad, v.grad), k) for k, v in params.items() if v.grad is not None and float(v.grad.abs().max()) > 0)
    gl, rl = g1["grid_low"].grad.cpu(), gcpu["grid_low"].grad
    print(stage, n, want_p, "fwd %.1e grid_low %.1e worst param %.1e %s" % (rel_err(raw, r2), rel_err(gl, rl), worst[0], worst[1]), flush=True)
    d = (gl - rl).abs()
    mx = float(rl.abs().max())
    bad = (d > 1e-3 * mx).nonzero()
    if len(bad):
        vox = sorted(set((int(b[2]), int(b[3]), int(b[4])) for b in bad))
        print("   bad voxels (z,y,x):", vox[:12], "n bad elements", len(bad))
        # which samples touch those voxels?  (low grid, trilinear cell of each point)
        from attentive_dfprior_b200 import synth
        b = inp["bound"]; D, H, W = rl.shape[2:]
        xn = (p - b[:, 0]) / (b[:, 1] - b[:, 0])
        ix, iy, iz = xn[:, 0] * (W - 1), xn[:, 1] * (H - 1), xn[:, 2] * (D - 1)
        for (z, y, x) in vox[:4]:
            near = ((ix.floor() - x).abs() <= 1) & ((iy.floor() - y).abs() <= 1) & ((iz.floor() - z).abs() <= 1)
            print("   voxel", (z, y, x), "samples nearby:", near.nonzero().flatten().tolist()[:20], "ratio mine/ref at voxel ch0..3:",
                  (gl[0, :4, z, y, x] / rl[0, :4, z, y, x]).tolist())
