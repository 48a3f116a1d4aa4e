"""CPU: the kernels' per-thread arithmetic (dfp_math.h / dfp_layout.h), run serially by the
host emulation in tests/hostemu, against the reference fixtures and the oracle.

This is a unit test of the *formulas and layouts* the sm_100a kernels use; it is not a
product path (the package has no CPU fallback).  The kernels themselves are checked by
the `-m gpu` tests.
"""
import ctypes as C
import os
import subprocess

import pytest
import torch

from conftest import GOLDEN_CASES, REPO, load_golden
from attentive_dfprior_b200 import _lib as L
from oracle import dfprior_oracle as O

SRC = os.path.join(REPO, "tests", "hostemu", "hostemu.cpp")
SO = os.path.join(REPO, "tests", "hostemu", "_hostemu.so")


@pytest.fixture(scope="module")
def emu():
    deps = [SRC, os.path.join(REPO, "attentive_dfprior_b200", "csrc", "dfp_math.h"),
            os.path.join(REPO, "attentive_dfprior_b200", "csrc", "dfp_layout.h"),
            os.path.join(REPO, "include", "dfprior_b200.h")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", SO, SRC])
    lib = C.CDLL(SO)
    lib.emu_render.restype = C.c_int
    lib.emu_render.argtypes = [C.POINTER(L.DfpScene), C.c_void_p, C.POINTER(L.DfpRays), C.POINTER(L.DfpCfg),
                               C.POINTER(L.DfpRenderOut), C.POINTER(L.DfpRenderOutGrad), C.POINTER(L.DfpGrads)]
    return lib


def ndhwc(g):   # [1,C,D,H,W] -> contiguous [D,H,W,C]
    return g[0].permute(1, 2, 3, 0).contiguous()


def make_cfg(stage, n_samples=32, n_surface=16):
    cfg = L.DfpCfg()
    cfg.stage = L.STAGES[stage]
    cfg.n_samples, cfg.n_surface, cfg.n_importance = n_samples, n_surface, 0
    tv = torch.linspace(0., 1., steps=n_samples)
    omt = 1. - tv
    ts = torch.linspace(0., 1., steps=n_surface).double()
    omts = 1. - ts
    for i in range(n_samples):
        cfg.t_vals[i] = float(tv[i]); cfg.one_minus_t[i] = float(omt[i])
    for i in range(n_surface):
        cfg.t_surf[i] = float(ts[i]); cfg.one_minus_t_surf[i] = float(omts[i])
    return cfg


def make_scene_struct(inp, grids_cl):
    sc = L.DfpScene()
    for name in ("low", "high", "color"):
        g = grids_cl[name]
        s = getattr(sc, name)
        s.data = g.data_ptr(); s.D, s.H, s.W, s.C = g.shape
    t = inp["tsdf_volume"]
    sc.tsdf.data = t.data_ptr()
    sc.tsdf.D, sc.tsdf.H, sc.tsdf.W = t.shape[2:]
    sc.tsdf.sD, sc.tsdf.sH, sc.tsdf.sW = t.stride()[2:]
    L.fill_bounds(sc.bound, inp["bound"])
    L.fill_bounds(sc.tsdf_bnds, inp["tsdf_bnds"])
    return sc


def pack(emu, sd):
    tensors = [v.contiguous().float() for v in sd.values()]
    arr = (C.c_void_p * 79)(*[t.data_ptr() for t in tensors])
    packed = torch.zeros(emu.emu_packed_floats())
    emu.emu_pack(C.cast(arr, C.POINTER(C.c_void_p)), C.c_void_p(packed.data_ptr()))
    return packed


def rel_err(a, b):
    """max |a-b| normalised by max(|b|_max, rms-ish floor)"""
    a, b = a.double(), b.double()
    scale = max(float(b.abs().max()), 1e-30)
    return float((a - b).abs().max()) / scale


def test_pack_roundtrip(emu):
    sd = O.init_decoder_params(3)
    packed = pack(emu, sd)
    flat = torch.zeros(L.N_PARAMS)
    emu.emu_unpack(C.c_void_p(packed.data_ptr()), C.c_void_p(flat.data_ptr()))
    ref = torch.cat([v.reshape(-1) for v in sd.values()])
    assert torch.equal(flat, ref)
    # every packed slot is hit at most once: packing ones then summing counts all parameters exactly
    ones = {k: torch.ones_like(v) for k, v in sd.items()}
    assert float(pack(emu, ones).sum()) == L.N_PARAMS


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_hostemu_matches_reference(emu, name):
    fx = load_golden(name)
    inp, ref = fx["inputs"], fx["outputs"]
    stage = fx["stage"]
    grids_cl = {k[5:]: ndhwc(v) for k, v in inp["grids"].items()}
    sc = make_scene_struct(inp, grids_cl)
    packed = pack(emu, inp["state_dict"])
    cfg = make_cfg(stage)
    ro, rd, gt = inp["rays_o"].contiguous(), inp["rays_d"].contiguous(), inp["gt_depth"].contiguous()
    N, S = ro.shape[0], 48
    rays = L.DfpRays(ro.data_ptr(), rd.data_ptr(), gt.data_ptr(), N, 0)
    depth = torch.zeros(N, dtype=torch.float64); unc = torch.zeros(N, dtype=torch.float64)
    color = torch.zeros(N, 3); weight = torch.zeros(N, S); z = torch.zeros(N, S, dtype=torch.float64)
    raw = torch.zeros(N, S, 4)
    out = L.DfpRenderOut(depth.data_ptr(), unc.data_ptr(), color.data_ptr(), weight.data_ptr(), z.data_ptr(), raw.data_ptr())

    # upstream gradients of the fixture's loss, taken from the oracle's autograd
    params = {k: v.clone().requires_grad_(True) for k, v in inp["state_dict"].items()}
    grids = {k: v.clone().requires_grad_(True) for k, v in inp["grids"].items()}
    ro_t, rd_t = ro.clone().requires_grad_(True), rd.clone().requires_grad_(True)
    torch.set_num_threads(1)
    d_o, u_o, c_o, w_o = O.render_batch_ray(params, grids, rd_t, ro_t, inp["tsdf_volume"], inp["tsdf_bnds"], inp["bound"],
                                            stage, gt)
    for t in (d_o, u_o, c_o, w_o):
        if t.requires_grad:
            t.retain_grad()
    if fx["mode"] == "mapper":
        loss = O.mapper_loss(d_o, c_o, w_o, gt, inp["gt_color"], stage, attn_term=True)
    else:
        # let the uncertainty carry gradient too so that its backward formula is exercised
        loss = O.tracker_loss(d_o, u_o, c_o, gt, inp["gt_color"]) + 0.3 * u_o.sum()
    loss.backward()
    zero = lambda t: torch.zeros_like(t)
    g_depth = d_o.grad if d_o.grad is not None else zero(d_o)
    g_unc = u_o.grad if u_o.grad is not None else zero(u_o)
    g_col = c_o.grad if c_o.grad is not None else zero(c_o)
    g_w = (w_o.grad if w_o.grad is not None else zero(w_o)).reshape(N, S).contiguous()
    og = L.DfpRenderOutGrad(g_depth.data_ptr(), g_unc.data_ptr(), g_col.contiguous().data_ptr(), g_w.data_ptr())
    gg = {k: torch.zeros_like(v) for k, v in grids_cl.items()}
    gp = torch.zeros_like(packed)
    g_ro, g_rd = torch.zeros(N, 3), torch.zeros(N, 3)
    grads = L.DfpGrads(gg["low"].data_ptr(), gg["high"].data_ptr(), gg["color"].data_ptr(), gp.data_ptr(),
                       g_ro.data_ptr(), g_rd.data_ptr(), None)
    rc = emu.emu_render(C.byref(sc), C.c_void_p(packed.data_ptr()), C.byref(rays), C.byref(cfg), C.byref(out),
                        C.byref(og), C.byref(grads))
    assert rc == 0

    # ---- forward: sampling chain bit-exact, floats within the north-star tolerances
    assert torch.equal(z, ref["z"]), "sample depths must be bit-exact"
    raw_ref = ref["raw"].reshape(N, S, 4)
    rms = float(raw_ref[..., 3][raw_ref[..., 3] != 100].pow(2).mean().sqrt())
    assert float((raw - raw_ref).abs().max()) <= 1e-4 * max(rms, 1.0) + 2e-4 * rms
    assert torch.equal(raw[..., 3] == 100, raw_ref[..., 3] == 100), "in-bound mask must be exact"
    assert torch.equal(weight != 1, ref["weight"].reshape(N, S) != 1), "band mask must be exact"
    assert rel_err(depth, ref["depth"]) <= 1e-4
    assert rel_err(unc, ref["uncertainty"]) <= 1e-4
    assert rel_err(color, ref["color"]) <= 1e-4 or stage != "color"
    assert rel_err(weight, ref["weight"].reshape(N, S)) <= 1e-4

    # ---- backward vs the oracle's autograd (1e-3 of the tensor's max)
    flat = torch.zeros(L.N_PARAMS)
    emu.emu_unpack(C.c_void_p(gp.data_ptr()), C.c_void_p(flat.data_ptr()))
    off = 0
    worst = ("", 0.0)
    for k, v in params.items():
        n = v.numel()
        mine = flat[off:off + n].reshape(v.shape)
        off += n
        if v.grad is None or float(v.grad.abs().max()) == 0:
            assert float(mine.abs().max()) == 0, k
            continue
        e = rel_err(mine, v.grad)
        if e > worst[1]:
            worst = (k, e)
    assert worst[1] <= 1e-3, worst
    for k, v in grids.items():
        mine = gg[k[5:]].permute(3, 0, 1, 2).unsqueeze(0)
        if v.grad is None:
            assert float(mine.abs().max()) == 0
        else:
            assert rel_err(mine, v.grad) <= 1e-3, k
    assert rel_err(g_ro, ro_t.grad) <= 1e-3
    assert rel_err(g_rd, rd_t.grad) <= 1e-3


def test_merged_launch_grid_split(emu):
    """The split every CTA of a merged decoder launch derives for itself (dfp_layout.h split_ctas): the decoders' CTA ranges
    tile the front of the grid without overlap, a decoder with work gets at least one CTA and never more CTAs than it has
    iterations, a decoder without work gets none, and the grid is used up whenever there is that much work."""
    import itertools
    import random
    rng = random.Random(0)
    cases = [(a, b, c, g) for a, b, c in itertools.product((0, 1, 2, 5, 40), repeat=3) for g in (1, 2, 3, 7, 148)]
    cases += [(rng.randrange(0, 30000), rng.randrange(0, 30000), rng.randrange(0, 30000), 148) for _ in range(300)]
    cases += [(rng.randrange(0, 4), rng.randrange(0, 300), rng.randrange(0, 300), rng.randrange(1, 149)) for _ in range(300)]
    for a, b, c, g in cases:
        it = (C.c_int64 * 3)(a, b, c)
        active = sum(1 for v in (a, b, c) if v > 0)
        if g < active:
            continue                      # the launchers size the grid as min(SMs, total iterations) >= active decoders
        first, count = (C.c_int * 3)(), (C.c_int * 3)()
        emu.emu_split_ctas(it, g, first, count)
        at = 0
        for d, v in enumerate((a, b, c)):
            assert first[d] == at, (a, b, c, g, list(first), list(count))
            assert (count[d] == 0) == (v == 0), (a, b, c, g, list(count))
            assert 0 <= count[d] <= v
            at += count[d]
        assert at <= g
        if a + b + c >= g:
            assert at == g or at >= g - 2, (a, b, c, g, list(count))       # spare CTAs only when the largest decoder is saturated
