"""CPU: the oracle against the reference outputs of tests/golden/extra_*.pt (oracle/make_golden_extra.py):
gt_depth=None, uncertainty gradients, eval_points_tsdf / sample_grid_tsdf, the Mesher's lattice and point query."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import dfprior_oracle as O


def load_extra(name):
    fx = torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)
    inp = fx["inputs"]
    X, Y, Z = inp["tsdf_xyz"].shape
    inp["tsdf_volume"] = inp["tsdf_xyz"].reshape(1, 1, X, Y, Z).permute(0, 1, 4, 3, 2)
    return fx


def extra_loss(fx, depth, unc, color, weight):
    i = fx["inputs"]
    if fx["loss_kind"] == "nodepth":
        return 0.05 * depth.sum() + unc.sum() + 0.2 * torch.abs(i["gt_color"].to(color.device) - color).sum() \
            + torch.abs(weight - torch.ones_like(weight)).sum()
    gt = i["gt_depth"].to(depth.device)
    m = gt > 0
    return (torch.abs(gt - depth) / torch.sqrt(unc + 1e-10))[m].sum()


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


def _maxdiff(a, b):
    return float((a.detach().double() - b.detach().double()).abs().max())


@pytest.mark.parametrize("name", ["extra_nodepth_color", "extra_unc_high"])
def test_oracle_matches_extra_render_fixture(name):
    fx = load_extra(name)
    i, o = fx["inputs"], fx["outputs"]
    params = {k: v.clone().requires_grad_(True) for k, v in i["state_dict"].items()}
    grids = {k: v.clone().requires_grad_(True) for k, v in i["grids"].items()}
    gt = i["gt_depth"] if fx["with_depth"] else None
    (depth, unc, color, w), aux = O.render_batch_ray(params, grids, i["rays_d"], i["rays_o"], i["tsdf_volume"], i["tsdf_bnds"],
                                                     i["bound"], fx["stage"], gt, return_aux=True)
    S = 48 if fx["with_depth"] else 32                                    # Renderer.py:140-142: N_surface = 0 without depth
    assert aux["z"].shape[1] == S and torch.equal(aux["z"], o["z"])
    assert depth.dtype == torch.float64 and o["depth"].dtype == torch.float64      # the reference returns f64 here too
    assert _maxdiff(depth, o["depth"]) <= 1e-9 and _maxdiff(unc, o["uncertainty"]) <= 1e-9
    assert _maxdiff(color, o["color"]) <= 1e-6 and _maxdiff(w, o["weight"]) <= 1e-6
    loss = extra_loss(fx, depth, unc, color, w)
    loss.backward()
    assert abs(float(loss) - float(o["loss"])) <= 1e-6 * max(1.0, abs(float(o["loss"])))
    for k, g in fx["grads"]["grids"].items():
        if g is not None and float(g.abs().max()) > 0:
            assert _maxdiff(grids[k].grad, g) <= 1e-5 * float(g.abs().max()), k
    n = 0
    for k, g in fx["grads"]["params"].items():
        if g is not None and float(g.abs().max()) > 0:
            assert _maxdiff(params[k].grad, g) <= 1e-5 * max(1.0, float(g.abs().max())), k
            n += 1
    assert n >= 23


def test_oracle_tsdf_sampling_matches_reference():
    fx = load_extra("extra_tsdf_points")
    i, o = fx["inputs"], fx["outputs"]
    for tag in ("f64", "f32"):
        p = i["p64" if tag == "f64" else "p32"]
        v = O.trilinear_gather(i["tsdf_volume"], p, i["tsdf_bnds"])       # Renderer.py:73-81
        assert torch.equal(v.reshape(1, -1), o["vals_" + tag])
        assert torch.equal(v.reshape(1, 1, -1), o["grid_" + tag].reshape(1, 1, -1))


def test_oracle_mesher_lattice_and_query_match_reference():
    fx = load_extra("extra_mesher")
    i, o = fx["inputs"], fx["outputs"]
    pts = O.get_grid_uniform(i["marching_cubes_bound"], i["resolution"])
    assert torch.equal(pts, o["grid_points"])
    for st, ref in o["ret"].items():
        raw, _ = O.eval_points(i["state_dict"], pts, i["grids"], i["tsdf_volume"], i["tsdf_bnds"], i["bound"], st,
                               points_batch_size=700)
        assert _maxdiff(raw, ref) <= 1e-6, st
    big = O.get_grid_uniform(i["scannet0000_marching_cubes_bound"], 256)            # scannet0000 shapes (16.8 M points)
    assert big.shape[0] == o["scannet0000_n"] == 256 ** 3
    assert hashlib.sha256(big.numpy().tobytes()).hexdigest() == o["scannet0000_sha256"]
    assert torch.equal(big[o["scannet0000_rows"]], o["scannet0000_sample"])


def test_product_mesher_lattice_matches_reference():
    """attentive_dfprior_b200.Mesher.get_grid_uniform is host code: checked here without a GPU."""
    from types import SimpleNamespace
    from attentive_dfprior_b200.mesher import Mesher
    fx = load_extra("extra_mesher")
    i, o = fx["inputs"], fx["outputs"]
    me = Mesher.__new__(Mesher)
    me.marching_cubes_bound = i["marching_cubes_bound"]
    g = me.get_grid_uniform(i["resolution"])
    assert torch.equal(g["grid_points"], o["grid_points"])
    for a, b in zip(g["xyz"], o["xyz"]):
        assert np.array_equal(np.asarray(a, dtype=np.float64), b.numpy())
    me.marching_cubes_bound = i["scannet0000_marching_cubes_bound"]
    big = me.get_grid_uniform(256)["grid_points"]
    assert hashlib.sha256(big.numpy().tobytes()).hexdigest() == o["scannet0000_sha256"]
