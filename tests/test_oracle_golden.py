"""CPU: the oracle restatement reproduces the reference's own outputs (tests/golden/*.pt).

The fixtures were produced by oracle/make_golden.py running the unmodified reference;
the oracle uses the same ATen ops in the same order, so on one thread it is bit-exact.
A small absolute slack is allowed for multi-threaded reductions only.
"""
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden
from oracle import dfprior_oracle as O


def _maxdiff(a, b):
    return float((a.double() - b.double()).abs().max())


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


def test_golden_cases_present():
    assert len(GOLDEN_CASES) >= 5


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_reference_fixture(name):
    fx = load_golden(name)
    i, o = fx["inputs"], fx["outputs"]
    params = {k: v.clone().requires_grad_(True) for k, v in i["state_dict"].items()}
    grids = {k: v.clone().requires_grad_(True) for k, v in i["grids"].items()}
    if fx["mode"] == "tracker":
        cam = i["cam"]
        ct = i["cam_tensor"].clone().requires_grad_(True)
        c2w = O.get_camera_from_tensor(ct)
        rays_o, rays_d = O.get_rays_from_uv(i["pix_i"], i["pix_j"], c2w, cam["fx"], cam["fy"], cam["cx"], cam["cy"])
        rays_o.retain_grad(); rays_d.retain_grad()
        assert _maxdiff(rays_o, i["rays_o"]) == 0 and _maxdiff(rays_d, i["rays_d"]) == 0
    else:
        rays_o, rays_d = i["rays_o"], i["rays_d"]
    (depth, unc, color, w), aux = O.render_batch_ray(
        params, grids, rays_d, rays_o, i["tsdf_volume"], i["tsdf_bnds"], i["bound"], fx["stage"],
        i["gt_depth"], return_aux=True)

    # sample depths and points are pure f64 arithmetic: bit-exact
    assert torch.equal(aux["z"], o["z"])
    assert torch.equal(aux["pts"].reshape(-1, 3), o["pts"])
    assert _maxdiff(aux["raw"].reshape(-1, 4), o["raw"]) <= 1e-6
    assert _maxdiff(depth, o["depth"]) <= 1e-9
    assert _maxdiff(unc, o["uncertainty"]) <= 1e-9
    assert _maxdiff(color, o["color"]) <= 1e-6
    assert _maxdiff(w, o["weight"]) <= 1e-6
    assert depth.dtype == torch.float64 and unc.dtype == torch.float64
    assert color.dtype == torch.float32 and w.shape == (rays_o.shape[0], 48, 1)

    if fx["mode"] == "mapper":
        loss = O.mapper_loss(depth, color, w, i["gt_depth"], i["gt_color"], fx["stage"], attn_term=True)
        loss.backward()
        assert abs(float(loss) - float(o["loss"])) <= 1e-6 * max(1.0, abs(float(o["loss"])))
        for k, g in fx["grads"]["grids"].items():
            if g is None:
                assert grids[k].grad is None or float(grids[k].grad.abs().max()) == 0
            else:
                assert _maxdiff(grids[k].grad, g) <= 1e-5 * max(1.0, float(g.abs().max()))
        n_checked = 0
        for k, g in fx["grads"]["params"].items():
            if g is None:
                assert params[k].grad is None or float(params[k].grad.abs().max()) == 0
            else:
                assert _maxdiff(params[k].grad, g) <= 1e-5 * max(1.0, float(g.abs().max())), k
                n_checked += 1
        assert n_checked >= 23
    else:
        loss = O.tracker_loss(depth, unc, color, i["gt_depth"], i["gt_color"])
        loss.backward()
        assert abs(float(loss) - float(o["loss"])) <= 1e-6 * abs(float(o["loss"]))
        for key, t in (("cam_tensor", ct), ("rays_o", rays_o), ("rays_d", rays_d)):
            g = fx["grads"][key]
            assert _maxdiff(t.grad, g) <= 1e-5 * max(1.0, float(g.abs().max())), key


def test_state_dict_names_and_count():
    """79 tensors, 86 029 parameters, reference names."""
    fx = load_golden(GOLDEN_CASES[0])
    sd = fx["inputs"]["state_dict"]
    mine = O.init_decoder_params(0)
    assert list(sd.keys()) == list(mine.keys())
    assert len(sd) == 79
    assert sum(v.numel() for v in sd.values()) == 86029
    for k in sd:
        assert sd[k].shape == mine[k].shape, k


def test_out_of_bound_sample_is_opaque():
    """Out-of-bound points get occ=100 -> alpha=1 (Renderer.py:64, common.py:236)."""
    fx = load_golden("map_low_s0")
    i = fx["inputs"]
    p = torch.tensor([[50.0, 0.0, 0.0], [0.0, 0.0, 0.0]], dtype=torch.float64)
    raw, w = O.eval_points(i["state_dict"], p, i["grids"], i["tsdf_volume"], i["tsdf_bnds"], i["bound"], "low")
    assert float(raw[0, 3]) == 100.0 and float(raw[1, 3]) != 100.0
    assert float(torch.sigmoid(10 * raw[0, 3])) == 1.0


def test_empty_band_gives_unit_weight_and_low_only():
    """No point inside the TSDF band => w == 1 and occ == low decoder only (decoder.py:332-336)."""
    fx = load_golden("map_high_s1")
    i = fx["inputs"]
    tsdf = torch.ones_like(i["tsdf_volume"])
    g = torch.Generator().manual_seed(0)
    b = i["bound"]
    p = (torch.rand(64, 3, generator=g, dtype=torch.float64) * (b[:, 1] - b[:, 0]) * 0.9 + b[:, 0])
    raw_h, w = O.df_forward(i["state_dict"], p, i["grids"], tsdf, i["tsdf_bnds"], b, "high")
    raw_l, _ = O.df_forward(i["state_dict"], p, i["grids"], tsdf, i["tsdf_bnds"], b, "low")
    assert torch.all(w == 1)
    assert torch.equal(raw_h[:, 3], raw_l[:, 3])


def test_composite_weights_sum_at_most_one():
    g = torch.Generator().manual_seed(1)
    raw = torch.randn(7, 48, 4, generator=g)
    z = torch.sort(torch.rand(7, 48, generator=g, dtype=torch.float64), -1)[0]
    depth, var, rgb, wts = O.composite(raw, z)
    assert float(wts.sum(-1).max()) <= 1.0 + 1e-5
    assert torch.all(var >= 0)
