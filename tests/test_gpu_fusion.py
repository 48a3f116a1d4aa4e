"""GPU: the TSDF integrate kernel (dfp_tsdf_integrate) against the oracle bit for bit and against the reference's CPU mode,
and Mesher.point_masks (dfp_point_masks) against the reference's own method."""
import numpy as np
import pytest
import torch

from oracle import fusion_oracle as FO
from test_fusion_oracle import load, run_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_tsdf_integrate_matches_oracle_and_reference():
    from attentive_dfprior_b200.fusion import TSDFVolume
    fx = load()
    i, o = fx["inputs"], fx["outputs"]
    vol = TSDFVolume(i["vol_bnds"].numpy().copy(), voxel_size=i["voxel_size"], device=DEV)
    for c, d, p in i["frames"]:
        vol.integrate(c.numpy(), d.numpy(), i["K"].numpy(), p.numpy(), obs_weight=1.)
    tsdf, color, bnds = vol.get_volume()
    t_o, w_o, c_o, b_o, dim = run_oracle(fx)
    assert np.array_equal(vol._vol_dim, dim) and np.allclose(bnds, b_o)
    assert np.array_equal(vol._weight_vol.cpu().numpy(), w_o), "same voxels observed the same number of times"
    assert np.array_equal(tsdf, t_o) and np.array_equal(color, c_o), "kernel and oracle share the f32 arithmetic: bit-exact"
    same_w = vol._weight_vol.cpu().numpy() == o["weight"].numpy()
    assert same_w.mean() >= 0.995
    assert ((np.abs(tsdf - o["tsdf"].numpy()) <= 1e-4) | ~same_w).all()
    prior, tb = vol.tsdf_prior()                                   # get_tsdf.py:95-97 layout, consumed by the renderer
    assert prior.shape == (1, 1, dim[2], dim[1], dim[0]) and not prior.is_contiguous()
    assert float(prior[0, 0, 3, 2, 1]) == float(tsdf[1, 2, 3]) and tb.shape == (3, 2)


def test_integrate_room0_volume_runs_at_full_size():
    """The 758 x 574 x 451 room0 volume (196 M voxels): one frame integrates; observed voxels lie in the frustum."""
    from attentive_dfprior_b200 import synth
    from attentive_dfprior_b200.fusion import TSDFVolume
    b = synth.round_bound(synth.ROOM0_BOUND).numpy()
    vol = TSDFVolume(b, voxel_size=4.0 / 256, device=DEV)
    assert tuple(vol._vol_dim) == (758, 574, 451)
    H, W = 680, 1200
    K = np.array([[600.0, 0, 599.5], [0, 600.0, 339.5], [0, 0, 1]])
    pose = np.eye(4); pose[:3, 3] = [3.0, 1.0, 0.0]
    depth = np.full((H, W), 2.5, dtype=np.float32)
    color = np.full((H, W, 3), 128, dtype=np.uint8)
    vol.integrate(color, depth, K, pose)
    w = vol._weight_vol
    n_obs = int((w > 0).sum())
    assert 1_000_000 < n_obs < 60_000_000
    t = vol._tsdf_vol[w > 0]
    assert float(t.max()) <= 1.0 and float(t.min()) >= -1.0
    assert float(vol._color_vol[w > 0].max()) == 128 * 65536 + 128 * 256 + 128


def test_point_masks_match_reference():
    import attentive_dfprior_b200 as A
    from attentive_dfprior_b200.config import default_cfg, slam_namespace
    fx = load()["point_masks"]
    cam = fx["cam"]
    bound = torch.tensor([[-3.0, 3.0]] * 3, dtype=torch.float64)
    mesher = A.Mesher(default_cfg(), None, slam_namespace(bound, bound.to(DEV), cam), points_batch_size=3000)
    kfs = [{"est_c2w": k["est_c2w"].to(DEV), "depth": k["depth"].to(DEV)} for k in fx["keyframes"]]
    seen, fore, unseen = mesher.point_masks(fx["points"], kfs, fx["est"].to(DEV), 3, DEV)
    agree = (seen == fx["seen"].numpy()) & (fore == fx["forecast"].numpy()) & (unseen == fx["unseen"].numpy())
    assert agree.mean() >= 0.9995, agree.mean()                   # a point within an ulp of a frustum edge may land on either side
    assert seen.dtype == np.bool_ and (seen | fore | unseen).all() and not (seen & fore).any()
    seen, fore, unseen = mesher.point_masks(fx["points"], kfs, fx["est"].to(DEV), 3, DEV, get_mask_use_all_frames=True)
    agree = (seen == fx["seen_all"].numpy()) & (fore == fx["forecast_all"].numpy())
    assert agree.mean() >= 0.9995, agree.mean()
