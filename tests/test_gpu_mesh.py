"""GPU: the mesh extraction that follows the Mesher's dense query (include/dfprior_mesh.h, attentive_dfprior_b200/mesher.py)
against its numpy restatement (oracle/mesh_oracle.py): marching cubes bit-exact (vertex positions, vertex order, faces), culling
and component filter, the PLY writer, and Mesher.get_mesh end to end on the small synthetic scene."""
import os
import struct

import numpy as np
import pytest
import torch

from oracle import mesh_oracle as M
from test_gpu_parity import DEV

pytestmark = pytest.mark.gpu


def fields():
    ax = [np.linspace(-1.2, 1.2, n) for n in (21, 17, 26)]                   # three different sizes and spacings
    g = np.stack(np.meshgrid(*ax, indexing="ij"), -1)
    sp = tuple(float(a[1] - a[0]) for a in ax)
    org = (-1.2, -1.2, -1.2)
    yield "sphere", (np.linalg.norm(g, axis=-1) - 0.8).astype(np.float32), 0.0, sp, org
    r = np.sqrt(g[..., 0] ** 2 + g[..., 1] ** 2)
    yield "torus_level", (np.sqrt((r - 0.7) ** 2 + g[..., 2] ** 2)).astype(np.float32), 0.25, sp, org
    rng = np.random.default_rng(5)
    rough = rng.standard_normal((14, 11, 13)).astype(np.float32)             # ambiguous faces, multi-polygon cells
    yield "rough", rough, 0.1, (0.5, 0.25, 1.0), (3.0, -2.0, 0.5)
    occ = rng.uniform(0, 1, (9, 9, 9)).astype(np.float32)
    occ[3:6, 3:6, 3:6] = 100.0                                               # the Mesher's "outside the bound" value (Mesher.py:448)
    yield "occupancy_with_100", occ, 0.5, (1.0, 1.0, 1.0), (0.0, 0.0, 0.0)


@pytest.mark.parametrize("case", list(fields()), ids=lambda c: c[0])
def test_marching_cubes_matches_oracle_bit_exactly(case):
    from attentive_dfprior_b200 import mesher as MS
    _, vol, level, sp, org = case
    tri, key = M.marching_cubes(vol, level, sp, org)
    v_ref, f_ref = M.weld(tri, key)
    v, f = MS.marching_cubes(torch.from_numpy(vol).to(DEV), level, sp, org)
    assert v.dtype == torch.float32 and f.dtype == torch.int64
    assert tuple(f.shape) == f_ref.shape and tuple(v.shape) == v_ref.shape
    assert np.array_equal(f.cpu().numpy(), f_ref)
    assert np.array_equal(v.cpu().numpy(), v_ref), float(np.abs(v.cpu().numpy() - v_ref).max())


def test_empty_and_degenerate_volumes():
    from attentive_dfprior_b200 import mesher as MS
    v, f = MS.marching_cubes(torch.ones(8, 8, 8, device=DEV), 0.0)
    assert v.shape == (0, 3) and f.shape == (0, 3)
    with pytest.raises(Exception):
        MS.marching_cubes(torch.ones(1, 8, 8, device=DEV), 0.0)
    with pytest.raises(ValueError):
        MS.marching_cubes(torch.ones(8, 8, 8), 0.0)


def test_culling_components_and_ply(tmp_path):
    from attentive_dfprior_b200 import mesher as MS
    ax = np.linspace(-1.2, 1.2, 26)
    g = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1)
    two = np.minimum(np.linalg.norm(g - np.array([0.55, 0, 0]), axis=-1) - 0.4, np.linalg.norm(g + np.array([0.6, 0, 0]), axis=-1) - 0.25).astype(np.float32)
    sp, org = (float(ax[1] - ax[0]),) * 3, (-1.2,) * 3
    v, f = MS.marching_cubes(torch.from_numpy(two).to(DEV), 0.0, sp, org)
    v_ref, f_ref = M.weld(*M.marching_cubes(two, 0.0, sp, org))
    seen = v_ref[:, 0] > 0
    kept = MS.cull_unseen_faces(f, seen)
    assert np.array_equal(kept.cpu().numpy(), M.cull_unseen(f_ref, seen))
    lab, area = M.components(v_ref, f_ref)
    v1, f1 = MS.filter_components(v, f, largest=True)
    assert f1.shape[0] == int((lab == np.argmax(area)).sum()) and v1.shape[0] == len(np.unique(f_ref[lab == np.argmax(area)]))
    v2, f2 = MS.filter_components(v, f, min_area=float(area.min()) * 0.5)
    assert f2.shape[0] == f.shape[0]
    v3, f3 = MS.filter_components(v, f, min_area=float(area.min()) * 1.5)
    assert f3.shape[0] == f1.shape[0] and torch.equal(v3[f3], v1[f1])
    path = os.path.join(tmp_path, "m.ply")
    col = torch.randint(0, 255, (v1.shape[0], 3), dtype=torch.uint8)
    MS.write_ply(path, v1, f1, col)
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    assert f"element vertex {v1.shape[0]}".encode() in head and f"element face {f1.shape[0]}".encode() in head
    assert len(body) == v1.shape[0] * 15 + f1.shape[0] * 13
    x, y, z, r_, g_, b_ = struct.unpack_from("<fffBBB", body, 0)
    assert (x, y, z) == tuple(float(t) for t in v1[0].cpu()) and (r_, g_, b_) == tuple(int(t) for t in col[0])


def test_get_mesh_end_to_end(tmp_path):
    """Mesher.get_mesh on the small synthetic scene: query -> marching cubes -> (culling against two keyframes) -> colours ->
    PLY.  The surface must be the level set of the very occupancy lattice query_occupancy returns (oracle on that lattice)."""
    import attentive_dfprior_b200 as A
    from attentive_dfprior_b200 import synth
    from attentive_dfprior_b200.config import default_cfg, slam_namespace
    scene = synth.make_scene("small", seed=3, device=DEV, high_std=1e-2, channels_last=True)
    torch.manual_seed(3)
    model = A.DF().to(DEV)
    cfg = default_cfg()
    cfg["meshing"]["resolution"] = 40
    cfg["meshing"]["level_set"] = 0.0
    mesher = A.Mesher(cfg, None, slam_namespace(scene.bound, scene.tsdf_bnds, scene.cam), points_batch_size=20000)
    grids = scene.grids
    out = mesher.get_mesh(None, grids, model, [], None, 0, scene.tsdf_volume, DEV, color=True, clean_mesh=False)
    grid = mesher.get_grid_uniform(40)
    z = mesher.query_occupancy(grids, model, scene.tsdf_volume, DEV, points=grid["grid_points"])
    ax = grid["xyz"]
    vol = z.reshape(ax[1].shape[0], ax[0].shape[0], ax[2].shape[0]).permute(1, 0, 2).contiguous().cpu().numpy()
    sp = tuple(float(ax[k][2] - ax[k][1]) for k in range(3))
    org = tuple(float(ax[k][0]) for k in range(3))
    v_ref, f_ref = M.weld(*M.marching_cubes(vol, 0.0, sp, org))
    if f_ref.shape[0] == 0:
        assert out is None
        pytest.skip("random-init decoders gave no zero crossing on this lattice")
    verts, faces, colors = out
    assert np.array_equal(faces.cpu().numpy(), f_ref) and np.array_equal(verts.cpu().numpy(), v_ref)
    assert colors.dtype == torch.uint8 and tuple(colors.shape) == (verts.shape[0], 3)
    # with keyframes: faces whose vertices no keyframe sees go, small components go, and a file is written
    kfs = []
    for s in (1, 2):
        c2w = synth.make_pose(scene, s).to(DEV)
        kfs.append({"est_c2w": c2w, "depth": torch.full((scene.cam["H"], scene.cam["W"]), 3.0, device=DEV)})
    path = os.path.join(tmp_path, "scene.ply")
    out2 = mesher.get_mesh(path, grids, model, kfs, None, 0, scene.tsdf_volume, DEV, color=True, clean_mesh=True,
                           remove_small_geometry_threshold=0.0)
    if out2 is not None:
        v2, f2, c2 = out2
        assert f2.shape[0] <= faces.shape[0] and int(f2.max()) < v2.shape[0] and c2.shape[0] == v2.shape[0]
        assert os.path.getsize(path) > 0
