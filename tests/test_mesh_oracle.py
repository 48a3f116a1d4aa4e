"""CPU: the generated marching-cubes tables (oracle/mc_tables.py) and the numpy restatement of the mesh extraction that follows the
Mesher query (oracle/mesh_oracle.py).  skimage / trimesh are not available here, so the triangulation cannot be pinned to the
reference's library ("parity unpinned" for it); what is checked are the properties every correct extraction has."""
import numpy as np
import pytest

from oracle import mesh_oracle as M
from oracle.mc_tables import EDGES, tables


def lattice(n, lo=-1.2, hi=1.2):
    ax = np.linspace(lo, hi, n)
    g = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1)
    return g, (ax[1] - ax[0],) * 3, (lo,) * 3


def closed_and_oriented(faces):
    de = {}
    for a, b, c in faces:
        for e in ((a, b), (b, c), (c, a)):
            de[e] = de.get(e, 0) + 1
    return all(v == 1 for v in de.values()) and all((b, a) in de for (a, b) in de)


def test_tables_are_complete_and_complementary():
    edges, count, tri = tables()
    assert edges.shape == (12, 2) and count.shape == (256,) and tri.shape[0] == 256
    assert count[0] == 0 and count[255] == 0 and count.max() == tri.shape[1] == 5
    for m in range(256):
        cut = {k for k, (a, b) in enumerate(EDGES) if ((m >> a) & 1) != ((m >> b) & 1)}
        used = set(int(e) for e in tri[m, :count[m]].reshape(-1))
        assert used == cut, m                                   # every cut edge carries a vertex, no other edge does
        assert (tri[m, count[m]:] == -1).all()
        assert count[m] == len(cut) - 2 * loops(tri[m, :count[m]]) if cut else count[m] == 0


def loops(t):
    """number of closed polygons a fan triangulation came from: connected components of its triangles"""
    parent = {}

    def find(x):
        while parent.setdefault(x, x) != x:
            x = parent[x]
        return x
    for a, b, c in t:
        ra, rb, rc = find(int(a)), find(int(b)), find(int(c))
        parent[rb] = ra
        parent[find(int(c))] = ra
    return len({find(x) for x in list(parent)})


@pytest.mark.parametrize("n", [9, 24])
def test_sphere_is_closed_oriented_and_has_the_right_area(n):
    g, sp, org = lattice(n)
    vol = (np.linalg.norm(g, axis=-1) - 0.8).astype(np.float32)
    tri, key = M.marching_cubes(vol, 0.0, sp, org)
    v, f = M.weld(tri, key)
    assert closed_and_oriented(f)
    assert v.shape[0] - 3 * f.shape[0] // 2 + f.shape[0] == 2                        # Euler characteristic of a sphere
    nrm = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    assert ((nrm * v[f].mean(1)).sum(1) > 0).all()                                   # normals point out of the inside (value < level)
    area = 0.5 * np.linalg.norm(nrm, axis=1).sum()
    assert abs(area - 4 * np.pi * 0.64) < (0.08 if n == 9 else 0.01) * 4 * np.pi * 0.64
    assert np.abs(np.linalg.norm(v, axis=1) - 0.8).max() < 2.5 * sp[0] ** 2          # vertices sit on the interpolated level set


def test_torus_and_random_fields_stay_watertight():
    g, sp, org = lattice(28)
    r = np.sqrt(g[..., 0] ** 2 + g[..., 1] ** 2)
    torus = (np.sqrt((r - 0.7) ** 2 + g[..., 2] ** 2) - 0.25).astype(np.float32)
    v, f = M.weld(*M.marching_cubes(torus, 0.0, sp, org))
    assert closed_and_oriented(f) and v.shape[0] - 3 * f.shape[0] // 2 + f.shape[0] == 0     # genus 1
    rng = np.random.default_rng(0)
    for _ in range(3):                                           # rough fields hit the ambiguous faces and multi-loop cases
        vol = rng.standard_normal((12, 12, 12)).astype(np.float32)
        vol[0], vol[-1], vol[:, 0], vol[:, -1], vol[:, :, 0], vol[:, :, -1] = 5, 5, 5, 5, 5, 5     # positive shell: surface closes inside
        v, f = M.weld(*M.marching_cubes(vol, 0.0))
        assert f.shape[0] > 100 and closed_and_oriented(f)


def test_culling_and_components():
    g, sp, org = lattice(26)
    two = np.minimum(np.linalg.norm(g - np.array([0.55, 0, 0]), axis=-1) - 0.4, np.linalg.norm(g + np.array([0.6, 0, 0]), axis=-1) - 0.25)
    v, f = M.weld(*M.marching_cubes(two.astype(np.float32), 0.0, sp, org))
    lab, area = M.components(v, f)
    assert len(area) == 2 and abs(area.max() - 4 * np.pi * 0.16) < 0.03 * 4 * np.pi * 0.16
    assert abs(area.min() - 4 * np.pi * 0.0625) < 0.06 * 4 * np.pi * 0.0625
    seen = v[:, 0] > 0                                           # the small sphere (x < 0) is unseen
    kept = M.cull_unseen(f, seen)
    assert kept.shape[0] == int((lab == np.argmax(area)).sum())
