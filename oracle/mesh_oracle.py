"""CPU restatement of the mesh extraction that follows the Mesher's dense query (test infrastructure; only tests/ may use it).

The reference runs skimage.measure.marching_cubes on the queried occupancy lattice, converts the vertices back to world
coordinates, drops the faces none of whose vertices was seen, and keeps large connected components
(src/utils/Mesher.py:463-512).  skimage / trimesh / open3d do not exist in this image, so this oracle cannot be pinned to
their outputs ("parity unpinned" for the triangulation itself); it pins the CUDA kernels to a plain numpy statement of the
same algorithm with the generated tables of oracle/mc_tables.py, and the tests check the properties any correct extraction
has (watertight, consistently oriented, vertices on the level set of the trilinear interpolant's edges, area of a sphere).
"""
from __future__ import annotations

import numpy as np

from .mc_tables import tables


def marching_cubes(vol, level, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0)):
    """vol [nx, ny, nz] float32.  Returns (tri [T, 3, 3] float32 world positions, key [T, 3] int64 global edge ids), cells in
    C order, triangles in table order.  A corner is inside when its value is below ``level``."""
    vol = np.ascontiguousarray(vol, dtype=np.float32)
    nx, ny, nz = vol.shape
    edges, count, tri = tables()
    level = np.float32(level)
    sp = np.asarray(spacing, dtype=np.float32)
    org = np.asarray(origin, dtype=np.float32)
    inside = vol < level
    mask = np.zeros((nx - 1, ny - 1, nz - 1), dtype=np.int32)
    for c in range(8):
        dx, dy, dz = c & 1, (c >> 1) & 1, (c >> 2) & 1
        mask |= inside[dx:nx - 1 + dx, dy:ny - 1 + dy, dz:nz - 1 + dz].astype(np.int32) << c
    cells = np.argwhere(count[mask] > 0)                       # C order
    out_t, out_k = [], []
    for x, y, z in cells:
        m = mask[x, y, z]
        for t in range(count[m]):
            pos, key = [], []
            for e in tri[m, t]:
                a, b = edges[e]
                ca = np.array([x + (a & 1), y + ((a >> 1) & 1), z + ((a >> 2) & 1)])
                axis = int(np.log2(a ^ b))
                va, vb = vol[tuple(ca)], vol[tuple(ca + np.eye(3, dtype=np.int64)[axis])]
                tt = np.float32(level - va) / np.float32(vb - va)
                p = ca.astype(np.float32)
                p[axis] = np.float32(p[axis] + tt)
                pos.append(np.float32(p * sp) + org)
                key.append(((int(ca[0]) * ny + int(ca[1])) * nz + int(ca[2])) * 3 + axis)
            out_t.append(pos)
            out_k.append(key)
    if not out_t:
        return np.zeros((0, 3, 3), np.float32), np.zeros((0, 3), np.int64)
    return np.asarray(out_t, dtype=np.float32), np.asarray(out_k, dtype=np.int64)


def weld(tri, key):
    """Triangle soup -> (verts [V, 3], faces [F, 3]): vertices on the same lattice edge are one vertex (ascending edge id)."""
    uniq, first, inv = np.unique(key.reshape(-1), return_index=True, return_inverse=True)
    return tri.reshape(-1, 3)[first], inv.reshape(-1, 3).astype(np.int64)


def cull_unseen(faces, seen):
    """Mesher.py:489-493: a face goes when ALL its vertices are unseen."""
    keep = ~(~seen[faces]).all(axis=1)
    return faces[keep]


def components(verts, faces):
    """Connected components of the face graph (faces sharing a vertex) and their areas: (label per face, area per label)."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    nv = verts.shape[0]
    i = np.concatenate([faces[:, 0], faces[:, 1]])
    j = np.concatenate([faces[:, 1], faces[:, 2]])
    g = coo_matrix((np.ones(i.shape[0], dtype=np.int8), (i, j)), shape=(nv, nv))
    _, vlab = connected_components(g, directed=False)
    flab = vlab[faces[:, 0]]
    a = 0.5 * np.linalg.norm(np.cross(verts[faces[:, 1]] - verts[faces[:, 0]], verts[faces[:, 2]] - verts[faces[:, 0]]), axis=1)
    area = np.bincount(flab, weights=a, minlength=int(vlab.max()) + 1 if nv else 0)
    return flab, area
