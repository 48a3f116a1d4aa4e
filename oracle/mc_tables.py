"""Marching-cubes case tables, generated (not transcribed): test infrastructure AND the source of the table header the CUDA
kernel includes (tools/gen_mc_tables.py writes attentive_dfprior_b200/csrc/dfp_mesh_tables.h from this module).

The reference extracts its mesh with skimage.measure.marching_cubes (Lewiner's variant, src/utils/Mesher.py:463-480);
skimage is not available in this image, so its case tables cannot be used or compared against.  This module derives a
*face-consistent* table from first principles:

* corner i of a cell sits at (i & 1, (i >> 1) & 1, (i >> 2) & 1) in (x, y, z) index steps; the 12 edges join corners that
  differ in one bit; a corner is "inside" when its value is below the level;
* on every face the cut edges are joined into segments by a rule that only looks at that face's four corner signs (two cut
  edges: one segment; four cut edges: each inside corner is cut off separately), so the two cells sharing a face always
  agree and the surface is watertight;
* segments are oriented so that the inside is on the same side everywhere, chained into closed loops and fan-triangulated.

tables(): (edge_corners [12][2], tri_count [256], tri_edges [256][MAX_TRI][3]).
"""
from __future__ import annotations

import functools

import numpy as np

CORNER = np.array([[(i >> 0) & 1, (i >> 1) & 1, (i >> 2) & 1] for i in range(8)], dtype=np.float64)
EDGES = [(a, b) for a in range(8) for b in range(a + 1, 8) if bin(a ^ b).count("1") == 1]        # 12, lower corner first
EDGE_ID = {e: k for k, e in enumerate(EDGES)}
# faces: (axis, side) -> corners in cyclic order
FACES = []
for axis in range(3):
    for side in (0, 1):
        u, v = [a for a in range(3) if a != axis]
        cyc = []
        for (du, dv) in ((0, 0), (1, 0), (1, 1), (0, 1)):
            c = [0, 0, 0]
            c[axis], c[u], c[v] = side, du, dv
            cyc.append(c[0] | (c[1] << 1) | (c[2] << 2))
        n = np.zeros(3)
        n[axis] = 1.0 if side else -1.0
        FACES.append((cyc, n))


def _edge(a, b):
    return EDGE_ID[(min(a, b), max(a, b))]


def _mid(e):
    a, b = EDGES[e]
    return 0.5 * (CORNER[a] + CORNER[b])


EDGE_FACES = [{f for f, (cyc, _) in enumerate(FACES) if a in cyc and b in cyc} for a, b in EDGES]


def _triangulate(loop):
    """Triangles of a closed polygon through cut edges, none of whose diagonals lies in a face of the cell: a diagonal between
    two cut edges of the same face would coincide with a segment or diagonal of the neighbouring cell (an edge with four
    triangles).  Polygons have at most 8 corners here; the first admissible triangulation in enumeration order is taken."""
    def ok(a, b):
        return not (EDGE_FACES[a] & EDGE_FACES[b])

    def rec(poly):
        if len(poly) == 3:
            return [tuple(poly)]
        # split off a triangle (poly[0], poly[k], poly[-1]) ... standard enumeration over the apex of the edge (poly[-1], poly[0])
        for k in range(1, len(poly) - 1):
            left, right = poly[:k + 1], poly[k:]
            if (k > 1 and not ok(poly[0], poly[k])) or (k < len(poly) - 2 and not ok(poly[k], poly[-1])):
                continue
            a = rec(left) if len(left) >= 3 else []
            b = rec(right) if len(right) >= 3 else []
            if a is None or b is None:
                continue
            return a + [(poly[0], poly[k], poly[-1])] + b
        return None

    out = rec(list(loop))
    assert out is not None, ("no face-free triangulation", loop)
    assert len(out) == len(loop) - 2
    return out


def _case(mask):
    inside = [(mask >> i) & 1 for i in range(8)]
    nxt = {}                                       # cut edge -> next cut edge along the oriented loop
    for cyc, n in FACES:
        s = [inside[c] for c in cyc]
        cut = [k for k in range(4) if s[k] != s[(k + 1) % 4]]          # face edge k joins cyc[k], cyc[k+1]
        segs = []
        if len(cut) == 2:
            ins = next(cyc[k] for k in range(4) if s[k])
            segs.append((cut[0], cut[1], ins))
        elif len(cut) == 4:                         # ambiguous face: cut every inside corner off on its own
            for k in range(4):
                if s[k]:
                    segs.append(((k - 1) % 4, k, cyc[k]))
        for k1, k2, ins in segs:
            e1, e2 = _edge(cyc[k1], cyc[(k1 + 1) % 4]), _edge(cyc[k2], cyc[(k2 + 1) % 4])
            p1, p2 = _mid(e1), _mid(e2)
            # seen from outside the cell (along -n), keep the inside corner on the left of the directed segment
            side = np.dot(np.cross(p2 - p1, n), CORNER[ins] - p1)
            if side < 0:
                e1, e2 = e2, e1
            assert e1 not in nxt, "orientation rule is inconsistent"
            nxt[e1] = e2
    tris = []
    seen = set()
    for start in sorted(nxt):
        if start in seen:
            continue
        loop = [start]
        seen.add(start)
        e = nxt[start]
        while e != start:
            assert e not in seen
            seen.add(e)
            loop.append(e)
            e = nxt[e]
        assert len(loop) >= 3
        tris.extend(_triangulate(loop))
    assert set(nxt.values()) == set(nxt.keys())
    return tris


@functools.lru_cache(maxsize=1)
def tables():
    cases = [_case(m) for m in range(256)]
    max_tri = max(len(c) for c in cases)
    count = np.array([len(c) for c in cases], dtype=np.int32)
    tri = np.full((256, max_tri, 3), -1, dtype=np.int32)
    for m, c in enumerate(cases):
        for t, abc in enumerate(c):
            tri[m, t] = abc
    return np.array(EDGES, dtype=np.int32), count, tri


if __name__ == "__main__":
    e, c, t = tables()
    print("edges", e.tolist())
    print("max triangles per cell", t.shape[1], "total", int(c.sum()))
