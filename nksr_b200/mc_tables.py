"""Procedural marching-cubes tables for the dual-cell triangulation (DESIGN.md SPEC S9).

Corner c = (cx<<2)|(cy<<1)|cz; the 12 edges are enumerated axis-major (x edges, y edges,
z edges), each from its lower to its upper corner.  A corner is *inside* when f > 0
(reference convention: occupancy test `evaluate_f_bar(x) > 0`, models/loss.py:99; outward
normal = -grad f, models/loss.py:192-196).

The triangle table is generated, not transcribed: on every cube face the contour is traced
with the marching-squares rule "an ambiguous face separates its inside corners".  The rule
depends only on the face's own corner signs, so adjacent cells always agree and the mesh is
watertight wherever cells exist.  Segments are chained into closed loops and fan-triangulated
with the orientation that makes triangle normals point from inside to outside.

`tools/gen_mc_tables.py` writes these tables to `nksr_b200/csrc/mc_tables.inc` for the CUDA
kernels; `tests/test_cpu_oracle.py` checks the committed .inc against this generator.
"""
from __future__ import annotations

import numpy as np

CORNERS = np.array([[(c >> 2) & 1, (c >> 1) & 1, c & 1] for c in range(8)], dtype=np.int32)


def _edges():
    out = []
    for ax in range(3):
        bit = 1 << (2 - ax)
        for c in range(8):
            if not c & bit:
                out.append((c, c | bit, ax))
    return np.array(out, dtype=np.int32)


EDGES = _edges()  # (12, 3): corner_a, corner_b, axis


def _faces():
    faces = []
    for ax in range(3):
        u, v = (ax + 1) % 3, (ax + 2) % 3
        for side in (0, 1):
            quad = []
            for a, b in ((0, 0), (1, 0), (1, 1), (0, 1)):
                p = [0, 0, 0]
                p[ax], p[u], p[v] = side, a, b
                quad.append((p[0] << 2) | (p[1] << 1) | p[2])
            faces.append(quad[::-1] if side == 0 else quad)
    return faces


def build_tables():
    """-> (tri_table int8 [256, 15] padded with -1, tri_count int32 [256])."""
    eid = {}
    for e, (a, b, _) in enumerate(EDGES):
        eid[(int(a), int(b))] = eid[(int(b), int(a))] = e
    faces = _faces()
    all_tris = []
    for case in range(256):
        ins = [(case >> c) & 1 for c in range(8)]
        nxt = {}
        for quad in faces:
            s = [ins[c] for c in quad]
            if sum(s) in (0, 4):
                continue
            for i in range(4):
                if s[i] and not s[i - 1]:          # start of a run of inside corners
                    j = i
                    while s[(j + 1) % 4]:
                        j += 1
                    e_before = eid[(quad[i - 1], quad[i])]
                    e_after = eid[(quad[j % 4], quad[(j + 1) % 4])]
                    nxt[e_after] = e_before
        tris, seen = [], set()
        for start in sorted(nxt):
            if start in seen:
                continue
            loop, cur = [start], nxt[start]
            seen.add(start)
            while cur != start:
                loop.append(cur)
                seen.add(cur)
                cur = nxt[cur]
            for t in range(1, len(loop) - 1):
                tris.append((loop[0], loop[t + 1], loop[t]))
        all_tris.append(tris)
    maxt = max(len(t) for t in all_tris)
    assert maxt <= 5
    table = np.full((256, 15), -1, dtype=np.int8)
    count = np.zeros(256, dtype=np.int32)
    for c, tris in enumerate(all_tris):
        count[c] = len(tris)
        for t, tri in enumerate(tris):
            table[c, 3 * t:3 * t + 3] = tri
    return table, count
