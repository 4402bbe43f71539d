"""Opt-in variants that were written after the round's GPU budget was spent.  They are skipped unless
NKSR_EXPERIMENTAL=1 so that an unvalidated path cannot turn the suite red; the first GPU call of the next
round runs them (`NKSR_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu`).

  * row_layout = 'interleaved' (nksr_build_rows modes 3 / 4, k_gram_blocks4, k_gram_fill<..., INTER>):
    same floating-point operations in the same order as the line layout, so the system must be BITWISE equal.
"""
import os

import numpy as np
import pytest
import torch

from oracle import nksr_oracle as O
from tests import clouds

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.environ.get("NKSR_EXPERIMENTAL"),
                                 reason="opt-in variant awaiting its first GPU validation (NKSR_EXPERIMENTAL=1)")]


def _np(t):
    return t.detach().cpu().numpy()


def _setup(cuda, L, W, approx, C=4):
    import nksr_b200
    xyz, _ = clouds.shapenet_like(3000)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    svh = nksr_b200.SparseFeatureHierarchy(W, L, cuda).build_point_splatting(t(xyz))
    osvh = O.OracleSVH(W, L).build_point_splatting(xyz)
    rng = np.random.default_rng(7)
    feats = [(0.5 + 0.2 * rng.normal(size=(osvh.n(l), C))).astype(np.float32) for l in range(L)]
    nxyz = np.concatenate([osvh.centers(d) for d in range(min(2, L))])
    nval = -(nxyz / np.linalg.norm(nxyz, axis=1, keepdims=True)).astype(np.float32)
    field = nksr_b200.KernelField(svh, None, [t(f) for f in feats], approx)
    return field, t(xyz), t(nxyz), t(nval), (1e4 / xyz.shape[0], 1e4 / nxyz.shape[0] * W * W, 1.0)


@pytest.mark.parametrize("L,W,approx", [(4, 0.02, False), (4, 0.02, True), (3, 0.03, False), (1, 0.06, False)])
def test_interleaved_rows_are_the_line_rows_transposed(cuda, L, W, approx):
    field, xyz, nxyz, nval, _ = _setup(cuda, L, W, approx)
    _, _, _, _, e0 = field._sorted_rows(xyz, 0)
    _, _, _, _, e3 = field._sorted_rows(xyz, 3)
    e0, e3 = _np(e0), _np(e3)                                   # (m, L, 32) and (m, 1, 32, 4)
    assert e3.shape == (xyz.shape[0], 1, 32, 4)
    for l in range(4):
        assert np.array_equal(e3[:, 0, :, l], e0[:, l, :] if l < L else np.zeros_like(e0[:, 0, :]))
    _, _, _, _, e1 = field._sorted_rows(nxyz, 1, nval)
    _, _, _, _, e4 = field._sorted_rows(nxyz, 4, nval)
    e1, e4 = _np(e1).reshape(nxyz.shape[0], L, 3, 32), _np(e4)  # (k, L, 3, 32) and (k, 3, 32, 4)
    for l in range(4):
        for a in range(3):
            assert np.array_equal(e4[:, a, :, l], e1[:, l, a, :] if l < L else np.zeros_like(e1[:, 0, 0, :]))


@pytest.mark.parametrize("L,W,approx,split", [(4, 0.02, False, None), (4, 0.02, True, 4), (4, 0.02, False, 1),
                                              (3, 0.03, False, 2), (2, 0.04, True, None), (1, 0.06, False, None)])
def test_interleaved_layout_gives_the_same_system_bitwise(cuda, L, W, approx, split):
    field, xyz, nxyz, nval, (pw, nw, rw) = _setup(cuda, L, W, approx)
    out = []
    for layout in ("lines", "interleaved"):
        field.solver_config.update(keep_system=True, max_iter=0, row_layout=layout, block_split_level=split)
        field.solve(xyz, nxyz, nval, pw, nw, rw)
        s = field.system
        out.append([_np(x).copy() for x in (s.rowptr, s.col, s.val, s.rhs, s.diag)])
    for a, b in zip(*out):
        assert np.array_equal(a, b)

