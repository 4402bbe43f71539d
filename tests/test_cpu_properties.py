"""Property tests of the CPU oracle (hypothesis): the invariants DESIGN.md's SPEC promises, on random
small clouds -- partition of unity of the basis, parent closure, symmetry / positive definiteness
and pattern containment of the Gram matrix, translation behaviour of the integer indexing."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import nksr_oracle as O


@st.composite
def small_cloud(draw):
    n = draw(st.integers(8, 120))
    seed = draw(st.integers(0, 2 ** 16))
    scale = draw(st.sampled_from([0.3, 1.0, 7.0]))
    shift = draw(st.sampled_from([0.0, -13.37, 101.5]))
    rng = np.random.default_rng(seed)
    xyz = (rng.normal(size=(n, 3)) * scale + shift).astype(np.float32)
    W = float(draw(st.sampled_from([0.05, 0.11, 0.5]))) * scale
    L = draw(st.integers(1, 4))
    return xyz, W, L, seed


@settings(max_examples=25, deadline=None)
@given(small_cloud())
def test_hierarchy_invariants(c):
    xyz, W, L, _ = c
    svh = O.OracleSVH(W, L).build_point_splatting(xyz)
    base = svh.locate(xyz)
    assert (base >= 0).all()                                   # every point lies in an active voxel on every level
    for l in range(L):
        k = svh.keys[l]
        assert (np.diff(k) > 0).all()                          # sorted, unique
        if l + 1 < L:
            assert np.isin(np.unique(k >> 3), svh.keys[l + 1]).all()      # parent closed
        # containing voxel = reference formula floor(x / W_l)
        ijk = svh.ijk(l)[base[l]]
        ref = np.floor(xyz / np.float32(svh.level_w(l))).astype(np.int32)
        assert (ijk == ref).all()
        # the 8 voxels whose centres surround a point are all active (trilinear support complete)
        hl = O.quantize_half(xyz, W).astype(np.int64) >> l
        b8 = ((hl - 1) >> 1)[:, None, :] + O._OFF8[None]
        assert (svh.lookup(l, b8) >= 0).all()


@settings(max_examples=20, deadline=None)
@given(small_cloud(), st.booleans())
def test_basis_partition_of_unity_and_rows(c, approx):
    xyz, W, L, seed = c
    svh = O.OracleSVH(W, L).build_point_splatting(xyz)
    base = svh.locate(xyz)
    for l in range(L):
        ones = np.ones((svh.n(l), 1), np.float32)              # constant feature 1 -> phi = 1 -> K = B^3
        nbr, K, dK = O.level_rows(svh, l, xyz, base[l], ones, True, approx)
        full = (nbr >= 0).all(axis=1)                          # rows whose whole 27-stencil is active
        assert np.allclose(K[full].sum(axis=1), 1.0, atol=1e-9)           # quadratic B-splines sum to one ...
        assert np.allclose(dK[full].sum(axis=2), 0.0, atol=1e-6 / svh.level_w(l))   # ... gradients to zero
        assert (K >= -1e-12).all() and (K.sum(axis=1) <= 1.0 + 1e-9).all()


@settings(max_examples=12, deadline=None)
@given(small_cloud())
def test_gram_matrix_properties(c):
    xyz, W, L, seed = c
    svh = O.OracleSVH(W, L).build_point_splatting(xyz)
    rng = np.random.default_rng(seed)
    feats = [(0.5 + 0.2 * rng.normal(size=(svh.n(l), 3))).astype(np.float32) for l in range(L)]
    nxyz = svh.centers(0)
    nval = rng.normal(size=nxyz.shape)
    A, b, E = O.build_system(svh, feats, xyz, nxyz, nval, 2.0, 0.3 * W * W, 1.0)
    assert abs(A - A.T).max() <= 1e-9 * abs(A).max()
    v = rng.normal(size=A.shape[0])
    assert v @ (A @ v) > 0
    P = O.structural_pattern(svh)
    assert (A - A.multiply(P)).count_nonzero() == 0            # stored pattern contains every nonzero
    x, it, res = O.pcg(A, b, 1e-8, 5 * A.shape[0] + 50)
    assert res <= 1e-8                                          # CG converges on an SPD system
