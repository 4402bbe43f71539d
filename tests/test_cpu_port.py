"""The two CPU restatements must agree: oracle/nksr_oracle.py (numpy/scipy) and
oracle/nksr_oracle_cpu.cpp (C++/OpenMP, written independently).  With no reference vectors available
 this cross-check is what pins the oracle."""
import numpy as np
import pytest

from oracle import cpu_port as P
from oracle import nksr_oracle as O
from tests import clouds


@pytest.fixture(scope="module", autouse=True)
def _built():
    P.lib()


@pytest.mark.parametrize("cloud,W,L", [("shapenet", 0.02, 4), ("blob", 0.1, 3), ("sphere", 0.05, 1)])
def test_hierarchy_keys_bit_exact(cloud, W, L):
    xyz = {"shapenet": clouds.shapenet_like(3000)[0], "blob": clouds.offset_blob(8000)[0],
           "sphere": clouds.sphere(2000)[0]}[cloud]
    a, b = P.CpuSvh(xyz, W, L), O.OracleSVH(W, L).build_point_splatting(xyz)
    for l in range(L):
        assert np.array_equal(a.keys(l), b.keys[l])
        assert np.array_equal(a.centers(l), b.centers(l))


@pytest.mark.parametrize("C,approx", [(4, False), (8, True)])
def test_system_and_solution_agree(C, approx):
    xyz, _ = clouds.shapenet_like(1500)
    W, L = 0.03, 3
    a, b = P.CpuSvh(xyz, W, L), O.OracleSVH(W, L).build_point_splatting(xyz)
    rng = np.random.default_rng(5)
    feats = [(0.5 + 0.2 * rng.normal(size=(b.n(l), C))).astype(np.float32) for l in range(L)]
    nxyz = np.concatenate([b.centers(0), b.centers(1)])
    nval = rng.normal(size=nxyz.shape).astype(np.float32)
    sysm = P.CpuSystem(a, feats, xyz, nxyz, nval, 3.0, 0.05, 1.0, approx)
    A, rhs = sysm.to_scipy()
    A_ref, b_ref, _ = O.build_system(b, feats, xyz, nxyz, nval, 3.0, 0.05, 1.0, approx)
    assert A.nnz == A_ref.nnz
    assert abs(A - A_ref).max() <= 2e-7 * abs(A_ref).max()            # fp32 storage of the entries
    assert np.abs(rhs - b_ref).max() <= 2e-7 * np.abs(b_ref).max()
    x, it, res = sysm.pcg(1e-6, 4000)
    xo, ito, reso = O.pcg(A_ref, b_ref, 1e-6, 4000, dtype=np.float32)
    assert res <= 1e-6 and abs(it - ito) <= 0.1 * ito + 3
    assert np.linalg.norm(A_ref @ x - b_ref) <= 1e-5 * np.linalg.norm(b_ref)


@pytest.mark.parametrize("cloud,W,L", [("shapenet", 0.02, 4), ("blob", 0.1, 3)])
def test_neighbour_table_and_pooling_are_the_numpy_ones(cloud, W, L):
    """the OpenMP 27-neighbour table / 27-neighbourhood sum of the C++ hierarchy (what the CPU baseline's stand-in
    features run on) against the numpy restatement: indices identical, sums bitwise identical (same fp64 additions in
    the same slot order)"""
    xyz = {"shapenet": clouds.shapenet_like(3000)[0], "blob": clouds.offset_blob(8000)[0]}[cloud]
    a = P.CpuSvh(xyz, W, L)
    b = O.OracleSVH(W, L).build_from_keys([a.keys(l) for l in range(L)])
    rng = np.random.default_rng(0)
    for l in range(L):
        nb = b.nbr27(l)
        assert np.array_equal(a.nbr27(l), nb.astype(np.int32))
        acc = rng.normal(size=(a.n(l), 4))
        ref = np.zeros_like(acc)
        for s in range(27):
            ok = nb[:, s] >= 0
            ref[ok] += acc[nb[ok, s]]
        assert np.array_equal(a.pool27(l, acc), ref)


def test_baseline_helpers_keep_the_arithmetic():
    """bench.py's CPU arm runs these on all cores; threading / vectorising them must not change a bit: the scatter sum
    (np.add.at order), the blocked PCA normals, and the stand-in features with and without the C++ tables"""
    import torch
    from oracle import normals, pipeline
    from nksr_b200.network import NKSRNetwork
    rng = np.random.default_rng(1)
    idx = rng.integers(0, 500, 20000)
    val = rng.normal(size=(20000, 4))
    ref = np.zeros((500, 4))
    np.add.at(ref, idx, val)
    assert np.array_equal(pipeline._scatter_sum(idx, val, 500), ref)
    xyz = rng.normal(size=(40000, 3)).astype(np.float32)
    nn, _ = normals.knn_indices(xyz, 16)
    n1, w1 = normals._pca_block(np.asarray(xyz, np.float64), nn)
    n2, w2 = normals.pca_normals(xyz, nn, workers=4)
    assert np.array_equal(n1, n2) and np.array_equal(w1, w2)
    pts, nrm = clouds.sphere(5000, noise=0.002)
    svh = P.CpuSvh(pts, 0.05, 3)
    osvh = O.OracleSVH(0.05, 3).build_from_keys([svh.keys(l) for l in range(3)])
    net = NKSRNetwork(dict(kernel_dim=4, tree_depth=3, adaptive_depth=2))
    fa, na = pipeline.standin_features(osvh, svh, pts, nrm, net)

    class NoTables:                                            # the numpy tables: hide the C++ pooling
        locate = svh.locate
    old = pipeline._pool27
    pipeline._pool27 = lambda o, l, acc, cpp=None: old(o, l, acc, None)
    try:
        fb, nb_ = pipeline.standin_features(osvh, NoTables, pts, nrm, net)
    finally:
        pipeline._pool27 = old
    for l in range(3):
        assert np.array_equal(fa[l], fb[l]) and np.array_equal(na[l], nb_[l])
