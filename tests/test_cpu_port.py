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
