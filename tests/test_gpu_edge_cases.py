"""Edge cases of the CUDA path: colliding (duplicate) points, non-finite input, tiny and ragged clouds,
queries far outside the band, repeated solves on one hierarchy."""
import numpy as np
import pytest
import torch

from oracle import nksr_oracle as O
from tests import clouds

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy()


def test_duplicate_points_collapse_to_one_voxel_set(cuda):
    import nksr_b200
    p = np.tile(np.array([[0.31, -0.77, 1.05]], np.float32), (5000, 1))
    svh = nksr_b200.SparseFeatureHierarchy(0.1, 4, cuda).build_point_splatting(torch.from_numpy(p).to(cuda))
    osvh = O.OracleSVH(0.1, 4).build_point_splatting(p)
    for l in range(4):
        assert svh.num_voxels(l) == 8 and np.array_equal(_np(svh.keys[l]), osvh.keys[l])
    # 5000 identical position rows + one normal per voxel: system still assembles and solves
    feats = [torch.full((8, 4), 0.5, device=cuda) for _ in range(4)]
    field = nksr_b200.KernelField(svh, None, feats)
    field.solver_config.update(keep_system=True, tol=1e-6, check_every=1)
    nxyz = svh.get_voxel_centers(0)
    nval = torch.tensor([[0.0, 0.0, -1.0]], device=cuda).expand(8, 3).contiguous()
    field.solve(torch.from_numpy(p).to(cuda), nxyz, nval, 2.0, 0.01, 1.0)
    A_ref, b_ref, _ = O.build_system(osvh, [np.full((8, 4), 0.5, np.float32)] * 4, p, _np(nxyz), _np(nval), 2.0, 0.01, 1.0)
    import scipy.sparse as sp
    s = field.system
    A = sp.csr_matrix((_np(s.val).astype(np.float64), _np(s.col), _np(s.rowptr)), shape=(s.n, s.n))
    assert abs(A - A_ref).max() <= 5e-4 * abs(A_ref).max()
    assert np.abs(_np(s.rhs) - b_ref).max() <= 5e-4 * max(np.abs(b_ref).max(), 1e-12)
    # 5000 coincident rows make the system nearly rank deficient: only require a usable fp32 solve
    alpha = _np(field.alpha).astype(np.float64)
    assert np.isfinite(alpha).all() and field.solve_info["relative_residual"] <= 1e-3


def test_non_finite_and_out_of_range_inputs_raise(cuda):
    import nksr_b200
    bad = torch.tensor([[0.0, 0.0, 0.0], [float("nan"), 0.0, 0.0]], device=cuda)
    with pytest.raises(RuntimeError):
        nksr_b200.SparseFeatureHierarchy(0.1, 4, cuda).build_point_splatting(bad)
    far = torch.tensor([[0.0, 0.0, 0.0], [1.0e7, 0.0, 0.0]], device=cuda)
    with pytest.raises(RuntimeError):
        nksr_b200.SparseFeatureHierarchy(0.1, 4, cuda).build_point_splatting(far)


def test_queries_outside_the_band_evaluate_to_zero(cuda):
    import nksr_b200
    xyz, nrm = clouds.sphere(2000)
    rec = nksr_b200.Reconstructor(cuda, tree_depth=3)
    field = rec.reconstruct(torch.from_numpy(xyz).to(cuda), torch.from_numpy(nrm).to(cuda), voxel_size=0.05)
    q = torch.tensor([[50.0, 50.0, 50.0], [-3.0, 2.0, 9.0], [1e6, 0.0, 0.0], [float("inf"), 0.0, 0.0]], device=cuda)
    r = field.evaluate_f(q, grad=True)
    assert torch.equal(r.value, torch.zeros(4, device=cuda)) and torch.equal(r.gradient, torch.zeros(4, 3, device=cuda))
    assert not field.mask_field.mask(q).any()


def test_two_separate_components_and_resolve(cuda):
    """ragged input: two blobs of very different size; the hierarchy is reused for a second solve."""
    import nksr_b200
    a, na = clouds.sphere(6000, radius=0.3, centre=(0.0, 0.0, 0.0))
    b, nb = clouds.sphere(40, radius=0.05, centre=(2.0, 1.0, -1.0), seed=9)
    xyz, nrm = np.concatenate([a, b]), np.concatenate([na, nb])
    t = lambda v: torch.from_numpy(v).to(cuda)
    rec = nksr_b200.Reconstructor(cuda, tree_depth=3)
    field = rec.reconstruct(t(xyz), t(nrm), voxel_size=0.04, solver_tol=1e-6)
    mesh = field.extract_dual_mesh()
    v = _np(mesh.v)
    big = np.linalg.norm(v, axis=1) < 1.0
    assert big.sum() > 500 and abs(np.median(np.linalg.norm(v[big], axis=1)) - 0.3) < 0.01
    alpha1 = field.alpha.clone()
    # a second reconstruction of the same cloud reproduces the coefficients (the stand-in network
    # pools features with torch index_add_, whose atomics reorder fp32 sums: equal to rounding only)
    info1 = dict(field.solve_info)
    field2 = rec.reconstruct(t(xyz), t(nrm), voxel_size=0.04, solver_tol=1e-6)
    assert torch.allclose(field2.alpha, alpha1, rtol=1e-3, atol=1e-5 * float(alpha1.abs().max()))
    assert abs(field2.solve_info["iterations"] - info1["iterations"]) <= 0.03 * info1["iterations"] + 3
    for l in range(3):
        assert torch.equal(field2.svh.keys[l], field.svh.keys[l])
