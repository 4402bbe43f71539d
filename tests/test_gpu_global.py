"""GPU tests of the global-solve driver (nksr_b200/dist_solve.py) in a single process (world = 1):
the torch/NCCL-driven PCG must reproduce the in-library PCG of the ordinary path.  The two-rank
version of this check is tools/check_global_solve.py (run with torchrun on 2 GPUs)."""
import numpy as np
import pytest
import torch

from tests import clouds

pytestmark = pytest.mark.gpu


def test_global_solve_world1_matches_single_gpu_path(cuda):
    import nksr_b200
    from nksr_b200 import dist_solve as ds
    xyz, nrm = clouds.sphere(30000, noise=0.001)
    t = lambda a: torch.from_numpy(a).to(cuda)
    rec = nksr_b200.Reconstructor(cuda, tree_depth=3)
    ref = rec.reconstruct(t(xyz), t(nrm), voxel_size=0.03, solver_tol=1e-6)
    glob = ds.reconstruct_global(rec, t(xyz), t(nrm), 0.03, solver_tol=1e-6)
    assert glob.owned.all() and glob.solve_info["halo_recv"] == 0
    for l in range(3):
        assert torch.equal(ref.svh.keys[l], glob.svh.keys[l])
    a, b = ref.alpha.double(), glob.alpha.double()
    assert float((a - b).abs().max()) <= 2e-3 * float(a.abs().max())
    q = t(xyz[:2000])
    assert float((ref.evaluate_f(q).value - glob.evaluate_f(q).value).abs().max()) < 1e-3
    mesh = ds.extract_global_mesh(glob, mise_iter=1)
    r = np.linalg.norm(mesh.v.cpu().numpy(), axis=1)
    assert mesh.f.shape[0] > 1000 and abs(np.median(r) - 0.35) < 0.004
