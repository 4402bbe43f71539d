"""Numbers behind tests/test_gpu_pipeline.py::test_chunked_reconstruction_blends_and_welds (seam continuity of the
partition-of-unity blend, parked chunks): prints every quantity the test bounds.  usage: python tools/chunk_diag.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nksr_b200  # noqa: E402
from tests import clouds  # noqa: E402

cuda = torch.device("cuda:0")
xyz, nrm = clouds.sphere(60_000, radius=3.5, noise=0.005)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
rec = nksr_b200.Reconstructor(cuda)
field = rec.reconstruct(t(xyz), t(nrm), detail_level=None, chunk_size=4.0, solver_tol=1e-5)
whole = rec.reconstruct(t(xyz), t(nrm), detail_level=None, voxel_size=0.1, solver_tol=1e-5)
mesh = field.extract_dual_mesh(mise_iter=1)
rng = np.random.default_rng(0)
q = (xyz[rng.integers(0, xyz.shape[0], 4000)] * rng.uniform(0.995, 1.005, (4000, 1))).astype(np.float32)
qa, qb = q.copy(), q.copy()
qa[:, 0], qb[:, 0] = -5e-5, 5e-5
fa, fb = field.evaluate_f(t(qa)).value, field.evaluate_f(t(qb)).value
wa, wb = whole.evaluate_f(t(qa)).value, whole.evaluate_f(t(qb)).value
scale = float(whole.evaluate_f(t(q * 1.03)).value.abs().median())
bound = torch.zeros_like(fa)
n_blend = torch.zeros_like(fa)
for k, fk in enumerate(field.fields):
    wk = field._weights(t(qa), k)
    jk = (fk.evaluate_f(t(qa)).value - fk.evaluate_f(t(qb)).value).abs()
    bound = torch.maximum(bound, torch.where(wk > 0, jk, torch.zeros_like(jk)))
    n_blend += (wk > 0).float()
excess = (fa - fb).abs() - bound
out = dict(scale=scale, n_blend_min=float(n_blend.min()), blend_jump_max=float((fa - fb).abs().max()),
           whole_jump_max=float((wa - wb).abs().max()), bound_max=float(bound.max()),
           excess_max=float(excess.max()), excess_p99=float(excess.quantile(0.99)),
           blend_vs_whole_median=float((fa - wa).abs().median()), blend_vs_whole_p90=float((fa - wa).abs().quantile(0.9)))
rec.chunk_tmp_device = torch.device("cpu")
parked = rec.reconstruct(t(xyz), t(nrm), detail_level=None, chunk_size=4.0, solver_tol=1e-5)
mp = parked.extract_dual_mesh(mise_iter=1)
out.update(parked_on_cpu=all(f_.svh.device.type == "cpu" for f_ in parked.fields),
           parked_faces_equal=bool(mp.f.shape == mesh.f.shape and torch.equal(mp.f, mesh.f)),
           parked_v_maxdiff=float((mp.v - mesh.v).abs().max()) if mp.v.shape == mesh.v.shape else -1.0)
print(json.dumps(out))
