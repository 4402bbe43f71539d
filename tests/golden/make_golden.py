"""Generates tests/golden/*.npz from the CPU oracle (run from the repo root).

The reference holds no golden vectors for this path, so these are
SELF-goldens: they freeze the oracle's output for small seeded inputs so that neither the oracle
nor the CUDA path can drift silently.  Regenerate only on a deliberate SPEC change.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import nksr_oracle as O  # noqa: E402
from tests import clouds  # noqa: E402


def case(name, xyz, W, L, C, seed, approx):
    svh = O.OracleSVH(W, L).build_point_splatting(xyz)
    rng = np.random.default_rng(seed)
    feats = [(0.5 + 0.2 * rng.normal(size=(svh.n(l), C))).astype(np.float32) for l in range(L)]
    ad = min(2, L)
    nxyz = np.concatenate([svh.centers(d) for d in range(ad)])
    nval = -(nxyz - xyz.mean(0)) / np.linalg.norm(nxyz - xyz.mean(0), axis=1, keepdims=True)
    nval = nval.astype(np.float32)
    pw, nw = 1e4 / xyz.shape[0], 1e4 / nxyz.shape[0] * W * W
    A, b, _ = O.build_system(svh, feats, xyz, nxyz, nval, pw, nw, 1.0, approx)
    A.sort_indices()
    alpha, it, res = O.pcg(A, b, 1e-8, 5000)
    q = (xyz[:64] + np.float32(0.25 * W)).astype(np.float32)
    f, g = O.evaluate_f(svh, feats, alpha, q, grad=True, approx_kernel_grad=approx)
    v, tri = O.extract_dual_mesh(svh, lambda p: O.evaluate_f(svh, feats, alpha, p), 1, 1)
    out = dict(xyz=xyz, W=np.float32(W), L=L, C=C, approx=approx, nxyz=nxyz, nval=nval, pw=pw, nw=nw,
               A_indptr=A.indptr, A_indices=A.indices, A_data=A.data, b=b, alpha=alpha, pcg_iters=it,
               q=q, f=f, g=g, mesh_v=v, mesh_f=tri)
    for l in range(L):
        out[f"keys{l}"] = svh.keys[l]
        out[f"feat{l}"] = feats[l]
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz"), **out)
    print(name, "n", A.shape[0], "nnz", A.nnz, "iters", it, "V", v.shape[0], "T", tri.shape[0])


if __name__ == "__main__":
    case("sphere256_L3_C4", clouds.sphere(256, radius=0.3, noise=0.002, seed=21)[0], 0.1, 3, 4, 1, False)
    case("blob400_L2_C8_approx", (clouds.offset_blob(400, seed=9, scale=0.4)[0]), 0.12, 2, 8, 2, True)
