"""CPU restatement of one whole reconstruction -- TEST / BASELINE INFRASTRUCTURE ONLY (parity unpinned,
like every file under oracle/; see oracle/nksr_oracle.py).

Restates the wiring of the open training model (models/nksr_net.py:41-133) that
`nksr.Reconstructor.reconstruct` runs for one cloud (examples/recons_simple.py:25-27,
examples/recons_waymo_cpu.py:48-63):

    feature = normal | unit view direction (sensor - xyz)          models/nksr_net.py:48-52
    SVH(voxel_size, depth).build_point_splatting(xyz)              models/nksr_net.py:57-62
    network.encoder / network.unet -> basis / normal features      models/nksr_net.py:73-78
    KernelField(dec_svh, interpolators, basis_features, approx)    models/nksr_net.py:91-96
    normal_xyz = voxel centres of the finest adaptive_depth levels models/nksr_net.py:100
    solve(pos_xyz, normal_xyz, -normal_features, 1e4/N, 1e4/K*W^2, 1)   models/nksr_net.py:101-112

The network is PyTorch in the reference and stays PyTorch here (BASELINE.json north_star): the caller
hands in the *same* torch modules the product uses (moved to the CPU); only the pooling that feeds
them (a CUDA kernel in the product) is restated in numpy.  Heavy arithmetic (hierarchy, Gram system,
PCG, field evaluation) runs in the C++/OpenMP restatement (oracle/nksr_oracle_cpu.cpp).
"""
from __future__ import annotations

import numpy as np

from . import cpu_port as P
from . import nksr_oracle as O


def _scatter_sum(index, values, n):
    """out[j] = sum of values[i] over index[i] == j, added in the order of i (what np.add.at does, one bincount per
    column: the same fp64 additions in the same order, without np.add.at's per-element dispatch)"""
    return np.stack([np.bincount(index, weights=values[:, c], minlength=n) for c in range(values.shape[1])], axis=1)


def _children_sum(keys_fine, keys_coarse, acc_fine):
    """sum of the (<= 8) children of every coarse voxel (parent key = child key >> 3)."""
    par = np.searchsorted(keys_coarse, keys_fine >> 3)
    return _scatter_sum(par, acc_fine, keys_coarse.shape[0])


def _pool27(osvh, l, acc, svh_cpp=None):
    """sum over the 27-neighbourhood, slots added in table order.  With the C++ hierarchy at hand the table and the sum
    come from it (OpenMP; tests/test_cpu_port.py holds the two bitwise equal), else from the numpy restatement."""
    if svh_cpp is not None:
        return svh_cpp.pool27(l, acc)
    nb = osvh.nbr27(l)                                   # (n,27) index or -1
    out = np.zeros_like(acc)
    for s in range(27):
        ok = nb[:, s] >= 0
        out[ok] += acc[nb[ok, s]]
    return out


def standin_features(osvh, svh_cpp, xyz, point_feat, network):
    """numpy restatement of nksr_b200.network.NKSRNetwork.encoder/unet (the seeded stand-in that replaces
    the closed sparse-conv U-Net): per level the 27-neighbourhood-smoothed sums of [feature, 1] over the
    points of each voxel; heads = the caller's torch modules run on the CPU."""
    import torch
    L = osvh.depth
    base0 = svh_cpp.locate(xyz)[0].astype(np.int64)
    src = np.concatenate([point_feat.astype(np.float64), np.ones((xyz.shape[0], 1))], axis=1)
    pooled, acc = [], None
    for l in range(L):
        if l == 0:
            ok = base0 >= 0
            acc = _scatter_sum(base0[ok], src[ok], osvh.n(0))
        else:
            acc = _children_sum(osvh.keys[l - 1], osvh.keys[l], acc)
        pooled.append(_pool27(osvh, l, acc, svh_cpp))
    C = network.kernel_dim
    basis, normal, up = {}, {}, None
    for l in range(L - 1, -1, -1):
        s = pooled[l]
        cnt = s[:, 3:4]
        mean = s[:, :3] / np.maximum(cnt, 1.0)
        nrm = mean / (np.linalg.norm(mean, axis=1, keepdims=True) + 1e-6)
        if up is not None:
            par = np.searchsorted(osvh.keys[l + 1], osvh.keys[l] >> 3)
            nrm = np.where(cnt > 0, nrm, up[par])
        up = nrm
        x = torch.from_numpy(np.concatenate([nrm, np.log1p(cnt)], axis=1).astype(np.float32))
        with torch.no_grad():
            b = (1.0 + 0.1 * torch.tanh(network.basis_heads[l](x))) / (C ** 0.5)
            z = network.interpolators[l](b)
        basis[l] = z.numpy().astype(np.float32)
        normal[l] = nrm.astype(np.float32)
    return [basis[l] for l in range(L)], [normal[l] for l in range(L)]


def reconstruct(xyz, normal=None, sensor=None, voxel_size=0.1, depth=4, adaptive_depth=2, network=None,
                approx_kernel_grad=False, solver_tol=1e-5, max_iter=2000, feats=None, normal_feats=None):
    """Returns dict(svh (CpuSvh), osvh (OracleSVH, same keys), feats, alpha, system (CpuSystem), iterations,
    relres, normal_xyz, normal_value).  `feats` / `normal_feats` override the stand-in network."""
    xyz = np.ascontiguousarray(xyz, np.float32)
    if normal is not None:
        pf = np.asarray(normal, np.float32)
    else:
        view = np.asarray(sensor, np.float32) - xyz
        pf = view / (np.linalg.norm(view, axis=-1, keepdims=True) + np.float32(1e-6))
    svh = P.CpuSvh(xyz, voxel_size, depth)
    osvh = O.OracleSVH(voxel_size, depth).build_from_keys([svh.keys(l) for l in range(depth)])
    if feats is None:
        feats, normal_feats = standin_features(osvh, svh, xyz, pf, network)
    ad = min(adaptive_depth, depth)
    nxyz = np.concatenate([svh.centers(d) for d in range(ad)])
    nval = -np.concatenate([normal_feats[d] for d in range(ad)]).astype(np.float32)
    W = float(np.float32(voxel_size))
    sysm = P.CpuSystem(svh, feats, xyz, nxyz, nval, 1e4 / xyz.shape[0], 1e4 / nxyz.shape[0] * (W ** 2), 1.0,
                       approx_kernel_grad)
    alpha, it, res = sysm.pcg(solver_tol, max_iter)
    return dict(svh=svh, osvh=osvh, feats=feats, alpha=alpha, system=sysm, iterations=it, relres=res,
                normal_xyz=nxyz, normal_value=nval)
