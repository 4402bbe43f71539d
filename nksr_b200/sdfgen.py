"""sdfgen -- host mirror of the reference's native extension `ext.sdfgen` (ext/__init__.py:18-23).

    ext.sdfgen.sdf_from_points(queries, ref_xyz, ref_normal, nb_points, stdv, compute_grad=False, imls=False,
                               adaptive_knn=0) -> [sdf] or [sdf, grad]          ext/sdfgen/bind.cpp:9-15

Call sites: dataset/av_gt_geometry.py:63-78 (`nb_points=8, stdv=3.0, adaptive_knn=8`, ground-truth SDF of the training
crops) and models/loss.py:85 (`8, 0.02`).  Same arguments, same return convention, CUDA tensors only.  The reference
builds a tinyflann kd-tree per call (sdf_from_points.cu:150-156); here the reference points are hashed once per call into
the multi-level voxel hash of nksr_b200 and one kernel searches and votes (csrc/sdfgen.cu).
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import _lib
from ._lib import call, stream_ptr

START_LEVEL = 1


def sdf_from_points(queries: torch.Tensor, ref_xyz: torch.Tensor, ref_normal: torch.Tensor, nb_points: int,
                    stdv: float, compute_grad: bool = False, imls: bool = False,
                    adaptive_knn: int = 0) -> List[torch.Tensor]:
    _lib.require_cuda(queries, "queries")
    _lib.require_cuda(ref_xyz, "ref_xyz")
    from .reconstructor import _knn_hash
    dev = queries.device
    q = queries.detach().to(torch.float32).contiguous()
    ref = ref_xyz.detach().to(dev, torch.float32).contiguous()
    nrm = ref_normal.detach().to(dev, torch.float32).contiguous()
    if not 1 <= int(nb_points) <= 64 or int(adaptive_knn) > 64:
        raise ValueError("nb_points and adaptive_knn must be in 1..64")
    n_ref, m = ref.shape[0], q.shape[0]
    if n_ref == 0:
        raise _lib.NksrError("sdf_from_points: empty reference cloud")
    perm, svh, _, ranges, origin = _knn_hash(ref)
    xs, ns = ref[perm].contiguous(), nrm[perm].contiguous()
    o3 = (C.c_float * 3)(*[float(v) for v in origin.tolist()])
    st = stream_ptr(dev)
    ref_std = None
    if adaptive_knn > 0:                                  # sdf_from_points.cu:158-166
        ref_std = torch.empty(n_ref, dtype=torch.float32, device=dev)
        call("nksr_knn_mean_distance", svh.view(), xs, ranges, n_ref, C.addressof(o3), xs, n_ref, int(adaptive_knn),
             START_LEVEL, ref_std, st)
    sdf = torch.empty(m, dtype=torch.float32, device=dev)
    grad = torch.empty((m, 3), dtype=torch.float32, device=dev) if compute_grad else None
    call("nksr_sdf_from_points", svh.view(), xs, ns, ref_std, ranges, n_ref, C.addressof(o3), q, m, int(nb_points),
         float(stdv), int(bool(imls)), START_LEVEL, sdf, grad, st)
    return [sdf, grad] if compute_grad else [sdf]
