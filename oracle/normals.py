"""CPU restatement of the normal-estimation preprocess -- TEST / BASELINE INFRASTRUCTURE ONLY.

Follows the reference's open CPU twin line by line (examples/recons_waymo_cpu.py:21-41), which is what
`nksr.get_estimate_normal_preprocess_fn(64, 85.0)` does on the GPU (examples/recons_waymo.py:36):

    :26  indices, normal = pcu.estimate_point_cloud_normals_knn(xyz, 64)    kNN (k = 64) PCA normals
    :32-33 view_dir = (sensor - xyz) / (|sensor - xyz| + 1e-6)
    :34-36 flip normals with  <view_dir, normal> < 0
    :38-39 keep |cos| > cos(85 deg)

`point_cloud_utils` is a third-party dependency absent from /root/reference (environment.yml pins no
version); its published algorithm is restated here: the k nearest neighbours of a point INCLUDING the point
itself, the 3x3 covariance of those neighbours about their mean, the unit eigenvector of the smallest
eigenvalue.  Parity unpinned (no golden vectors in the reference).  scipy's cKDTree does the search.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree


def knn_indices(xyz: np.ndarray, k: int, workers: int = -1):
    """(N,k) indices and (N,k) distances of the k nearest points (self included), nearest first."""
    tree = cKDTree(np.asarray(xyz, np.float64))
    d, idx = tree.query(np.asarray(xyz, np.float64), k=k, workers=workers)
    return idx.reshape(xyz.shape[0], -1), d.reshape(xyz.shape[0], -1)


def pca_normals(xyz: np.ndarray, idx: np.ndarray):
    """unit eigenvector of the smallest eigenvalue of the neighbourhood covariance (float64);
    also returns the eigenvalues (ascending) so that tests can skip degenerate neighbourhoods."""
    p = np.asarray(xyz, np.float64)[idx]                         # (N,k,3)
    c = p - p.mean(axis=1, keepdims=True)
    cov = np.einsum('nki,nkj->nij', c, c) / idx.shape[1]
    w, v = np.linalg.eigh(cov)
    return v[:, :, 0], w


def estimate_normal_preprocess(xyz: np.ndarray, sensor: np.ndarray, knn: int = 64, max_angle_deg: float = 85.0,
                               workers: int = -1):
    """Returns xyz', normal' (float32) of the kept points, the keep mask, and (normals of ALL points, cos, eigvals)."""
    idx, _ = knn_indices(xyz, min(knn, xyz.shape[0]), workers)
    n, ev = pca_normals(xyz, idx)
    view = np.asarray(sensor, np.float64) - np.asarray(xyz, np.float64)
    view = view / (np.linalg.norm(view, axis=-1, keepdims=True) + 1e-6)
    cos = np.sum(view * n, axis=1)
    n = np.where((cos < 0.0)[:, None], -n, n)
    keep = np.abs(cos) > np.cos(np.deg2rad(max_angle_deg))
    return (np.ascontiguousarray(xyz[keep], np.float32), np.ascontiguousarray(n[keep], np.float32), keep,
            (n, cos, ev))
