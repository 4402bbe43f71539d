"""ctypes wrapper of oracle/nksr_oracle_cpu.cpp (C++/OpenMP CPU restatement) -- TEST / BASELINE
INFRASTRUCTURE ONLY, parity unpinned like oracle/nksr_oracle.py.  Used by tests/test_cpu_port.py (checked
against the numpy oracle) and by bench.py's CPU arm (`cpu_baseline.kind = "port"`)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libnksr_oracle_cpu.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        L.nksr_cpu_svh_build.restype = C.c_void_p
        L.nksr_cpu_svh_build.argtypes = [C.c_void_p, C.c_int64, C.c_float, C.c_int]
        L.nksr_cpu_svh_free.argtypes = [C.c_void_p]
        L.nksr_cpu_svh_count.restype = C.c_int64
        L.nksr_cpu_svh_count.argtypes = [C.c_void_p, C.c_int]
        L.nksr_cpu_svh_keys.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.nksr_cpu_svh_centers.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.nksr_cpu_build_system.restype = C.c_void_p
        L.nksr_cpu_build_system.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                            C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_int]
        L.nksr_cpu_system_free.argtypes = [C.c_void_p]
        L.nksr_cpu_system_n.restype = C.c_int64
        L.nksr_cpu_system_n.argtypes = [C.c_void_p]
        L.nksr_cpu_system_nnz.restype = C.c_int64
        L.nksr_cpu_system_nnz.argtypes = [C.c_void_p]
        L.nksr_cpu_system_copy.argtypes = [C.c_void_p] * 5
        L.nksr_cpu_pcg.restype = C.c_int
        L.nksr_cpu_pcg.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
        L.nksr_cpu_threads.restype = C.c_int
        L.nksr_cpu_evaluate.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.nksr_cpu_svh_from_keys.restype = C.c_void_p
        L.nksr_cpu_svh_from_keys.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_float, C.c_int]
        L.nksr_cpu_locate.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.nksr_cpu_structural_counts.argtypes = [C.c_void_p, C.c_void_p]
        L.nksr_cpu_structural_row.restype = C.c_int64
        L.nksr_cpu_structural_row.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.nksr_cpu_nbr27.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.nksr_cpu_pool27.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class CpuSvh:
    def __init__(self, xyz: np.ndarray, voxel_size: float, depth: int, keys=None):
        self.depth = depth
        self.voxel_size = float(np.float32(voxel_size))
        if keys is not None:                       # adopt sorted unique keys per level (pruned hierarchies)
            ks = [np.ascontiguousarray(k, np.int64) for k in keys]
            ptrs = (C.c_void_p * depth)(*[k.ctypes.data for k in ks])
            cnt = np.array([k.shape[0] for k in ks], np.int64)
            self.h = lib().nksr_cpu_svh_from_keys(ptrs, _p(cnt), np.float32(voxel_size), depth)
            return
        self.xyz = np.ascontiguousarray(xyz, np.float32)
        self.h = lib().nksr_cpu_svh_build(_p(self.xyz), self.xyz.shape[0], np.float32(voxel_size), depth)

    def offsets(self):
        return np.concatenate([[0], np.cumsum([self.n(l) for l in range(self.depth)])]).astype(np.int64)

    def locate(self, xyz):
        q = np.ascontiguousarray(xyz, np.float32)
        out = np.empty((self.depth, q.shape[0]), np.int32)
        lib().nksr_cpu_locate(self.h, _p(q), q.shape[0], _p(out))
        return out

    def nbr27(self, l):
        """(n_l, 27) int32 neighbour table, -1 = inactive (slot order of nksr_oracle._OFF27)"""
        out = np.empty((self.n(l), 27), np.int32)
        lib().nksr_cpu_nbr27(self.h, int(l), _p(out))
        return out

    def pool27(self, l, acc):
        """sum of `acc` (n_l, C) float64 over the 27-neighbourhood of every voxel, slots added in table order"""
        acc = np.ascontiguousarray(acc, np.float64)
        nb = self.nbr27(l)
        out = np.empty_like(acc)
        lib().nksr_cpu_pool27(_p(nb), acc.shape[0], _p(acc), acc.shape[1], _p(out))
        return out

    def structural_counts(self):
        """stored entries per row of the SPEC S6 pattern (own + transposed)."""
        out = np.empty(int(self.offsets()[-1]), np.int32)
        lib().nksr_cpu_structural_counts(self.h, _p(out))
        return out

    def structural_row(self, row, capacity):
        buf = np.empty(int(capacity), np.int32)
        m = lib().nksr_cpu_structural_row(self.h, int(row), _p(buf))
        return buf[:m].copy()

    def evaluate(self, feats, alpha, xyz, grad=False, approx=False):
        """f (M,) [and grad (M,3)] in float64 at the queries (SPEC S3/S4)."""
        feats = [np.ascontiguousarray(f, np.float32) for f in feats]
        ptrs = (C.c_void_p * len(feats))(*[f.ctypes.data for f in feats])
        q = np.ascontiguousarray(xyz, np.float32)
        a = np.ascontiguousarray(alpha, np.float32)
        f = np.empty(q.shape[0], np.float64)
        g = np.empty((q.shape[0], 3), np.float64) if grad else None
        lib().nksr_cpu_evaluate(self.h, ptrs, feats[0].shape[1], _p(a), _p(q), q.shape[0], int(grad), int(approx),
                                _p(f), _p(g) if grad else None)
        return (f, g) if grad else f

    def __del__(self):
        if getattr(self, "h", None):
            lib().nksr_cpu_svh_free(self.h)
            self.h = None

    def n(self, l):
        return int(lib().nksr_cpu_svh_count(self.h, l))

    def keys(self, l):
        out = np.empty(self.n(l), np.int64)
        lib().nksr_cpu_svh_keys(self.h, l, _p(out))
        return out

    def centers(self, l):
        out = np.empty((self.n(l), 3), np.float32)
        lib().nksr_cpu_svh_centers(self.h, l, _p(out))
        return out


class CpuSystem:
    def __init__(self, svh: CpuSvh, feats, pos_xyz, nrm_xyz, nrm_val, w_pos, w_nrm, w_reg, approx=False):
        feats = [np.ascontiguousarray(f, np.float32) for f in feats]
        ptrs = (C.c_void_p * len(feats))(*[f.ctypes.data for f in feats])
        pos = np.ascontiguousarray(pos_xyz, np.float32)
        nx = np.ascontiguousarray(nrm_xyz, np.float32).reshape(-1, 3)
        nv = np.ascontiguousarray(nrm_val, np.float32).reshape(-1, 3)
        self.h = lib().nksr_cpu_build_system(svh.h, ptrs, feats[0].shape[1], _p(pos), pos.shape[0], _p(nx), _p(nv),
                                             nx.shape[0], w_pos, w_nrm, w_reg, int(approx))
        self.n = int(lib().nksr_cpu_system_n(self.h))
        self.nnz = int(lib().nksr_cpu_system_nnz(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            lib().nksr_cpu_system_free(self.h)
            self.h = None

    def to_scipy(self):
        rowptr = np.empty(self.n + 1, np.int64)
        col = np.empty(self.nnz, np.int32)
        val = np.empty(self.nnz, np.float32)
        rhs = np.empty(self.n, np.float32)
        lib().nksr_cpu_system_copy(self.h, _p(rowptr), _p(col), _p(val), _p(rhs))
        return sp.csr_matrix((val.astype(np.float64), col, rowptr), shape=(self.n, self.n)), rhs

    def pcg(self, tol=1e-5, max_iter=2000):
        x = np.empty(self.n, np.float32)
        res = C.c_double(0.0)
        it = lib().nksr_cpu_pcg(self.h, tol, max_iter, _p(x), C.byref(res))
        return x, int(it), float(res.value)


def reconstruct(xyz: np.ndarray, voxel_size: float, depth: int = 4, channels: int = 4, approx: bool = True,
                tol: float = 1e-4):
    """The CPU twin of one bench step (same constraint wiring as bench.py's numpy arm): hierarchy, constant
    features, normal constraints at the centres of the two finest levels, assembly, PCG."""
    svh = CpuSvh(xyz, voxel_size, depth)
    feats = [np.full((svh.n(l), channels), 0.5, np.float32) for l in range(depth)]
    ad = min(2, depth)
    nxyz = np.concatenate([svh.centers(d) for d in range(ad)])
    nval = np.tile(np.array([[0.0, 0.0, -1.0]], np.float32), (nxyz.shape[0], 1))
    sysm = CpuSystem(svh, feats, xyz, nxyz, nval, 1e4 / xyz.shape[0], 1e4 / nxyz.shape[0] * voxel_size ** 2, 1.0, approx)
    x, it, res = sysm.pcg(tol, 2000)
    return dict(n=sysm.n, nnz=sysm.nnz, iterations=it, relres=res)


def threads() -> int:
    return int(lib().nksr_cpu_threads())
