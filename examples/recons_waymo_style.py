"""Headless twin of the reference's examples/recons_waymo.py:24-41: sensor-oriented cloud, normal
estimation preprocess, approx_kernel_grad / solver_tol / fused_mode knobs, optional chunking.
Uses the synthetic outdoor scene of bench.py (the Waymo asset is a network download).

    python examples/recons_waymo_style.py [points] [chunk_size]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import nksr  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    chunk = float(sys.argv[2]) if len(sys.argv) > 2 else None
    device = torch.device("cuda:0")
    xyz, sensor = bench.make_cloud("dev_outdoor_1M", 4, 0, points=n)
    reconstructor = nksr.Reconstructor(device)
    reconstructor.chunk_tmp_device = torch.device("cpu")
    field = reconstructor.reconstruct(
        xyz.to(device), sensor=sensor.to(device), detail_level=None,
        approx_kernel_grad=True, solver_tol=1e-4, fused_mode=True,
        chunk_size=chunk,
        preprocess_fn=nksr.get_estimate_normal_preprocess_fn(64, 85.0))
    mesh = field.extract_dual_mesh(mise_iter=1)
    print(f"{n} points -> {mesh.v.shape[0]} vertices, {mesh.f.shape[0]} faces")
