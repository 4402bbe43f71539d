"""Headless twin of the reference's examples/recons_simple.py:16-29 -- same three nksr calls
(`Reconstructor(device)`, `.reconstruct(xyz, normal, detail_level=1.0)`, `.extract_dual_mesh(mise_iter=1)`)
but input is a binary/ASCII PLY path (e.g. the reference's assets/bunny.ply) or a synthetic sphere, and the
result is written as OBJ instead of opening a viewer.

    python examples/recons_simple.py [cloud.ply] [out.obj]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nksr  # noqa: E402  (alias package -> nksr_b200)


def read_ply_xyz_normal(path):
    """minimal PLY reader: float32 x y z nx ny nz [+ ignored props], binary_little_endian or ascii."""
    with open(path, "rb") as f:
        header, fmt, n, props = [], None, 0, []
        while True:
            line = f.readline().decode("ascii", "ignore").strip()
            header.append(line)
            if line.startswith("format"):
                fmt = line.split()[1]
            elif line.startswith("element vertex"):
                n = int(line.split()[2])
            elif line.startswith("property") and len(props) < 64 and "list" not in line:
                props.append((line.split()[1], line.split()[2]))
            elif line == "end_header":
                break
        names = [p[1] for p in props]
        if fmt == "ascii":
            data = np.loadtxt(f, max_rows=n, dtype=np.float64)
            cols = {nm: data[:, i] for i, nm in enumerate(names)}
        else:
            tmap = {"float": "<f4", "float32": "<f4", "double": "<f8", "uchar": "u1", "uint8": "u1", "int": "<i4"}
            dt = np.dtype([(nm, tmap[t]) for t, nm in props])
            data = np.frombuffer(f.read(n * dt.itemsize), dtype=dt, count=n)
            cols = {nm: data[nm] for nm in names}
    xyz = np.stack([cols["x"], cols["y"], cols["z"]], 1).astype(np.float32)
    nrm = np.stack([cols["nx"], cols["ny"], cols["nz"]], 1).astype(np.float32)
    return xyz, nrm


if __name__ == "__main__":
    device = torch.device("cuda:0")
    if len(sys.argv) > 1 and os.path.exists(sys.argv[1]):
        xyz, nrm = read_ply_xyz_normal(sys.argv[1])
    else:
        rng = np.random.default_rng(0)
        nrm = rng.normal(size=(30000, 3)).astype(np.float32)
        nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        xyz = (nrm * 0.4 + rng.normal(size=nrm.shape) * 0.002).astype(np.float32)
    input_xyz = torch.from_numpy(xyz).float().to(device)
    input_normal = torch.from_numpy(nrm).float().to(device)

    reconstructor = nksr.Reconstructor(device)
    field = reconstructor.reconstruct(input_xyz, input_normal, detail_level=1.0)
    mesh = field.extract_dual_mesh(mise_iter=1)

    out = sys.argv[2] if len(sys.argv) > 2 else "recons_simple.obj"
    v, f = mesh.v.cpu().numpy(), mesh.f.cpu().numpy() + 1
    with open(out, "w") as fh:
        fh.writelines(f"v {a:.6f} {b:.6f} {c:.6f}\n" for a, b, c in v)
        fh.writelines(f"f {a} {b} {c}\n" for a, b, c in f)
    print(f"{xyz.shape[0]} points -> {v.shape[0]} vertices, {f.shape[0]} faces -> {out}; {reconstructor.last_stats}")
