"""Seeded synthetic oriented point clouds shared by the tests."""
import numpy as np


def sphere(n, radius=0.35, noise=0.002, seed=0, centre=(0.0, 0.0, 0.0)):
    rng = np.random.default_rng(seed)
    p = rng.normal(size=(n, 3))
    p /= np.linalg.norm(p, axis=1, keepdims=True)
    xyz = p * radius + rng.normal(size=(n, 3)) * noise + np.asarray(centre)
    return xyz.astype(np.float32), p.astype(np.float32)


def shapenet_like(n=3000, noise=0.005, seed=2):
    """cfg2: sphere r=.35 U torus R=.3 r=.1 U box .5^3 surfaces (area weighted), noisy."""
    rng = np.random.default_rng(seed)
    areas = np.array([4 * np.pi * 0.35 ** 2, 4 * np.pi ** 2 * 0.3 * 0.1, 6 * 0.25])
    cnt = rng.multinomial(n, areas / areas.sum())
    pts, nrm = [], []
    p = rng.normal(size=(cnt[0], 3)); p /= np.linalg.norm(p, axis=1, keepdims=True)
    pts.append(p * 0.35); nrm.append(p)
    u, v = rng.uniform(0, 2 * np.pi, cnt[1]), rng.uniform(0, 2 * np.pi, cnt[1])
    pts.append(np.stack([(0.3 + 0.1 * np.cos(v)) * np.cos(u), (0.3 + 0.1 * np.cos(v)) * np.sin(u), 0.1 * np.sin(v)], 1))
    nrm.append(np.stack([np.cos(v) * np.cos(u), np.cos(v) * np.sin(u), np.sin(v)], 1))
    face = rng.integers(0, 6, cnt[2]); ab = rng.uniform(-0.25, 0.25, (cnt[2], 2))
    q = np.zeros((cnt[2], 3)); m = np.zeros((cnt[2], 3))
    for f in range(6):
        s = face == f; ax = f // 2; sg = 1.0 if f % 2 else -1.0
        q[s, ax] = 0.25 * sg; q[s, (ax + 1) % 3] = ab[s, 0]; q[s, (ax + 2) % 3] = ab[s, 1]; m[s, ax] = sg
    pts.append(q); nrm.append(m)
    xyz = np.concatenate(pts) + rng.normal(size=(n, 3)) * noise
    return xyz.astype(np.float32), np.concatenate(nrm).astype(np.float32)


def offset_blob(n, seed=5, scale=3.0, shift=(-17.3, 41.9, -5.25)):
    """random smooth blob far from the origin with negative coordinates (range / sign tests)."""
    rng = np.random.default_rng(seed)
    p = rng.normal(size=(n, 3)); p /= np.linalg.norm(p, axis=1, keepdims=True)
    r = 1.0 + 0.2 * np.sin(3 * p[:, 0]) * np.cos(2 * p[:, 1])
    return (p * r[:, None] * scale + np.asarray(shift)).astype(np.float32), p.astype(np.float32)


def read_ply_xyz_normal(path):
    """minimal reader for binary little-endian PLY files whose vertex element starts with float32
    x y z nx ny nz (the layout of the reference's assets/bunny.ply, examples/common.py:19-22)."""
    with open(path, "rb") as f:
        raw = f.read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    header = raw[:end].decode("ascii").splitlines()
    assert "format binary_little_endian 1.0" in header
    n = next(int(l.split()[2]) for l in header if l.startswith("element vertex"))
    props = []
    for l in header[header.index(next(h for h in header if h.startswith("element vertex"))) + 1:]:
        if not l.startswith("property"):
            break
        props.append(l.split())
    assert [p[2] for p in props[:6]] == ["x", "y", "z", "nx", "ny", "nz"] and all(p[1] == "float" for p in props[:6])
    size = {"float": 4, "uchar": 1, "double": 8, "int": 4, "uint": 4}
    stride = sum(size[p[1]] for p in props)
    body = np.frombuffer(raw, dtype=np.uint8, count=n * stride, offset=end).reshape(n, stride)
    v = body[:, :24].copy().view(np.float32).reshape(n, 6)
    return np.ascontiguousarray(v[:, :3]), np.ascontiguousarray(v[:, 3:6])


def bunny():
    """BASELINE.json configs[0]: assets/bunny.ply as shipped by the reference (10 000 oriented points)."""
    import os
    return read_ply_xyz_normal(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bunny.ply"))
