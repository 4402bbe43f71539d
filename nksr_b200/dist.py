"""Multi-GPU plumbing: spatial chunks of one cloud are sharded over the ranks of one node.

The reference reconstructs chunks serially on one GPU (examples/recons_by_chunk.py:27-29,
NKSR-USAGE.md:88-120); chunks are independent units, so here every rank (one process per GPU,
torch.distributed) reconstructs and meshes the chunks it owns and only the final mesh pieces travel
(variable-size gather to rank 0).  No collective sits on the data path of the solve.  Works on
`nccl` (GPU tensors) and on `gloo` (CPU tensors; used by the CPU tests of this host logic).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def assign_chunks(weights: Sequence[float], world: int) -> List[int]:
    """Owner rank of every chunk: longest-processing-time-first greedy on the chunk weights
    (point counts).  Deterministic: ties break towards the lower chunk index / lower rank."""
    order = sorted(range(len(weights)), key=lambda i: (-float(weights[i]), i))
    load = [0.0] * world
    owner = [0] * len(weights)
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += float(weights[i])
    return owner


def max_over_ranks(value: float, device, group=None) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def gather_mesh(v: torch.Tensor, f: torch.Tensor, dst: int = 0, group=None):
    """Variable-size gather of (V,3) float vertices and (T,3) int64 faces to `dst`; face indices are
    rebased.  Returns (v, f) on dst, (None, None) elsewhere."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return v, f
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = torch.tensor([v.shape[0], f.shape[0]], dtype=torch.int64, device=v.device)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    nv = [int(s[0]) for s in all_sizes]
    nf = [int(s[1]) for s in all_sizes]
    vmax, fmax = max(max(nv), 1), max(max(nf), 1)
    vp = torch.zeros((vmax, 3), dtype=torch.float32, device=v.device)
    fp = torch.zeros((fmax, 3), dtype=torch.int64, device=v.device)
    vp[: v.shape[0]] = v
    fp[: f.shape[0]] = f
    vs = [torch.zeros_like(vp) for _ in range(world)] if rank == dst else None
    fs = [torch.zeros_like(fp) for _ in range(world)] if rank == dst else None
    dist.gather(vp, vs, dst=dst, group=group)
    dist.gather(fp, fs, dst=dst, group=group)
    if rank != dst:
        return None, None
    out_v, out_f, off = [], [], 0
    for r in range(world):
        out_v.append(vs[r][: nv[r]])
        out_f.append(fs[r][: nf[r]] + off)
        off += nv[r]
    return torch.cat(out_v), torch.cat(out_f)


def reconstruct_distributed(reconstructor, xyz: torch.Tensor, normal: Optional[torch.Tensor] = None,
                            sensor: Optional[torch.Tensor] = None, chunk_size: float = 51.2, preprocess_fn=None,
                            mise_iter: int = 1, group=None, **solver_kwargs):
    """Every rank holds the full cloud (or at least its own chunks + halo), reconstructs the chunks
    it owns and meshes them; rank 0 receives the merged mesh.  Returns (field_of_local_chunks, mesh|None)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    cidx = torch.floor(xyz / chunk_size).long()
    cores, counts = torch.unique(cidx, dim=0, return_counts=True)
    owner = assign_chunks(counts.tolist(), world)
    from .meshing import DualMesh
    from .reconstructor import DEFAULT_VOXEL_SIZE
    if rank not in owner:                        # more ranks than chunks: contribute an empty piece
        v, f = gather_mesh(torch.zeros((0, 3), device=xyz.device),
                           torch.zeros((0, 3), dtype=torch.int64, device=xyz.device), 0, group)
        return None, (None if v is None else DualMesh(v=v, f=f, c=None))
    # a rank that fails (e.g. none of its chunks holds enough points) must still enter the collectives, or the others
    # hang in the gather: contribute an empty piece, then re-raise
    field, err = None, None
    try:
        field = reconstructor._reconstruct_chunks(
            xyz, normal, sensor, DEFAULT_VOXEL_SIZE, float(chunk_size), preprocess_fn,
            solver_kwargs.get("approx_kernel_grad", False), solver_kwargs.get("solver_tol", 1e-5),
            solver_kwargs.get("fused_mode", True), solver_kwargs.get("solver_max_iter", 2000),
            chunk_filter=lambda k, n: owner[k] == rank)
        mesh = field.extract_dual_mesh(mise_iter=mise_iter)
        lv, lf = mesh.v, mesh.f
    except Exception as e:          # noqa: BLE001 -- re-raised below, after the collective
        err = e
        lv = torch.zeros((0, 3), device=xyz.device)
        lf = torch.zeros((0, 3), dtype=torch.int64, device=xyz.device)
    v, f = gather_mesh(lv, lf, 0, group)
    if err is not None:
        raise err
    return field, (None if v is None else DualMesh(v=v, f=f, c=None))
