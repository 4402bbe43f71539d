"""One GLOBAL solve sharded over several GPUs (mapping B; north_star:
"NCCL-over-NVLink only for the CG dot-product/norm allreduce and halo exchange at chunk boundaries").

The cloud is cut into slabs along its longest axis.  Rank r keeps the points of its slab plus a
halo of `halo_voxels` coarsest voxels on either side, builds the hierarchy, features, kernel rows
and Gram rows of that region with the ordinary single-GPU kernels, and OWNS the unknowns whose
voxel centre lies inside its slab.  Because every ingredient of a Gram row is a function of the
points within a few coarsest voxels, the rows of owned unknowns are bit-for-bit the rows of the
single-GPU system; the halo unknowns only serve as columns.  Conjugate gradients then run on the
union of the owned rows:

    per iteration:  halo exchange of p (neighbour send/recv of the boundary entries)
                    SpMV on the local CSR (nksr_spmv, the same hand-written kernel)
                    two fp64 dot products, each ONE all-reduce of a device scalar

The vector updates of this driver are torch ops (a handful of n-vector passes next to an
8·nnz-byte SpMV); nothing is synchronised with the host except the residual test.
Works with `nccl` on GPUs; the key matching / ownership logic is plain tensor code tested on gloo.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import call, stream_ptr
from .fields import KernelField, LayerField
from .svh import SparseFeatureHierarchy, SparseIndexGrid


def slab_bounds(coord: torch.Tensor, world: int, quantum: float) -> List[float]:
    """world+1 increasing bounds along one axis, interior ones at point-count quantiles snapped to
    multiples of `quantum` (the coarsest voxel size): no voxel centre of any level lies on a bound."""
    qs = torch.quantile(coord.double().cpu()[:: max(1, coord.numel() // 2_000_000)],
                        torch.linspace(0, 1, world + 1, dtype=torch.float64)[1:-1]) if world > 1 else coord.new_zeros(0)
    inner = [round(float(q) / quantum) * quantum for q in qs]
    for i in range(1, len(inner)):                         # strictly increasing
        inner[i] = max(inner[i], inner[i - 1] + quantum)
    return [-float("inf")] + inner + [float("inf")]


def owner_of(coord: torch.Tensor, bounds: List[float]) -> torch.Tensor:
    """rank owning a coordinate: number of interior bounds <= coord."""
    inner = torch.tensor(bounds[1:-1], dtype=torch.float64, device=coord.device)
    if inner.numel() == 0:
        return torch.zeros(coord.shape[0], dtype=torch.long, device=coord.device)
    return torch.searchsorted(inner, coord.double().contiguous(), right=True)


class HaloPlan:
    """Who sends which unknowns to whom.  send_idx[r]: my owned unknowns rank r needs;
    recv_idx[r]: my halo unknowns owned by rank r (same order on both sides)."""

    def __init__(self, send_idx, recv_idx, group=None):
        self.send_idx, self.recv_idx, self.group = send_idx, recv_idx, group

    def exchange(self, vec: torch.Tensor):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return vec
        rank = dist.get_rank(self.group)
        ops, recv_bufs = [], {}
        for r, idx in enumerate(self.send_idx):
            if r != rank and idx.numel():
                ops.append(dist.P2POp(dist.isend, vec[idx].contiguous(), r, self.group))
        for r, idx in enumerate(self.recv_idx):
            if r != rank and idx.numel():
                recv_bufs[r] = torch.empty(idx.numel(), dtype=vec.dtype, device=vec.device)
                ops.append(dist.P2POp(dist.irecv, recv_bufs[r], r, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for r, buf in recv_bufs.items():
            vec[self.recv_idx[r]] = buf
        return vec


def build_halo_plan(level_keys: List[torch.Tensor], owner: List[torch.Tensor], offsets: List[int], group=None):
    """level_keys[l]: sorted Morton keys of my local voxels; owner[l]: owning rank of each.
    Voxels are matched across ranks by (level, key) -- bit-exact keys make this an integer join."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = level_keys[0].device
    need = [[None] * len(level_keys) for _ in range(world)]          # need[r][l] = keys I want from r
    recv_idx = [[] for _ in range(world)]
    for l, (keys, own) in enumerate(zip(level_keys, owner)):
        for r in range(world):
            if r == rank:
                continue
            sel = torch.nonzero(own == r).reshape(-1)
            need[r][l] = keys[sel].cpu()
            recv_idx[r].append(sel + offsets[l])
    recv_idx = [torch.cat(v) if v else torch.zeros(0, dtype=torch.long, device=dev) for v in recv_idx]
    if world == 1:
        return HaloPlan([torch.zeros(0, dtype=torch.long, device=dev)], recv_idx, group)
    gathered = [None] * world
    dist.all_gather_object(gathered, need, group=group)                # gathered[s][r][l]: keys s wants from r
    send_idx = []
    for s in range(world):
        parts = []
        if s != rank:
            for l, keys in enumerate(level_keys):
                want = gathered[s][rank][l]
                if want is None or want.numel() == 0:
                    continue
                want = want.to(dev)
                pos = torch.searchsorted(keys, want).clamp(max=max(keys.numel() - 1, 0))
                if keys.numel() == 0 or not bool((keys[pos] == want).all()):
                    raise _lib.NksrError(f"rank {s} asks rank {rank} for voxels it does not hold on level {l}: "
                                         "halo too thin for this hierarchy")
                parts.append(pos + offsets[l])
        send_idx.append(torch.cat(parts) if parts else torch.zeros(0, dtype=torch.long, device=dev))
    return HaloPlan(send_idx, recv_idx, group)


def _gsum(t: torch.Tensor, group) -> torch.Tensor:
    s = t.sum(dtype=torch.float64).reshape(1)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(s, group=group)
    return s


def pcg_distributed(sysm, mask: torch.Tensor, plan: HaloPlan, tol: float, max_iter: int, check_every: int = 1,
                    group=None):
    """Jacobi-PCG on the rows selected by `mask` (1 = owned) of every rank's local CSR system."""
    n, dev = sysm.n, sysm.rhs.device
    st = stream_ptr(dev)
    dinv = torch.where(sysm.diag > 0, 1.0 / sysm.diag.clamp(min=1e-30), torch.zeros_like(sysm.diag)) * mask
    b = sysm.rhs * mask
    x = torch.zeros(n, dtype=torch.float32, device=dev)
    r = b.clone()
    z = r * dinv
    p = z.clone()
    plan.exchange(p)
    ap = torch.empty_like(p)
    rz = _gsum(r.double() * z.double(), group)
    bb = float(_gsum(b.double() * b.double(), group).item())
    if not bb > 0:
        return x, 0, 0.0
    it, rr = 0, bb
    while it < max_iter:
        call("nksr_spmv", sysm.rowptr, sysm.col, sysm.val, p, ap, n, st)
        pap = _gsum(p.double() * ap.double() * mask, group)
        alpha = (rz / pap).float()
        x.add_(p * alpha * mask)
        r.sub_(ap * alpha * mask)
        z = r * dinv
        rz_new = _gsum(r.double() * z.double(), group)
        beta = (rz_new / rz).float()
        rz = rz_new
        p = z + beta * p * mask
        plan.exchange(p)
        it += 1
        if it % check_every == 0 or it == max_iter:
            rr = float(_gsum(r.double() * r.double(), group).item())
            if not rr == rr or rr <= tol * tol * bb:
                break
    plan.exchange(x)                                           # halo coefficients for evaluation / meshing
    return x, it, (rr / bb) ** 0.5


def reconstruct_global(reconstructor, xyz: torch.Tensor, normal: torch.Tensor, voxel_size: float,
                       halo_voxels: int = 8, axis: Optional[int] = None, approx_kernel_grad: bool = False,
                       solver_tol: float = 1e-5, solver_max_iter: int = 2000, group=None):
    """All ranks hold the same oriented cloud; each reconstructs and owns one slab of ONE global
    system.  Returns a KernelField over this rank's slab+halo region with attributes
    `.owned` (per-unknown bool), `.owned_cells` (level-0 mask for meshing) and `.solve_info`."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = reconstructor.device
    xyz = xyz.detach().to(dev, torch.float32).contiguous()
    normal = normal.detach().to(dev, torch.float32).contiguous()
    L = reconstructor.tree_depth
    w_top = float(voxel_size) * (2 ** (L - 1))
    if axis is None:
        axis = int(torch.argmax(xyz.max(dim=0).values - xyz.min(dim=0).values).item())
    bounds = slab_bounds(xyz[:, axis], world, w_top)
    lo, hi = bounds[rank], bounds[rank + 1]
    H = halo_voxels * w_top
    c = xyz[:, axis]
    local = (c >= lo - H) & (c < hi + H)
    lx, ln = xyz[local].contiguous(), normal[local].contiguous()
    n_points_global = int(xyz.shape[0])                                   # every rank sees the whole cloud

    svh = SparseFeatureHierarchy(voxel_size, L, dev).build_point_splatting(lx)
    net = reconstructor.network
    enc = net.encoder(lx, ln, svh, 0)
    feats, dec_svh, _ = net.unet(enc, svh, adaptive_depth=reconstructor.adaptive_depth)
    field = KernelField(dec_svh, net.interpolators, feats.basis_features, approx_kernel_grad)
    ad = min(reconstructor.adaptive_depth, L)
    offs = svh.offsets
    # ownership of every unknown / normal location by voxel-centre coordinate (exact: integer ijk)
    owner, centres = [], []
    for l in range(L):
        g = SparseIndexGrid(svh, l)
        ijk = g.active_grid_coords()
        cen = (ijk[:, axis].double() + 0.5) * (float(voxel_size) * (2 ** l))
        owner.append(owner_of(cen, bounds))
        centres.append(g.grid_to_world(ijk))
    owned = torch.cat([o == rank for o in owner])
    k_global = torch.tensor([float(sum(int((owner[d] == rank).sum().item()) for d in range(ad)))], device=dev,
                            dtype=torch.float64)
    if world > 1:
        dist.all_reduce(k_global, group=group)
    k_global = float(k_global.item())
    normal_xyz = torch.cat([centres[d] for d in range(ad)])
    normal_value = torch.cat([feats.normal_features[d] for d in range(ad)])
    from .reconstructor import NORMAL_WEIGHT, POS_WEIGHT
    sysm = field.assemble(lx, normal_xyz, -normal_value, POS_WEIGHT / n_points_global,
                          NORMAL_WEIGHT / k_global * (float(voxel_size) ** 2), 1.0)
    plan = build_halo_plan(svh.keys, owner, offs, group)
    alpha, iters, relres = pcg_distributed(sysm, owned.float(), plan, solver_tol, solver_max_iter, 1, group)
    field.alpha = alpha
    field.owned = owned
    field.owned_cells = owner[0] == rank
    field.solve_info = {"iterations": iters, "relative_residual": relres, "n": sysm.n, "nnz": sysm.nnz,
                        "n_owned": int(owned.sum().item()), "halo_recv": int(sum(i.numel() for i in plan.recv_idx)),
                        "slab": (lo, hi), "axis": axis}
    field.set_mask_field(LayerField(dec_svh, ad))
    return field


def extract_global_mesh(field, mise_iter: int = 0, grid_upsample: int = 1, group=None):
    """Every rank meshes the dual cells whose min-corner voxel it owns; pieces are gathered on rank 0."""
    from .dist import gather_mesh
    mesh = field.extract_dual_mesh(grid_upsample=grid_upsample, mise_iter=mise_iter, cell_filter=field.owned_cells)
    v, f = gather_mesh(mesh.v, mesh.f, 0, group)
    return None if v is None else SimpleNamespace(v=v, f=f, c=None)
