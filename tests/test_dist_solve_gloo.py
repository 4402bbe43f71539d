"""world_size-2 gloo tests (CPU tensors) of the global-solve host logic in nksr_b200/dist_solve.py:
slab bounds, ownership, the (level, key) halo join and the neighbour exchange."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nksr_b200 import dist_solve as ds
    try:
        # two levels; "keys" double as coordinates: level-0 key k sits at x = k + 0.5, level-1 at 2k + 1
        bounds = [-float("inf"), 50.0, float("inf")]
        lo0, hi0 = (0, 60) if rank == 0 else (40, 100)                 # local = slab + halo of 10
        keys0 = torch.arange(lo0, hi0, dtype=torch.int64)
        keys1 = torch.arange(lo0 // 2, hi0 // 2, dtype=torch.int64)
        owner = [ds.owner_of(keys0.double() + 0.5, bounds), ds.owner_of(keys1.double() * 2 + 1.0, bounds)]
        offsets = [0, keys0.numel()]
        plan = ds.build_halo_plan([keys0, keys1], owner, offsets)
        owned = torch.cat([o == rank for o in owner])
        # every unknown carries a rank-independent global value; halo entries start poisoned
        truth = torch.cat([keys0.float() * 3.0, 1000.0 + keys1.float() * 7.0])
        vec = torch.where(owned, truth, torch.full_like(truth, -1.0))
        plan.exchange(vec)
        assert torch.equal(vec, truth), (rank, (vec != truth).nonzero().reshape(-1)[:5])
        assert sum(i.numel() for i in plan.recv_idx) == int((~owned).sum())
        # a rim voxel only ONE side holds (a per-rank preprocess may keep a point the owner dropped at the outer edge of
        # the requester's halo): the join leaves it out instead of failing, its entry keeps the caller's value
        if rank == 0:
            k0r = torch.cat([keys0, torch.tensor([1000], dtype=torch.int64)])      # key 1000: owned by rank 1, unknown there
        else:
            k0r = keys0
        own_r = [ds.owner_of(k0r.double() + 0.5, bounds), owner[1]]
        plan_r = ds.build_halo_plan([k0r, keys1], own_r, [0, k0r.numel()])
        truth_r = torch.cat([k0r.float() * 3.0, 1000.0 + keys1.float() * 7.0])
        owned_r = torch.cat([o == rank for o in own_r])
        vec_r = torch.where(owned_r, truth_r, torch.full_like(truth_r, -1.0))
        plan_r.exchange(vec_r)
        if rank == 0:
            assert vec_r[k0r.numel() - 1] == -1.0                                  # untouched
            keep = torch.ones_like(vec_r, dtype=torch.bool); keep[k0r.numel() - 1] = False
            assert torch.equal(vec_r[keep], truth_r[keep])
        else:
            assert torch.equal(vec_r, truth_r)
        # global dot product over owned entries = dot over the union exactly once
        s = ds._gsum(truth.double() * owned, None)
        ref = (torch.arange(0, 100).double() * 3.0).sum() + (1000.0 + torch.arange(0, 50).double() * 7.0).sum()
        assert abs(float(s) - float(ref)) < 1e-9
        # slab bounds: increasing, snapped to the quantum, same on every rank
        coord = torch.linspace(-3.0, 17.0, 10001)
        b = ds.slab_bounds(coord, 4, 0.8)
        assert len(b) == 5 and all(b[i] < b[i + 1] for i in range(4))
        assert all(abs(v / 0.8 - round(v / 0.8)) < 1e-9 for v in b[1:-1])
        got = [None] * world
        dist.all_gather_object(got, b)
        assert got[0] == got[1]
        # point all-to-all: every rank starts with a different share of one cloud and ends with its slab + halo
        g = torch.Generator().manual_seed(7)
        cloud = torch.rand(4000, 3, generator=g) * 100.0
        extra = torch.arange(4000, dtype=torch.float32)[:, None].repeat(1, 3)
        mine = slice(0, 1500) if rank == 0 else slice(1500, 4000)          # unequal shares
        px, pe = ds.route_points(cloud[mine, 0], bounds, 10.0, [cloud[mine], extra[mine]])
        lo_, hi_ = (-1e30, 60.0) if rank == 0 else (40.0, 1e30)
        want = (cloud[:, 0] >= lo_) & (cloud[:, 0] < hi_)
        assert px.shape[0] == int(want.sum()) and torch.equal(torch.sort(pe[:, 0]).values, extra[want, 0])
        assert torch.equal(cloud[pe[:, 0].long()], px)                      # rows stay together
        b2 = ds.slab_bounds(cloud[mine, 0], 2, 5.0)                         # bounds from distributed samples agree
        got = [None] * world
        dist.all_gather_object(got, b2)
        assert got[0] == got[1] and 40.0 <= b2[1] <= 60.0
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        out.put((rank, traceback.format_exc()[-400:]))
    finally:
        dist.destroy_process_group()


def test_halo_join_and_exchange_world2():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
