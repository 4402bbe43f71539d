"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in nksr_b200/dist.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nksr_b200 import dist as nd
    try:
        # chunk ownership: same answer on every rank, every chunk owned once, balanced
        weights = [100, 5, 60, 40, 7, 90, 3]
        owner = nd.assign_chunks(weights, world)
        gathered = [None] * world
        dist.all_gather_object(gathered, owner)
        assert all(g == owner for g in gathered)
        loads = [sum(w for w, o in zip(weights, owner) if o == r) for r in range(world)]
        assert sum(loads) == sum(weights) and max(loads) - min(loads) <= max(weights)
        # timing reduction = max over ranks
        assert nd.max_over_ranks(10.0 + rank, "cpu") == 10.0 + world - 1
        # variable-size mesh gather with index rebasing (rank r contributes r+2 triangles)
        nv = 3 * (rank + 2)
        v = torch.full((nv, 3), float(rank)) + torch.arange(nv)[:, None] * 0.001
        f = torch.arange(nv, dtype=torch.int64).reshape(-1, 3)
        gv, gf = nd.gather_mesh(v, f, 0)
        if rank == 0:
            assert gv.shape[0] == sum(3 * (r + 2) for r in range(world)) and gf.shape[0] == sum(r + 2 for r in range(world))
            # every face still references three vertices of its own rank
            owner_of_vertex = torch.floor(gv[:, 0] + 1e-6)
            assert (owner_of_vertex[gf].min(dim=1).values == owner_of_vertex[gf].max(dim=1).values).all()
            assert gf.max().item() == gv.shape[0] - 1
        else:
            assert gv is None and gf is None
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_chunk_sharding_logic_world2():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_assign_chunks_is_deterministic_and_complete():
    from nksr_b200.dist import assign_chunks
    w = [5, 5, 5, 1, 9, 2, 2, 8]
    a = assign_chunks(w, 4)
    assert a == assign_chunks(w, 4) and set(a) <= set(range(4)) and len(a) == len(w)
    assert assign_chunks([], 4) == [] and assign_chunks([3.0], 8) == [0]
