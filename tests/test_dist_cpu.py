"""N>1 host logic on CPU (gloo, world_size 2): z-slab sharding of the tiled request, the NCCL-id broadcast
plumbing of Engine.init_comm, and the max-over-ranks timing reduction used by bench.py."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from transfer_em_b200.utils import z_slab_range
    # (a) every rank's z-slab; gathered on rank 0 they must tile [0, nz) exactly
    nz = 29                                                  # 1024^3 request: ceil(1024/36) z tile-layers
    mine = torch.tensor(z_slab_range(nz, rank, world))
    slabs = [torch.zeros(2, dtype=torch.long) for _ in range(world)]
    dist.all_gather(slabs, mine)
    # (b) the broadcast closure EM2EM._init_distributed hands to Engine.init_comm
    def bcast(raw):
        obj = [raw]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]
    ident = bytes(range(128)) if rank == 0 else None
    got = bcast(ident)
    # (c) max-over-ranks of a per-rank elapsed time
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # (d) gradient averaging convention: sum over ranks * 1/world == global-batch mean of per-rank means
    g = torch.full((4,), float(rank + 1))
    dist.all_reduce(g)
    q.put((rank, [tuple(int(v) for v in s) for s in slabs], got, float(t), (g / world).tolist()))
    dist.destroy_process_group()


def test_two_rank_host_logic():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps: p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps: p.join(timeout=60)
    for rank, slabs, got, tmax, gmean in res:
        assert slabs == [(0, 15), (15, 29)]                  # contiguous, disjoint, complete
        assert got == bytes(range(128))                      # every rank holds rank 0's id
        assert tmax == 11.0
        assert gmean == [1.5] * 4


def test_z_slab_ranges_cover_any_split():
    from transfer_em_b200.utils import z_slab_range
    for nz in (1, 3, 8, 29, 64):
        for world in (1, 2, 4, 8):
            r = [z_slab_range(nz, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == nz
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
