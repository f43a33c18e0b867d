"""World-size-2 gloo tests (CPU) of the data-parallel host logic (dist.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from pdivgnn_b200 import dist as pd
    r, w, _ = pd.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)
    flat = torch.randn(167299)
    mine = flat.clone()
    pd.allreduce_flat_(flat)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    ref = sum(gathered) / world
    ok_mean = torch.allclose(flat, ref, atol=1e-7)
    # identical on every rank
    others = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(others, flat)
    ok_same = all(torch.equal(o, flat) for o in others)
    # parameter broadcast
    lin = torch.nn.Linear(4, 4)
    pd.broadcast_parameters(lin)
    ws = [torch.empty_like(lin.weight.data) for _ in range(world)]
    dist.all_gather(ws, lin.weight.data)
    ok_bc = all(torch.equal(x, ws[0]) for x in ws)
    out[rank] = (ok_mean, ok_same, ok_bc, pd.shard_indices(10, rank, world))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_and_sharding_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0][:3] == (True, True, True) and out[1][:3] == (True, True, True)
    assert out[0][3] == [0, 2, 4, 6, 8] and out[1][3] == [1, 3, 5, 7, 9]
    assert sorted(out[0][3] + out[1][3]) == list(range(10))


def test_single_process_is_identity():
    from pdivgnn_b200 import dist as pd
    f = torch.arange(8.0)
    assert torch.equal(pd.allreduce_flat_(f.clone()), f)
    assert pd.shard_indices(5, 0, 1) == [0, 1, 2, 3, 4]
