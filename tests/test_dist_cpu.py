"""N>1 path on CPU: two gloo ranks shard a batch exactly like one process would and gather it back."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "q-diffusion_b200")]
    from qdiff_b200 import dist as qdist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, c = qdist.shard_like_single_process((8, 4, 8, 8), seed=42, rank=rank, world=world, extra_shapes=[(8, 7, 16)])
    # stand-in for the per-rank trajectory: any per-sample function commutes with the sharding
    y = x * 2.0 + c.mean(dim=(1, 2))[:, None, None, None]
    full = qdist.gather_latents(y, world)
    if rank == 0:
        q.put(full)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process():
    from qdiff_b200 import dist as qdist
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, c = qdist.shard_like_single_process((8, 4, 8, 8), seed=42, rank=0, world=1, extra_shapes=[(8, 7, 16)])
    ref = x * 2.0 + c.mean(dim=(1, 2))[:, None, None, None]
    assert torch.equal(got, ref)


def test_shard_bounds():
    from qdiff_b200 import dist as qdist
    assert [qdist.shard_bounds(64, r, 8) for r in (0, 7)] == [(0, 8), (56, 64)]
    try:
        qdist.shard_bounds(10, 0, 4)
        raise AssertionError("expected ValueError")
    except ValueError:
        pass
