"""CPU suite: the N>1 host logic (cloud sharding + flat-bucket gradient all-reduce) with
world_size 2 over gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from equi_articulated_pose_b200 import dataparallel as dp


def test_shard_range_partitions_everything():
    for total in (1, 7, 8, 16, 61):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = dp.shard_range(total, r, world)
                seen += list(range(lo, hi))
            assert seen == list(range(total))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, lr, w = dp.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)                                   # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
    bucket = dp.FlatGradBucket(net.parameters())
    data = torch.arange(8 * 5, dtype=torch.float32).view(8, 5) / 40.0
    lo, hi = dp.shard_range(8, rank, world)
    bucket.zero_()
    net(data[lo:hi]).square().mean().backward()
    bucket.all_reduce_mean()
    q.put((rank, bucket.flat.clone(), [p.grad.data_ptr() == bucket.flat[o:o + 1].data_ptr() for p, o in
                                       zip(bucket.params, _offsets(bucket.params))]))
    dist.barrier()
    dist.destroy_process_group()


def _offsets(params):
    off, out = 0, []
    for p in params:
        out.append(off)
        off += p.numel()
    return out


def test_flat_bucket_allreduce_matches_full_batch():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
    data = torch.arange(8 * 5, dtype=torch.float32).view(8, 5) / 40.0
    net(data).square().mean().backward()                   # equal shards: mean of shard means == full mean
    want = torch.cat([p.grad.flatten() for p in net.parameters()])
    for rank, flat, views in got:
        assert torch.allclose(flat, want, atol=1e-6), rank
        assert all(views), "gradients must stay views of the flat bucket"
