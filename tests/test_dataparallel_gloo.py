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


def _spawn(target, world=2, attempts=2):
    """Run `target(rank, world, port, queue)` in `world` spawned processes and return their results; a failed rendezvous
    (the probed port taken in the meantime, a loaded host) is retried once on a fresh port."""
    import queue as _queue
    last = None
    for _ in range(attempts):
        port = _free_port()
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=target, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        try:
            got = [q.get(timeout=180) for _ in range(world)]
            for p in procs:
                p.join(timeout=60)
            if all(p.exitcode == 0 for p in procs):
                return got
            last = f"exit codes {[p.exitcode for p in procs]}"
        except _queue.Empty:
            last = "no result within 180 s"
        for p in procs:
            if p.is_alive():
                p.terminate()
            p.join(timeout=10)
    raise AssertionError(f"workers failed on {attempts} attempts: {last}")


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
    got = _spawn(_worker)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
    data = torch.arange(8 * 5, dtype=torch.float32).view(8, 5) / 40.0
    net(data).square().mean().backward()                   # equal shards: mean of shard means == full mean
    want = torch.cat([p.grad.flatten() for p in net.parameters()])
    for rank, flat, views in got:
        assert torch.allclose(flat, want, atol=1e-6), rank
        assert all(views), "gradients must stay views of the flat bucket"


# ---------------------------------------------------------------------------------- SyncBatchNorm protocol
def _syncbn_worker(rank, world, port, q):
    """The two-phase SyncBatchNorm protocol of ops.NormActFn (sums -> all-reduce of [2C sums | row count] -> finalize /
    apply), with the four C-ABI phases restated in torch fp64 on the CPU; unequal shards exercise the count that rides
    behind the sums."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dp.init_from_env(backend="gloo")
    from equi_articulated_pose_b200 import ops
    assert ops._sync_world(True) == world and ops._sync_world(None) == 1
    g = torch.Generator().manual_seed(7)
    c, slope, eps = 6, 0.01, 1e-5
    x_all = torch.randn(8, c, generator=g, dtype=torch.float64) * 2 + 0.5
    gy_all = torch.randn(8, c, generator=g, dtype=torch.float64)
    gamma = torch.rand(c, generator=g, dtype=torch.float64) + 0.5
    beta = torch.randn(c, generator=g, dtype=torch.float64) * 0.1
    lo, hi = (0, 5) if rank == 0 else (5, 8)
    x, gy = x_all[lo:hi], gy_all[lo:hi]
    rows = x.shape[0]
    # vgtkb_norm_sums
    scratch = torch.cat([x.sum(0), (x * x).sum(0), torch.tensor([float(rows)], dtype=torch.float64)])
    ops._all_reduce_sums(scratch, True)
    # vgtkb_norm_finalize (total_rows = 0: count from the buffer)
    n = scratch[2 * c]
    mean = scratch[:c] / n
    inv = torch.rsqrt((scratch[c:2 * c] / n - mean * mean).clamp_min(0) + eps)
    xh = (x - mean) * inv
    pre = xh * gamma + beta
    y = torch.where(pre > 0, pre, pre * slope)
    # vgtkb_norm_bwd_sums (affine gradients = local sums)
    dyp = torch.where(pre > 0, gy, gy * slope)
    s = torch.cat([dyp.sum(0), (dyp * xh).sum(0), torch.tensor([float(rows)], dtype=torch.float64)])
    gbeta, ggamma = s[:c].clone(), s[c:2 * c].clone()
    ops._all_reduce_sums(s, True)
    # vgtkb_norm_bwd_apply
    gx = gamma * inv * (dyp - s[:c] / s[2 * c] - xh * s[c:2 * c] / s[2 * c])
    q.put((rank, lo, hi, y, gx, ggamma, gbeta, float(n)))
    dist.barrier()
    dist.destroy_process_group()


def test_syncbn_protocol_matches_full_batch_batchnorm():
    res = sorted(_spawn(_syncbn_worker), key=lambda t: t[0])
    g = torch.Generator().manual_seed(7)
    c, slope, eps = 6, 0.01, 1e-5
    x_all = (torch.randn(8, c, generator=g, dtype=torch.float64) * 2 + 0.5).requires_grad_(True)
    gy_all = torch.randn(8, c, generator=g, dtype=torch.float64)
    gamma = (torch.rand(c, generator=g, dtype=torch.float64) + 0.5).requires_grad_(True)
    beta = (torch.randn(c, generator=g, dtype=torch.float64) * 0.1).requires_grad_(True)
    y_ref = torch.nn.functional.leaky_relu(
        torch.nn.functional.batch_norm(x_all, None, None, gamma, beta, training=True, eps=eps), slope)
    y_ref.backward(gy_all)
    y = torch.cat([r[3] for r in res])
    gx = torch.cat([r[4] for r in res])
    assert res[0][7] == 8.0 and res[1][7] == 8.0
    assert torch.allclose(y, y_ref.detach(), atol=1e-12)
    assert torch.allclose(gx, x_all.grad, atol=1e-12)
    assert torch.allclose(res[0][5] + res[1][5], gamma.grad, atol=1e-12)      # local sums add up to the global gradient
    assert torch.allclose(res[0][6] + res[1][6], beta.grad, atol=1e-12)


def test_flat_grad_bucket_single_process_semantics():
    """zero_() drops the gradients, backward writes fresh tensors, collect()/flat packs them with one concatenation and makes
    every .grad a view of the flat buffer; accumulation without zero_() happens in place; unused parameters read as zero."""
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3), torch.nn.Linear(3, 2))
    bucket = dp.FlatGradBucket(net.parameters())
    opt = torch.optim.Adam(bucket.params, lr=1e-2)
    x = torch.randn(8, 5)
    for _ in range(3):
        bucket.zero_()
        assert all(p.grad is None for p in net.parameters())
        net(x).square().mean().backward()
        want = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
        bucket.all_reduce_mean()                                  # world size 1: collect only
        assert torch.equal(want, bucket.flat)
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket._views))
        opt.step()
    before = bucket.flat.clone()
    net(x).square().mean().backward()                             # no zero_(): accumulates into the views in place
    assert not torch.equal(before, bucket.flat) and all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(bucket.params, bucket._views))
    bucket.zero_()
    net[0](x).sum().backward()                                    # only the first layer is reached
    f = bucket.flat
    assert float(f[:35].abs().max()) > 0 and float(f[42:].abs().max()) == 0.0
    bucket.zero_()
    net(x).square().mean().backward()
    want = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert torch.equal(bucket.flat, want)                         # `flat` collects on first access
