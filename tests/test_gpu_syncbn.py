"""GPU suite, 2 ranks over NCCL: SyncBatchNorm (the reference converts its model with
nn.SyncBatchNorm.convert_sync_batchnorm, SPConvNets/trainer_unsup_arti_align.py:430).

Two ranks with two clouds each must reproduce ONE process with the four clouds: the BatchNorm statistics span all
ranks, the loss is the mean of the per-rank means, the parameter gradients are averaged by the flat bucket.
Needs two visible devices (skipped otherwise): run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_syncbn.py -m gpu`.
"""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(dev, n_points):
    from equi_articulated_pose_b200 import blocks
    from oracle import so3 as O
    params = O.backbone_params(input_num=n_points)
    net = blocks.SO3Backbone(params)
    net.load_state_dict(O.init_backbone_state(params, seed=0), strict=False)
    return net.to(dev).train()


def _worker(rank, world, port, n_points, peer, q):
    import torch.distributed as dist
    from equi_articulated_pose_b200 import blocks, ops, dataparallel as dp
    from oracle import so3 as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dp.init_from_env(backend="nccl")
    dev = torch.device("cuda", rank)
    peer_ok = None
    if peer:
        # NVLink peer mailboxes: 300 back-to-back exchanges of varying length (slot reuse), rank-order sums
        mb = dp.PeerMailbox(dev)
        ops.set_peer_mailbox(mb)
        peer_ok = True
        for i in range(300):
            n = 1 + (i * 37) % 2049
            buf = (torch.arange(n, dtype=torch.float64, device=dev) + 1) * (rank + 1) * (i + 1)
            mb.all_reduce_sums(buf)
            want = (torch.arange(n, dtype=torch.float64, device=dev) + 1) * (i + 1) * (world * (world + 1) // 2)
            peer_ok = peer_ok and bool(torch.equal(buf, want))
    clouds = O.synthetic_cloud(2 * world, n_points, 4321)
    lo, hi = dp.shard_range(2 * world, rank, world)

    # norm_act alone, unequal shards (the row count rides behind the sums)
    g = torch.Generator().manual_seed(5)
    x_all = torch.randn(1, 3000, 64, generator=g)
    gy_all = torch.randn(1, 3000, 64, generator=g)
    gam, bet = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1
    cut = 1800
    sl = slice(0, cut) if rank == 0 else slice(cut, 3000)
    x = x_all[:, sl].contiguous().to(dev).requires_grad_(True)
    ga, be = gam.to(dev).requires_grad_(True), bet.to(dev).requires_grad_(True)
    rm, rv = torch.zeros(64, device=dev), torch.ones(64, device=dev)
    y = ops.norm_act(x, ga, be, None, rm, rv, 0.1, 1e-5, 0.01, False, sync_group=True)
    y.backward(gy_all[:, sl].contiguous().to(dev))
    norm_out = (y.detach().cpu(), x.grad.cpu(), ga.grad.cpu(), be.grad.cpu(), rm.cpu(), rv.cpu())

    net = _build(dev, n_points)
    blocks.convert_sync_batchnorm(net, peer_memory=peer)
    assert (ops._PEER_MAILBOX is not None) == bool(peer)
    bucket = dp.FlatGradBucket(net.parameters())
    bucket.zero_()
    out = net(clouds[lo:hi].to(dev))
    loss = out.feats.square().mean()
    loss.backward()
    bucket.all_reduce_mean()
    torch.cuda.synchronize()
    q.put((rank, out.feats.detach().cpu(), float(loss), bucket.flat.cpu(), norm_out, peer_ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 CUDA devices")
@pytest.mark.parametrize("peer", [False, True], ids=["nccl", "peer_memory"])
def test_two_rank_syncbn_matches_single_process(peer):
    import torch.multiprocessing as mp
    from equi_articulated_pose_b200 import ops, dataparallel as dp
    from oracle import so3 as O
    world, n_points, port = 2, 256, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_points, peer, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0

    if peer:
        assert res[0][5] is True and res[1][5] is True
    dev = torch.device("cuda", 0)
    # ---- norm_act: full batch on one device
    g = torch.Generator().manual_seed(5)
    x_all = torch.randn(1, 3000, 64, generator=g).to(dev).requires_grad_(True)
    gy_all = torch.randn(1, 3000, 64, generator=g).to(dev)
    ga = (torch.rand(64, generator=g) + 0.5).to(dev).requires_grad_(True)
    be = (torch.randn(64, generator=g) * 0.1).to(dev).requires_grad_(True)
    rm, rv = torch.zeros(64, device=dev), torch.ones(64, device=dev)
    y = ops.norm_act(x_all, ga, be, None, rm, rv, 0.1, 1e-5, 0.01, False)
    y.backward(gy_all)
    y2 = torch.cat([res[0][4][0], res[1][4][0]], 1)
    gx2 = torch.cat([res[0][4][1], res[1][4][1]], 1)
    assert torch.allclose(y2, y.detach().cpu(), atol=2e-6, rtol=1e-6)
    assert torch.allclose(gx2, x_all.grad.cpu(), atol=2e-6, rtol=1e-5)
    assert torch.allclose(res[0][4][2] + res[1][4][2], ga.grad.cpu(), atol=1e-3, rtol=1e-5)
    assert torch.allclose(res[0][4][3] + res[1][4][3], be.grad.cpu(), atol=1e-3, rtol=1e-5)
    for r in res:   # running statistics: global mean / unbiased global variance on every rank
        assert torch.allclose(r[4][4], rm.cpu(), atol=1e-6) and torch.allclose(r[4][5], rv.cpu(), atol=1e-6)

    # ---- backbone: 2 ranks x 2 clouds == 1 process x 4 clouds
    clouds = O.synthetic_cloud(2 * world, n_points, 4321)
    net = _build(dev, n_points)
    bucket = dp.FlatGradBucket(net.parameters())
    bucket.zero_()
    out = net(clouds.to(dev))
    loss = out.feats.square().mean()
    loss.backward()
    feats = torch.cat([res[0][1], res[1][1]], 0)
    ref = out.feats.detach().cpu()
    assert float((feats - ref).abs().max() / ref.abs().max()) < 1e-4
    assert abs(0.5 * (res[0][2] + res[1][2]) - float(loss)) < 1e-4 * abs(float(loss))
    gref = bucket.flat.cpu()
    for r in res:
        assert float((r[3] - gref).abs().max() / gref.abs().max()) < 3e-2      # fp32 gradient noise floor
    assert torch.equal(res[0][3], res[1][3])
