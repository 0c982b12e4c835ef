"""CPU, world_size 2, gloo: the N > 1 host logic of the path - batch sharding with no
data-path collective, the calibration allreduce(MAX) that makes the scales bit-identical on every rank, and the
bucketed gradient allreduce in front of the revised SGD."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from cnns_slfp_quantization_b200 import parallel, calibration
    r, w = parallel.init("gloo")
    assert (r, w) == (rank, world)
    # 1. shards are disjoint, ordered and cover the batch
    lo, hi = parallel.shard_batch(101, rank, world)
    spans = [None] * world
    dist.all_gather_object(spans, (lo, hi))
    assert spans[0][0] == 0 and spans[-1][1] == 101 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    # 2. calibration: per-rank maxima -> one allreduce(MAX) -> identical scales everywhere
    model = nn.Sequential(nn.Linear(4, 4))
    cal = calibration.ScaleCalibrator(model)
    cal.layers = [None, None, None]                                   # three quantized layers' worth of maxima
    rng = np.random.default_rng(rank)
    cal.act_max = torch.from_numpy(rng.uniform(1, 9, 3).astype(np.float32))
    cal.wgt_max = torch.from_numpy(rng.uniform(0.1, 2, 3).astype(np.float32))
    mine = (cal.act_max.clone(), cal.wgt_max.clone())
    ka, kw = cal.scales()
    both = [None] * world
    dist.all_gather_object(both, (mine[0].numpy(), mine[1].numpy()))
    want_a = np.max([b[0] for b in both], 0).astype(np.float64) / 15.5
    want_w = np.max([b[1] for b in both], 0).astype(np.float64) / 15.5
    assert (ka == want_a).all() and (kw == want_w).all()
    # 3. gradient allreduce: mean over ranks, several buckets
    torch.manual_seed(0)
    params = [nn.Parameter(torch.zeros(n)) for n in (1000, 7, 3000, 1)]
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    calls = parallel.allreduce_gradients(params, bucket_bytes=8000)
    assert calls >= 2
    for i, p in enumerate(params):
        assert torch.allclose(p.grad, torch.full_like(p, (i + 1) * (world + 1) / 2.0))
    # 3b. a parameter that got no gradient on this rank contributes zeros (same collectives on every rank)
    q = [nn.Parameter(torch.zeros(5)), nn.Parameter(torch.zeros(3))]
    q[0].grad = torch.full((5,), float(rank + 1))
    if rank == 0:
        q[1].grad = torch.full((3,), 4.0)
    parallel.allreduce_gradients(q)
    assert torch.allclose(q[0].grad, torch.full((5,), 1.5)) and torch.allclose(q[1].grad, torch.full((3,), 2.0))
    # 3c. GradientArena: bucketed allreduce launched from grad hooks DURING backward, gradients as views of one arena
    torch.manual_seed(1)
    net = nn.Sequential(nn.Linear(16, 64), nn.ReLU(), nn.Linear(64, 64), nn.ReLU(), nn.Linear(64, 8))
    ref_net = nn.Sequential(nn.Linear(16, 64), nn.ReLU(), nn.Linear(64, 64), nn.ReLU(), nn.Linear(64, 8))
    ref_net.load_state_dict(net.state_dict())
    arena = parallel.GradientArena(net.parameters(), bucket_bytes=2048)
    assert len(arena.buckets) >= 3
    xs = [torch.randn(4, 16, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
    for step in range(2):
        if step == 0:
            arena.zero_grad()
        else:
            net.zero_grad(set_to_none=True)          # what optimizer.zero_grad() does by default: the hook restores the views
        net(xs[rank]).square().sum().backward()
        n_reduced = arena.finish()
        assert n_reduced == len(arena.buckets)
        ref_net.zero_grad()
        for r in range(world):                        # the same average computed locally
            (ref_net(xs[r]).square().sum() / world).backward()
        for p, q_ in zip(net.parameters(), ref_net.parameters()):
            assert p.grad.data_ptr() == arena.views[p].data_ptr()
            assert torch.allclose(p.grad, q_.grad, rtol=1e-5, atol=1e-6)
    arena.close()
    # 4. predictions gathered in rank order; ragged shards (n_items % world != 0) are padded and trimmed
    top1 = parallel.gather_predictions(torch.tensor([rank * 10, rank * 10 + 1]))
    assert top1.tolist() == [0, 1, 10, 11]
    lo5, hi5 = parallel.shard_batch(5, rank, world)
    rag = parallel.gather_predictions(torch.arange(lo5, hi5), n_items=5)
    assert rag.tolist() == [0, 1, 2, 3, 4]
    lo1, hi1 = parallel.shard_batch(1, rank, world)          # rank 1 holds an empty shard
    assert parallel.gather_predictions(torch.arange(lo1, hi1), n_items=1).tolist() == [0]
    out.put((rank, ka.tolist()))
    dist.destroy_process_group()


def test_two_rank_host_logic_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1]            # bit-identical scales on both ranks
