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
    # 4. predictions gathered in rank order
    top1 = parallel.gather_predictions(torch.tensor([rank * 10, rank * 10 + 1]))
    assert top1.tolist() == [0, 1, 10, 11]
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
