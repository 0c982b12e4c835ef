#!/usr/bin/env python
"""Pinned host -> device bandwidth per rank and in aggregate (is the 8-GPU `e2e` figure of bench.py
the box's host-memory / PCIe ceiling?).  Every rank copies the bench's own per-step input (154 MB of float32, pinned)
to its GPU `--reps` times with cudaMemcpyAsync, all ranks at the same time (barrier before, max over ranks after).
    python tools/h2d_probe.py                      # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py
Prints one JSON line on rank 0: per-rank GB/s (min / max), aggregate GB/s, and the images/s ceiling it implies for
bench.py's e2e leg (602 112 input bytes per 224x224 image)."""
import argparse, json, os, sys, time
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=147)            # 256 x 3 x 224 x 224 x 4 bytes = 154 140 672 B = 147 MiB
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--streams", type=int, default=1, help="split every copy over this many streams")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
n = args.mb << 20
host = torch.empty(n, dtype=torch.uint8).pin_memory()
host.random_(0, 255)
devbuf = torch.empty(n, dtype=torch.uint8, device=dev)
streams = [torch.cuda.Stream(device=dev) for _ in range(args.streams)]
chunk = (n + args.streams - 1) // args.streams


def copy_once():
    for i, s in enumerate(streams):
        with torch.cuda.stream(s):
            devbuf[i * chunk:(i + 1) * chunk].copy_(host[i * chunk:(i + 1) * chunk], non_blocking=True)


for _ in range(3):
    copy_once()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(args.reps):
    copy_once()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
gbs = n * args.reps / dt / 1e9
t = torch.tensor([gbs, -gbs, dt], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    slowest = n * args.reps / float(t[2]) / 1e9
    print(json.dumps({"probe": "pinned_h2d", "n_gpus": world, "mb_per_copy": args.mb, "reps": args.reps, "streams": args.streams,
                      "per_rank_gbs_max": round(float(t[0]), 2), "per_rank_gbs_min": round(-float(t[1]), 2),
                      "aggregate_gbs": round(world * slowest, 2),
                      "e2e_images_per_s_ceiling": round(world * slowest * 1e9 / 602112)}), flush=True)
if world > 1:
    dist.destroy_process_group()
