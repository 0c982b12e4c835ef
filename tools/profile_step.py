#!/usr/bin/env python
"""One eager step of the bench workload between cudaProfilerStart/Stop, for ncu launch lists:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py [--batch 256] [--size 224]

Builds the same plan bench.py times (SLFP-8 ResNet-50), warms it up, then runs exactly one step
(weight re-quantization + the fused forward) inside the profiler range.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
from cnns_slfp_quantization_b200 import engine, nets_common as nc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--steps", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    model = bench.build_model_gpu(args.size, dev)
    plan = engine.compile_resnet50(model, args.batch, args.size, device=dev)
    plan.input.copy_(nc.synth_images(args.batch, args.size, seed=1234).to(dev))
    plan.prepare_weights()
    for _ in range(2):
        plan.run()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for _ in range(args.steps):
        plan.run()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("profiled", plan.launches_per_step * args.steps, "launches")


if __name__ == "__main__":
    main()
