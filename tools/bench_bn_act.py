"""Fused training BatchNorm + ReLU (utils/bn_act.bn_act, csrc/bn_act.cu) against the stock nn.BatchNorm2d + relu on a few
tensor shapes: precision of the batch statistics against float64, and forward + backward time of both.
    python tools/bench_bn_act.py
ncu of the four kernels on the [128, 256, 56, 56] tensor:
    ncu --set full --clock-control none -k "regex:bn_reduce|bn_apply|bn_bwd_apply" --launch-skip 170 --launch-count 4 -o out python tools/bench_bn_act.py
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn as nn
from cnns_slfp_quantization_b200.utils.bn_act import bn_act
dev = torch.device("cuda:0")
torch.manual_seed(2048)
for shape in [(2, 2048, 3, 3), (2, 1024, 5, 5), (128, 64, 56, 56), (128, 256, 56, 56)]:
    x = (torch.randn(shape, device=dev) * 1.7 + 0.3).contiguous(memory_format=torch.channels_last)
    bn = nn.BatchNorm2d(shape[1]).to(dev).train()
    bn.running_mean.zero_(); bn.momentum = 1.0
    y = bn_act(x, bn, relu=False)
    m64 = x.double().mean(dim=(0, 2, 3))
    v64 = x.double().var(dim=(0, 2, 3), unbiased=True)
    print(shape, "mean err", float((bn.running_mean.double() - m64).abs().max()), "var err", float((bn.running_var.double() - v64).abs().max()),
          "stock:", end=" ")
    bn2 = nn.BatchNorm2d(shape[1]).to(dev).train(); bn2.momentum = 1.0
    bn2(x)
    print(float((bn2.running_mean.double() - m64).abs().max()), float((bn2.running_var.double() - v64).abs().max()))
    # timing
    gy = torch.randn_like(x)
    xr = x.clone().requires_grad_(True)
    for name, fn in (("fused", lambda: bn_act(xr, bn, relu=True)), ("stock", lambda: torch.relu(bn2(xr)))):
        for _ in range(3):
            xr.grad = None; fn().backward(gy)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            xr.grad = None; fn().backward(gy)
        b.record(); torch.cuda.synchronize()
        print("   ", name, "fwd+bwd ms", a.elapsed_time(b) / 10, "GB moved(min)", x.numel() * 4 * 10 / 1e9)
