#!/usr/bin/env python
"""Timing of one QAT fine-tune step (BASELINE.json configs[3]): ResNet-50 SLFP-8 through the drop-in modules
(conv2d_Q forward + straight-through backward) + the revised SGD (DSGD), synthetic data.
    python tools/bench_qat.py [batch per GPU] [size] [steps]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_qat.py ...
Data-parallel over ranks: gradients are averaged with bucketed NCCL allreduce(SUM) before the revised-SGD step
(cnns_slfp_quantization_b200/parallel.py); time is the max over ranks.  Prints one JSON line on rank 0; SLFP_QAT_PROFILE=1 adds the per-entry-point CUDA-event times of one step."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cnns_slfp_quantization_b200 import nets_common as nc, calibration, parallel, _native as nv
from cnns_slfp_quantization_b200.nets_imgnet import ResNet50
from cnns_slfp_quantization_b200.utils.optimizer import DSGD

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
size = int(sys.argv[2]) if len(sys.argv) > 2 else 224
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rank, world = parallel.init()
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
m32 = ResNet50(32, scales=(np.ones(54), np.ones(54))).eval()
sd = nc.synth_state_dict(m32)
m32.load_state_dict(sd)
m32 = m32.to(dev)
ka, kw = calibration.calibrate_scales(m32, [nc.synth_images(4, min(size, 96)).to(dev)])
m = ResNet50(8, scales=(np.ones(54), np.ones(54)))
m.load_state_dict(sd)
nc.set_scales(m, ka, kw)                 # calibration order = module traversal order
m = m.to(dev).train()
opt = DSGD(m.parameters(), 8, lr=1e-3, momentum=0.9, weight_decay=5e-4)
x = nc.synth_images(batch, size, seed=1234 + rank).to(dev)
y = torch.randint(0, 1000, (batch,), device=dev)
crit = torch.nn.CrossEntropyLoss()


def step():
    opt.zero_grad()
    loss = crit(m(x), y)
    loss.backward()
    parallel.allreduce_gradients(m.parameters())
    opt.step()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    loss = step()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
if world > 1:
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t.item())
out = {"net": "resnet50_slfp8_qat", "n_gpus": world, "batch_per_gpu": batch, "size": size, "ms_per_step": round(ms, 3),
       "images_per_s": round(world * batch / ms * 1e3, 1), "loss": float(loss.detach())}
if os.environ.get("SLFP_QAT_PROFILE"):
    nv.profile = {}
    torch.cuda.profiler.start()          # ncu --profile-from-start off captures exactly this step
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    prof, nv.profile = nv.profile, None
    out["per_entry_point_ms"] = {k: round(sum(p.elapsed_time(q) for p, q, _ in v), 3) for k, v in prof.items()}
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    torch.distributed.destroy_process_group()
