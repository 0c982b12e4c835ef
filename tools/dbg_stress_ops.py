#!/usr/bin/env python
"""Debug aid: find a kernel (or a pair of consecutive kernels) of the bench plan that fails intermittently.
    python tools/dbg_stress_ops.py REPS WINDOW [first_op [last_op]]
runs ops[i : i + WINDOW] REPS times back to back for every i, stops at the first failure."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from cnns_slfp_quantization_b200 import engine, nets_common as nc, _native as nv
dev = torch.device('cuda', 0)
lib = nv.lib()
model = bench.build_model_gpu(224, dev)
plan = engine.compile_resnet50(model, 256, 224, device=dev)
plan.input.copy_(nc.synth_images(256, 224, seed=1234).to(dev))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
win = int(sys.argv[2]) if len(sys.argv) > 2 else 1
first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
last = int(sys.argv[4]) if len(sys.argv) > 4 else len(plan.ops) - 1
st = nv.stream()
plan.prepare_weights()
for op in plan.ops:
    op(st)
torch.cuda.synchronize()
t0 = time.time()
for i in range(first, min(last, len(plan.ops) - win) + 1):
    try:
        for _ in range(reps):
            for op in plan.ops[i:i + win]:
                op(st)
        torch.cuda.synchronize()
    except Exception as e:
        print("window", i, "..", i + win - 1, "of", len(plan.ops), "FAILED:", str(e)[:60], flush=True)
        sys.exit(1)
    if time.time() - t0 > 200:
        print("time limit at window", i); break
print("all windows ok: reps", reps, "window", win, flush=True)
if os.environ.get("DBG_WPREP"):
    try:
        for s in range(int(os.environ["DBG_WPREP"])):
            plan.prepare_weights()
            for op in plan.ops[:int(os.environ.get("DBG_NOPS", "1000"))]:
                op(st)
        torch.cuda.synchronize()
        print("wprep + ops ok", flush=True)
    except Exception as e:
        print("wprep + ops FAILED at step", s, str(e)[:300], flush=True); sys.exit(1)
