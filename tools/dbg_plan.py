#!/usr/bin/env python
"""Debug aid: run the bench plan op by op (synchronising after each) for several steps with the barrier-timeout
recorder installed; prints the failing op and the recorded (tag, K block, tile, block, warp)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from cnns_slfp_quantization_b200 import engine, nets_common as nc, _native as nv
dev = torch.device('cuda', 0)
lib = nv.lib()
buf = torch.zeros(4, dtype=torch.int64).pin_memory()
torch.zeros(1, device="cuda")
dptr = ctypes.c_void_p()
rt = ctypes.CDLL("libcudart.so.12")
assert rt.cudaHostGetDevicePointer(ctypes.byref(dptr), ctypes.c_void_p(buf.data_ptr()), 0) == 0
nv.check(lib.slfp_debug_set_buffer(dptr))
model = bench.build_model_gpu(224, dev)
plan = engine.compile_resnet50(model, 256, 224, device=dev)
plan.input.copy_(nc.synth_images(256, 224, seed=1234).to(dev))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
sync_each = len(sys.argv) <= 2
st = nv.stream()
if len(sys.argv) > 2 and sys.argv[2] == "graph":
    try:
        plan.prepare_weights(); plan.run(); torch.cuda.synchronize()
        plan.capture()
        for s in range(steps):
            plan()
            torch.cuda.synchronize()
        print("graph ok")
    except Exception as e:
        print("FAILED at step", s, str(e)[:100])
    v = int(buf[1]) & 0xffffffffffffffff
    tag = v >> 32
    print("timeouts", int(buf[0]), "tag", tag & 0xff, "kb", (tag >> 8) & 0xff, "ti", tag >> 16, "block", (v >> 8) & 0xffffff, "warp", v & 0xff)
    sys.exit(0)
try:
    for s in range(steps):
        plan.prepare_weights()
        for i, op in enumerate(plan.ops):
            op(st)
            if sync_each:
                try:
                    torch.cuda.synchronize()
                except Exception as e:
                    print('step', s, 'op', i, 'of', len(plan.ops), 'failed', str(e)[:80])
                    raise
        torch.cuda.synchronize()
    print('all ok')
except Exception as e:
    print("FAILED", str(e)[:100])
v = int(buf[1]) & 0xffffffffffffffff
tag = v >> 32
print("timeouts", int(buf[0]), "tag", tag & 0xff, "kb", (tag >> 8) & 0xff, "ti", tag >> 16, "block", (v >> 8) & 0xffffff, "warp", v & 0xff)
print([c[2] for c in plan.conv_flops][:8])
