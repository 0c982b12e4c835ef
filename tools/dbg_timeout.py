#!/usr/bin/env python
"""Run one microbench conv layer with the barrier-timeout recorder installed and print the record on failure."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cnns_slfp_quantization_b200 import _native as nv
import tools.microbench as mb
lib = nv.lib()
buf = torch.zeros(4, dtype=torch.int64).pin_memory()
cudart = torch.cuda.cudart()
torch.zeros(1, device="cuda")
dptr = ctypes.c_void_p()
rt = ctypes.CDLL("libcudart.so.12")
assert rt.cudaHostGetDevicePointer(ctypes.byref(dptr), ctypes.c_void_p(buf.data_ptr()), 0) == 0
nv.check(lib.slfp_debug_set_buffer(dptr))
try:
    mb.bench_conv(mode=sys.argv[1], only=sys.argv[2], iters=1)
    print("ok")
except Exception as e:
    print("FAILED", str(e)[:100])
v = int(buf[1]) & 0xffffffffffffffff
tag = v >> 32
print("timeouts", int(buf[0]), "tag", tag & 0xff, "kb/it", (tag >> 8) & 0xff, "ti", tag >> 16, "block", (v >> 8) & 0xffffff, "warp", v & 0xff)
