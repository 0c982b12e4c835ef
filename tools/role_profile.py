#!/usr/bin/env python
"""Where each role of the dense conv kernel spends its time (library built with SLFP_EXTRA_NVCC_FLAGS=-DSLFP_ROLE_PROFILE).

    python tools/role_profile.py res l3.c2     # microbench mode, layer filter

Prints, per role lead thread and averaged over CTAs, the share of its life spent in each pipeline wait.
"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cnns_slfp_quantization_b200 import _native as nv
import tools.microbench as mb
lib = nv.lib()
buf = torch.zeros(32, dtype=torch.int64).pin_memory()
torch.zeros(1, device="cuda")
dptr = ctypes.c_void_p()
rt = ctypes.CDLL("libcudart.so.12")
assert rt.cudaHostGetDevicePointer(ctypes.byref(dptr), ctypes.c_void_p(buf.data_ptr()), 0) == 0
nv.check(lib.slfp_debug_set_buffer(dptr))
mb.bench_conv(mode=sys.argv[1], only=sys.argv[2], iters=1)
torch.cuda.synchronize()
names = {8: ("code producer", "wait cempty", "expect_tx", "tma issue"), 12: ("weight producer", "wait empty", "-", "-"),
         16: ("mma issuer", "wait tempty", "wait full", "issue+commit"), 20: ("decode warp 0", "wait cfull", "wait empty", "work"),
         24: ("epilogue warp 0", "wait tfull", "leader: store drain (staged)", "barriers (staged)")}
for slot, (role, a, b, c) in names.items():
    va, vb, vc, tot = [int(buf[slot + i]) for i in range(4)]
    if tot == 0:
        continue
    print(f"{role:16s} total {tot/1e6:9.2f} Mcyc | {a}: {va/tot:6.1%} | {b}: {vb/tot:6.1%} | {c}: {vc/tot:6.1%}")
