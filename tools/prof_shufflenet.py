import json, sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from cnns_slfp_quantization_b200 import nets_common as nc, calibration, _native as nv
from cnns_slfp_quantization_b200.nets_cifar import ShuffleNetV2
dev = torch.device("cuda", 0)
size, batch = 224, 64
m32 = ShuffleNetV2(32).eval(); sd = nc.synth_state_dict(m32); m32.load_state_dict(sd); nc.set_scales(m32, np.ones(57), np.ones(57)); m32 = m32.to(dev)
ka, kw = calibration.calibrate_scales(m32, [nc.synth_images(8, size).to(dev)])
m = ShuffleNetV2(7).eval(); m.load_state_dict(sd); nc.set_scales(m, ka, kw); m = m.to(dev)
x = nc.synth_images(batch, size).to(dev).contiguous(memory_format=torch.channels_last)
with torch.no_grad():
    for _ in range(2): m(x)
    torch.cuda.synchronize()
    nv.profile = {}
    m(x); torch.cuda.synchronize()
    prof, nv.profile = nv.profile, None
print({k: (len(v), round(sum(a.elapsed_time(b) for a, b, _ in v), 3)) for k, v in prof.items()})
conv = prof.get("slfp_conv2d_fwd", [])
ts = sorted(((a.elapsed_time(b), i) for i, (a, b, _) in enumerate(conv)), reverse=True)[:8]
layers = nc.quantized_layers(m)
for t, i in ts:
    l = layers[i]
    print(round(t, 3), i, type(l).__name__, getattr(l, 'in_channels', None), getattr(l, 'out_channels', None), getattr(l, 'kernel_size', None), getattr(l, 'stride', None), getattr(l, 'groups', None))
