import sys, os
sys.path.insert(0, '/root/repo')
import torch, bench
from cnns_slfp_quantization_b200 import engine, nets_common as nc, _native as nv
dev=torch.device('cuda',0)
model=bench.build_model_gpu(224, dev)
plan=engine.compile_resnet50(model, 256, 224, device=dev)
plan.input.copy_(nc.synth_images(256,224,seed=1234).to(dev))
plan.prepare_weights(); torch.cuda.synchronize()
st=nv.stream()
for i,op in enumerate(plan.ops):
    op(st)
    try:
        torch.cuda.synchronize()
    except Exception as e:
        print('op',i,'failed', plan.conv_flops[max(0,i-3):i+1]); raise
print('all ok')
