#!/usr/bin/env python
"""Thin bench / compatibility driver with the REFERENCE's command line.

Mirrors the flags of cifar100_train_eval.py:47-71 / imgnet_train_eval.py:32-57 - `--Qbits {7,8,32}`, `--net`,
`--optimizer {SGD,DSGD,SSGD,NormalSGD,Adam,RMSprop}`, `--lr`, `--wd`, `--train_batch_size`, `--eval_batch_size`,
`--retrain`, `--pretrain`, `--pre_reference`, `--save_model`, `--max_epochs`, `--log_interval`, `--use_gpu`, `--cluster` -
and its log line `... cls_loss= %.5f (%d samples/sec)` (cifar100_train_eval.py:185-187), so a user of the reference can
run the same invocations against this framework.  What it is NOT: the reference's data pipeline.  There is no network
and no dataset in this environment, so batches are synthetic (`--steps_per_epoch` of them, nets_common.synth_images);
with `--pretrain PATH` a state_dict file of the reference loads unchanged (strict=False like :158-159).

    python tools/train_eval.py --net resnet50 --Qbits 8                      # eval loop (fused engine), samples/sec
    python tools/train_eval.py --net vgg16 --Qbits 8 --retrain --optimizer DSGD --max_epochs 1
    python tools/train_eval.py --net mobilenet --pre_reference               # calibration: max_inout_<net>.txt, max_weight_<net>.txt
    torchrun --nproc-per-node N tools/train_eval.py ...                      # data parallel (sharded batch, allreduce)
"""
import argparse
import os
import sys
import time
from datetime import datetime

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NETS = {  # name -> (constructor, input size, classes, fused-plan compiler name or None)
    "mobilenet": (lambda q: __import__("cnns_slfp_quantization_b200.nets_cifar", fromlist=["x"]).MobileNetV1_Q(3, q), 32, 100, "compile_mobilenetv1"),
    "vgg16": (lambda q: __import__("cnns_slfp_quantization_b200.nets_cifar", fromlist=["x"]).VGG16_Q(q), 32, 100, "compile_vgg16"),
    "shufflenetv2": (lambda q: __import__("cnns_slfp_quantization_b200.nets_cifar", fromlist=["x"]).ShuffleNetV2(q), 32, 100, "compile_shufflenetv2"),
    "resnet50": (lambda q: __import__("cnns_slfp_quantization_b200.nets_imgnet", fromlist=["x"]).ResNet50(q), 224, 1000, "compile_resnet50"),
    "mobilenet_imgnet": (lambda q: __import__("cnns_slfp_quantization_b200.nets_imgnet", fromlist=["x"]).MobileNetV1_Q(3, q), 224, 1000, "compile_mobilenetv1"),
    "alexnet": (lambda q: __import__("cnns_slfp_quantization_b200.nets_imgnet", fromlist=["x"]).AlexNet(q), 224, 1000, None),
    "squeezenet": (lambda q: __import__("cnns_slfp_quantization_b200.nets_imgnet", fromlist=["x"]).SqueezeNet(q), 224, 1000, None),
}


def main():
    ap = argparse.ArgumentParser(description="SLFP reference and retrain - B200-native hot path, the reference's flags")
    ap.add_argument("--root_dir", type=str, default="./")
    ap.add_argument("--log_name", type=str, default="synthetic")
    ap.add_argument("--retrain", action="store_true", default=False)
    ap.add_argument("--save_model", action="store_true", default=False)
    ap.add_argument("--pre_reference", action="store_true", default=False)
    ap.add_argument("--pretrain", type=str, nargs="?", const="", default=None,
                    help="state_dict file of the reference (its flag is a switch with a hard-coded path; here the path is the value)")
    ap.add_argument("--optimizer", type=str, default="SGD")
    ap.add_argument("--net", type=str, default="mobilenet", choices=sorted(NETS))
    ap.add_argument("--Qbits", type=int, default=32)
    ap.add_argument("--lr", type=float, default=0.0001)
    ap.add_argument("--wd", type=float, default=5e-4)
    ap.add_argument("--num", type=int, default=0)
    ap.add_argument("--train_batch_size", type=int, default=256)
    ap.add_argument("--eval_batch_size", type=int, default=128)
    ap.add_argument("--max_epochs", type=int, default=1)
    ap.add_argument("--log_interval", type=int, default=10)
    ap.add_argument("--use_gpu", type=str, default="0")
    ap.add_argument("--cluster", action="store_true", default=False)
    ap.add_argument("--steps_per_epoch", type=int, default=20, help="synthetic batches per epoch (no dataset in this environment)")
    ap.add_argument("--no_engine", action="store_true", help="evaluate through the module-level drop-in instead of the fused engine")
    cfg = ap.parse_args()
    if not cfg.cluster and "LOCAL_RANK" not in os.environ:
        os.environ.setdefault("CUDA_VISIBLE_DEVICES", cfg.use_gpu)            # cifar100_train_eval.py:81-82

    from cnns_slfp_quantization_b200 import calibration, engine, nets_common as nc, parallel
    from cnns_slfp_quantization_b200.utils.optimizer import DSGD, SSGD, NormalSGD
    rank, world = parallel.init()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    ctor, size, classes, comp_name = NETS[cfg.net]
    print("=> creating model", cfg.net, "...") if rank == 0 else None
    model = ctor(cfg.Qbits)
    if cfg.pretrain:
        model.load_state_dict(torch.load(cfg.pretrain, map_location="cpu"), False)
    else:
        model.load_state_dict(nc.synth_state_dict(model))
    model = model.to(dev)
    layers = nc.quantized_layers(model)

    if cfg.pre_reference:
        # calibration (cifar100_train_eval.py:213-301): fused abs-max kernel + ONE allreduce(MAX); same output files
        m32 = ctor(32)
        m32.load_state_dict(model.state_dict(), False)
        nc.set_scales(m32, np.ones(len(layers)), np.ones(len(layers)))
        m32 = m32.to(dev).eval()
        lo, hi = parallel.shard_batch(cfg.eval_batch_size, rank, world)
        batches = [nc.synth_images(cfg.eval_batch_size, size, seed=1234 + i)[lo:hi].to(dev) for i in range(max(1, 1000 // cfg.eval_batch_size))]
        ka, kw = calibration.calibrate_scales(m32, batches, divisor=1.0)
        if rank == 0:
            with open(f"max_inout_{cfg.net}.txt", "w") as f:
                for i, v in enumerate(ka):
                    f.write(f"Layer {i} Max Absolute Input:\n{v}\n\n")
            with open(f"max_weight_{cfg.net}.txt", "w") as f:
                for i, v in enumerate(kw):
                    f.write(f"Layer {i} Max Absolute weight:\n{v}\n\n")
            print(f"Results saved to max_inout_{cfg.net}.txt / max_weight_{cfg.net}.txt")
        nc.set_scales(model, ka / 15.5, kw / 15.5)

    opts = {"SSGD": lambda: SSGD(model.parameters(), qbit=cfg.Qbits, lr=cfg.lr, momentum=0.9, weight_decay=cfg.wd),
            "DSGD": lambda: DSGD(model.parameters(), qbit=cfg.Qbits, lr=cfg.lr, momentum=0.9, weight_decay=cfg.wd),
            "NormalSGD": lambda: NormalSGD(model.parameters(), qbit=cfg.Qbits, lr=cfg.lr, momentum=0.9, weight_decay=cfg.wd),
            "Adam": lambda: torch.optim.Adam(model.parameters(), cfg.lr),
            "RMSprop": lambda: torch.optim.RMSprop(model.parameters(), cfg.lr),
            "SGD": lambda: torch.optim.SGD(model.parameters(), cfg.lr, momentum=0.9, weight_decay=cfg.wd)}
    if cfg.optimizer not in opts:
        raise NameError(f"name '{cfg.optimizer}' is not defined")              # the reference's CustomSGD: NameError (:143-145)
    print("optimizer =>", cfg.optimizer) if rank == 0 else None
    optimizer = opts[cfg.optimizer]()
    sched = torch.optim.lr_scheduler.MultiStepLR(optimizer, [75, 85, 100], gamma=0.1)
    criterion = torch.nn.CrossEntropyLoss().to(dev)
    arena = parallel.GradientArena(model.parameters()) if (cfg.retrain and world > 1) else None
    g = torch.Generator().manual_seed(rank)

    def train(epoch):
        model.train()
        start = time.time()
        lo, hi = parallel.shard_batch(cfg.train_batch_size, rank, world)
        for batch_idx in range(cfg.steps_per_epoch):
            inputs = nc.synth_images(cfg.train_batch_size, size, seed=epoch * 100003 + batch_idx)[lo:hi].to(dev)
            targets = torch.randint(0, classes, (cfg.train_batch_size,), generator=g)[lo:hi].to(dev)
            outputs = model(inputs)
            loss = criterion(outputs, targets)
            arena.zero_grad() if arena is not None else optimizer.zero_grad()
            loss.backward()
            if arena is not None:
                arena.finish()
            optimizer.step()
            if batch_idx % cfg.log_interval == 0 and rank == 0:
                torch.cuda.synchronize()
                duration = time.time() - start
                print("%s epoch: %d step: %d cls_loss= %.5f (%d samples/sec)" %
                      (datetime.now(), epoch, batch_idx, loss.item(), cfg.train_batch_size * min(cfg.log_interval, batch_idx + 1) / max(duration, 1e-9)))
                start = time.time()

    def test(epoch):
        model.eval()
        lo, hi = parallel.shard_batch(cfg.eval_batch_size, rank, world)
        per = hi - lo
        plan = None
        if comp_name and cfg.Qbits in (7, 8) and not cfg.no_engine and per > 0:
            plan = getattr(engine, comp_name)(model, per, size, device=dev)
            plan.capture()
        correct, t0 = 0, time.time()
        for batch_idx in range(cfg.steps_per_epoch):
            inputs = nc.synth_images(cfg.eval_batch_size, size, seed=7_000_003 + batch_idx)[lo:hi].to(dev)
            targets = torch.randint(0, classes, (cfg.eval_batch_size,), generator=torch.Generator().manual_seed(batch_idx))[lo:hi].to(dev)
            with torch.no_grad():
                outputs = plan(inputs) if plan is not None else model(inputs)
            pred = parallel.gather_predictions(outputs.argmax(1), n_items=cfg.eval_batch_size)
            tall = parallel.gather_predictions(targets, n_items=cfg.eval_batch_size)
            correct += int(pred.eq(tall).sum().item())
        torch.cuda.synchronize()
        dt = time.time() - t0
        acc = 100.0 * correct / (cfg.steps_per_epoch * cfg.eval_batch_size)
        if rank == 0:
            print("%s------------------------------------------------------ Precision@1: %.2f%% (%d samples/sec, %s)\n" %
                  (datetime.now(), acc, cfg.steps_per_epoch * cfg.eval_batch_size / dt, "fused engine" if plan is not None else "modules"))
        return acc

    acc_data, acc_max = [], -1.0          # (the reference starts at 0; synthetic labels can score exactly 0 %)
    for epoch in range(cfg.max_epochs):
        sched.step()
        if cfg.retrain:
            train(epoch)
        acc_data.append(test(epoch))
        if cfg.save_model and max(acc_data) > acc_max and rank == 0:
            acc_max = max(acc_data)
            os.makedirs(os.path.join(cfg.root_dir, "ckpt"), exist_ok=True)
            torch.save(model.state_dict(), os.path.join(cfg.root_dir, "ckpt", f"{cfg.net}{cfg.num}_tmp.pth"))
            print("max acc :", acc_max, "\nsaving model....")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
