"""GPU: the thin driver with the reference's command line (tools/train_eval.py) - same flags,
same `samples/sec` log line, a state_dict written by it loads back through --pretrain, calibration files are written."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, cwd):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "train_eval.py")] + args, cwd=cwd, capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p.stdout


def test_reference_flags_train_eval_calibrate(tmp_path):
    cwd = str(tmp_path)
    out = _run(["--net", "vgg16", "--Qbits", "8", "--retrain", "--optimizer", "DSGD", "--max_epochs", "1", "--steps_per_epoch", "3",
                "--train_batch_size", "16", "--eval_batch_size", "16", "--log_interval", "1", "--save_model", "--root_dir", cwd], cwd)
    assert "optimizer => DSGD" in out
    assert len(re.findall(r"cls_loss= [0-9.]+ \(\d+ samples/sec\)", out)) == 3          # cifar100_train_eval.py:185-187
    assert "Precision@1:" in out and "fused engine" in out
    ckpt = os.path.join(cwd, "ckpt", "vgg160_tmp.pth")
    assert os.path.exists(ckpt)
    out = _run(["--net", "vgg16", "--Qbits", "8", "--pretrain", ckpt, "--pre_reference", "--steps_per_epoch", "2", "--eval_batch_size", "32"], cwd)
    assert os.path.exists(os.path.join(cwd, "max_inout_vgg16.txt")) and os.path.exists(os.path.join(cwd, "max_weight_vgg16.txt"))
    assert "Layer 0 Max Absolute Input:" in open(os.path.join(cwd, "max_inout_vgg16.txt")).read()
    out = _run(["--net", "mobilenet", "--Qbits", "7", "--steps_per_epoch", "2", "--eval_batch_size", "32"], cwd)
    assert "Precision@1:" in out
    # the reference's undefined optimizer name fails the same way (NameError, cifar100_train_eval.py:143-145)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "train_eval.py"), "--net", "vgg16", "--optimizer", "CustomSGD"], cwd=cwd,
                       capture_output=True, text=True, timeout=300)
    assert p.returncode != 0 and "NameError" in p.stderr
