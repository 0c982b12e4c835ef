"""Builds libslfp_b200.so (hand-written sm_100a CUDA kernels + the C ABI of include/slfp_b200.h).

In-tree build with nvcc; the .so is git-ignored but travels to the GPU box with the snapshot.
    python -m cnns_slfp_quantization_b200.build [--force]
"""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libslfp_b200.so")

# -fmad=false: no implicit multiply-add contraction.  The reference's float32 expressions round every
# product and sum separately (except ATen's add(alpha), which the kernels spell as explicit fmaf), and
# parity with it is bit-exact for the quantizers and the optimizer step.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "-shared", "-Wno-deprecated-gpu-targets"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def source_hash():
    """sha256 over every file the library is compiled from (+ the compiler flags).  It is compiled INTO the library
    (slfp_build_id()) and checked when the library is loaded (_native.lib()): a stale or foreign .so is refused, so a
    test / bench run proves the loaded binary was built from the sources next to it."""
    h = hashlib.sha256()
    deps = sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(HERE, "..", "include", "slfp_b200.h")]
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + os.environ.get("SLFP_EXTRA_NVCC_FLAGS", "").split()).encode())
    return h.hexdigest()[:16]


def built_hash():
    try:
        with open(LIB_PATH + ".hash") as f:
            return f.read().strip()
    except OSError:
        return None


def _stale():
    return not os.path.exists(LIB_PATH) or built_hash() != source_hash()


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    digest = source_hash()
    objs = []
    procs = []
    # one nvcc per translation unit, in parallel
    for src in sources():
        obj = os.path.join(LIB_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + os.environ.get("SLFP_EXTRA_NVCC_FLAGS", "").split() + \
            [f'-DSLFP_SOURCE_HASH="{digest}"', "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libslfp_b200.so")
    subprocess.check_call([nvcc, "-shared", "-o", LIB_PATH] + objs + ["-Wno-deprecated-gpu-targets"])
    for o in objs:
        os.remove(o)
    with open(LIB_PATH + ".hash", "w") as f:
        f.write(digest + "\n")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
