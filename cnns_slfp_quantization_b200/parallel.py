"""Data-parallel plumbing: one process per GPU, batch sharded over ranks, weights
and static scales replicated.  Inference needs no collective.  The two exchange steps of the path are
  * calibration: one allreduce(MAX) over the per-layer abs-max vector (calibration.ScaleCalibrator.scales);
  * QAT fine-tuning: gradient allreduce(SUM)/world before the revised-SGD step, bucketed so NCCL launches
    stay few; BatchNorm statistics stay per replica (the reference has no SyncBN).
torch.distributed (NCCL on GPUs, gloo in the CPU tests) is plumbing only.
"""
import os

import torch
import torch.distributed as dist


def init(backend=None):
    """Initialise the default process group from the torchrun environment (no-op for a single process)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 or dist.is_initialized():
        return int(os.environ.get("RANK", "0")), world
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group(backend)
    return dist.get_rank(), dist.get_world_size()


def bind_to_local_cpus(device_index):
    """Restrict this process to the CPUs NVML reports as local to its GPU, so that pinned host buffers allocated
    afterwards are first-touched on the GPU's own NUMA node (eight ranks staging 154 MB per step each otherwise cross the
    socket interconnect).  Returns the CPU list it bound to, or None when NVML has no answer or the local set does not
    intersect the CPUs this process may use (cgroup cpuset); never raises."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis and not vis.split(",")[0].strip().isdigit():
            return None
        phys = int(vis.split(",")[device_index]) if vis else device_index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        local = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        allowed = local & os.sched_getaffinity(0)
        if not allowed or allowed == os.sched_getaffinity(0):
            return None
        os.sched_setaffinity(0, allowed)
        return sorted(allowed)
    except Exception:
        return None


def shard_batch(n_items, rank, world):
    """Contiguous shard [lo, hi) of a batch of independent images for this rank."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


def allreduce_gradients(params, bucket_bytes=25 << 20, group=None):
    """Average gradients over ranks in flat buckets (~25 MB: launch-latency sized, not link sized)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    n_calls, bucket, size = 0, [], 0

    def flush():
        nonlocal bucket, size, n_calls
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        bucket, size = [], 0
        n_calls += 1

    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
    return n_calls


def gather_predictions(local_top1, group=None):
    """All ranks' top-1 vectors concatenated in rank order (for whole-batch accuracy bookkeeping)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_top1
    parts = [torch.empty_like(local_top1) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, local_top1, group=group)
    return torch.cat(parts)
