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
    """Average gradients over ranks in flat buckets (~25 MB: launch-latency sized, not link sized), AFTER backward has
    returned.  Simple form kept for tests and as the fallback of GradientArena; a parameter without a gradient on this
    rank contributes zeros, so every rank issues the same collectives whatever its local graph touched."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    params = [p for p in params if p.requires_grad]
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    grads = [p.grad for p in params]
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


class GradientArena:
    """Bucketed gradient allreduce OVERLAPPED with backward (the insertion point is between
    `loss.backward()` and `optimizer.step()`, cifar100_train_eval.py:177-179).

    All gradients live in ONE pre-flattened arena (p.grad is a view into it: no torch.cat, no copy-back).  Parameters
    are laid out in REVERSE registration order - roughly the order backward produces them - and cut into ~25 MB
    buckets; a post-accumulate-grad hook per parameter counts its bucket down and, when the last gradient of a bucket
    has arrived, launches that bucket's allreduce(SUM) asynchronously (NCCL runs it on its own stream while autograd
    keeps computing the earlier layers' dgrad / wgrad).  `finish()` waits for the handles and applies 1 / world.

        arena = GradientArena(model.parameters())
        for batch in loader:
            arena.zero_grad()                 # instead of optimizer.zero_grad(): keeps the views
            loss = criterion(model(x), y); loss.backward()
            arena.finish()                    # every bucket reduced and averaged
            optimizer.step()
    Single process (world 1): the hooks only keep the views in place; finish() is a no-op."""

    def __init__(self, params, bucket_bytes=25 << 20, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        order = list(reversed(self.params))
        assert order, "no parameters"
        dev, dt = order[0].device, order[0].dtype
        assert all(p.device == dev and p.dtype == dt for p in order), "one arena = one device and dtype"
        total = sum(p.numel() for p in order)
        self.flat = torch.zeros(total, dtype=dt, device=dev)
        self.views, self.bucket_of, self.buckets = {}, {}, []          # buckets: [lo, hi, n_params]
        off = lo = 0
        count = 0
        for p in order:
            v = self.flat[off:off + p.numel()].view_as(p)
            self.views[p] = v
            self.bucket_of[p] = len(self.buckets)
            p.grad = v
            off += p.numel()
            count += 1
            if (off - lo) * self.flat.element_size() >= bucket_bytes:
                self.buckets.append([lo, off, count])
                lo, count = off, 0
        if count:
            self.buckets.append([lo, off, count])
        self._pending = [b[2] for b in self.buckets]
        self._handles = []
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    def _on_grad(self, p):
        v = self.views[p]
        if p.grad is not v:                       # someone reset p.grad (zero_grad(set_to_none=True)): one copy, view restored
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
            p.grad = v
        b = self.bucket_of[p]
        self._pending[b] -= 1
        if self._pending[b] == 0 and self.world > 1:
            lo, hi, _ = self.buckets[b]
            self._handles.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def zero_grad(self):
        self.flat.zero_()
        for p in self.params:
            p.grad = self.views[p]

    def finish(self):
        """Wait for the in-flight buckets, reduce any bucket whose parameters got no gradient this step (zeros: keeps
        the collective sequence identical on every rank) and average."""
        if self.world > 1:
            for b, left in enumerate(self._pending):
                if left > 0:
                    lo, hi, _ = self.buckets[b]
                    self._handles.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            for h in self._handles:
                h.wait()
            self.flat.div_(self.world)
        n = len(self._handles)
        self._handles = []
        self._pending = [b[2] for b in self.buckets]
        return n

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def gather_predictions(local_top1, group=None, n_items=None):
    """All ranks' top-1 vectors concatenated in rank order (for whole-batch accuracy bookkeeping).  shard_batch gives the
    last ranks a shorter (possibly empty) shard when n_items % world != 0; all_gather needs equal sizes, so every rank
    pads to ceil(n_items / world) with -1 and the padding is trimmed after the gather (pass the global n_items; without it
    the shards must be equal)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_top1
    world = dist.get_world_size(group)
    if n_items is None:
        parts = [torch.empty_like(local_top1) for _ in range(world)]
        dist.all_gather(parts, local_top1, group=group)
        return torch.cat(parts)
    per = (n_items + world - 1) // world
    padded = torch.full((per,), -1, dtype=local_top1.dtype, device=local_top1.device)
    padded[:local_top1.numel()] = local_top1
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    out = []
    for r, part in enumerate(parts):
        lo, hi = shard_batch(n_items, r, world)
        out.append(part[:hi - lo])
    return torch.cat(out)
