"""Data-parallel plumbing for the slot-attention path.

The path shards by batch of clips only (every clip is independent in forward and backward; the
reference does the same with DDP + DistributedSampler: slowfast/models/build.py:79-83,
slowfast/datasets/loader.py:97).  The single exchange step is the parameter-gradient all-reduce.
`FlatGradAllReduce` does it as ONE collective on one contiguous buffer (the module's gradients are
only ~1.5-3.4 MB, so the cost is launch latency, not bandwidth) instead of DDP's per-bucket calls.
Works with any torch.distributed backend (nccl on B200s, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_clips, rank, world):
    """Contiguous [lo, hi) clip range of `rank`; the first n_clips % world ranks take one extra clip."""
    base, extra = divmod(n_clips, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_reduce_flat_(flat, group=None, average=True):
    """ONE collective, in place, on a flat gradient buffer: sum (or mean) over the ranks of `group`."""
    if not (dist.is_available() and dist.is_initialized()):
        return flat
    world = dist.get_world_size(group)
    if world == 1:
        return flat
    if average and dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)     # averaged inside the collective (NVLS / ring), no extra pass
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.div_(world)
    return flat


class GradSync:
    """Gradient exchange of a standalone data-parallel `SlotAttentionVideo` (what DDP does for the reference,
    slowfast/models/build.py:79-83), fused into its backward: the CUDA library already writes every parameter
    gradient into one flat fp32 buffer, so the module all-reduces THAT buffer once, in place, on the backward's
    stream, and hands views of it to autograd: no bucket copies, one NCCL launch per step.

        sync = attach_grad_sync(module)          # after init_process_group; parameters must start identical
        slots, attn = module(x); loss.backward() # .grad now holds the rank-averaged gradients

    Do not combine with DistributedDataParallel on the same module (the gradients would be reduced twice); when the
    module is embedded in a DDP-wrapped STEVE, leave it detached and DDP handles it like any other sub-module."""

    def __init__(self, group=None, average=True):
        self.group, self.average = group, average

    def __call__(self, flat):
        return all_reduce_flat_(flat, self.group, self.average)


def attach_grad_sync(module, group=None, average=True, broadcast_parameters=True):
    """Enable the fused flat-gradient all-reduce on `module`; rank 0's parameters are broadcast first (as DDP does)."""
    if broadcast_parameters and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    module.grad_sync = GradSync(group, average)
    return module.grad_sync


class FlatGradAllReduce:
    """Average the gradients of `params` across ranks with a single all-reduce on a flat buffer."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.flat = None

    def __call__(self, weight=1.0):
        """grad_i <- sum_r weight_r * grad_i^(r); pass weight = local_clips / global_clips for a batch-mean."""
        if not self.params:
            return
        p0 = self.params[0]
        n = sum(p.numel() for p in self.params)
        if self.flat is None or self.flat.numel() != n or self.flat.device != p0.device:
            self.flat = torch.empty(n, dtype=torch.float32, device=p0.device)
        views, off = [], 0
        for p in self.params:
            v = self.flat[off:off + p.numel()]
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad.reshape(-1))
            views.append(v)
            off += p.numel()
        if weight != 1.0:
            self.flat.mul_(weight)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        for p, v in zip(self.params, views):
            if p.grad is None:
                p.grad = v.view_as(p).clone()
            else:
                p.grad.copy_(v.view_as(p))
