"""Data-parallel plumbing for the slot-attention path.

The path shards by batch of clips only (every clip is independent in forward and backward; the
reference does the same with DDP + DistributedSampler: slowfast/models/build.py:79-83,
slowfast/datasets/loader.py:97).  The single exchange step is the parameter-gradient all-reduce, on ONE
contiguous buffer (the module's gradients are only ~1.5-3.4 MB, so the cost is latency, not bandwidth):
  * `PeerGradSync` — the library's own kernel over NVLink / NVSwitch peer memory (csrc/savi_allreduce.cu,
    include/focus_savi.h: savi_allreduce_peers): the CUDA backward writes its flat gradient buffer straight into
    symmetric memory and one kernel per rank sums all peers' buffers (one `multimem.ld_reduce` through the switch where
    the allocation has a multicast mapping).  One node, NCCL process group, <= 16 ranks;
  * `GradSync` — one NCCL (or gloo, in the CPU tests) all-reduce of the same buffer: the fallback and the multi-node path.
"""
import ctypes
import os
import warnings

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


class PeerGradSync:
    """The same exchange as `GradSync`, as ONE kernel of this library over peer memory instead of an NCCL call.

    `buffer(n, device)` hands the backward a flat fp32 gradient buffer in symmetric memory (allocated and exchanged with the
    peers on first use: a collective, so every rank must reach its first backward); `__call__(flat)` launches
    savi_allreduce_peers on the current stream and returns a NEW local tensor with the rank-averaged (or summed) gradients —
    the symmetric buffer itself is reused by the next step, autograd keeps views of the returned one."""

    def __init__(self, group=None, average=True, multicast=True):
        self.group = group if group is not None else dist.group.WORLD
        self.average, self.multicast = average, multicast
        self.buf = self.hdl = None

    def buffer(self, n, device):
        cap = (n + 3) // 4 * 4                              # the kernel moves 16-byte vectors
        if self.buf is None or self.buf.numel() < cap or self.buf.device != device:
            import torch.distributed._symmetric_memory as symm
            self.buf = symm.empty(cap, dtype=torch.float32, device=device)
            self.buf.zero_()
            self.hdl = symm.rendezvous(self.buf, self.group)
            w = self.hdl.world_size
            vp = ctypes.c_void_p
            self._bufs = (vp * w)(*[int(p) for p in self.hdl.buffer_ptrs])
            self._pads = (vp * w)(*[int(p) for p in self.hdl.signal_pad_ptrs])
            mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0) if self.multicast else 0     # 0: no NVLS mapping for this allocation
            self._mc = vp(mc if mc else None)
            self._pad_bytes = int(symm.get_signal_pad_size())
        return self.buf[:n]

    @property
    def uses_multicast(self):
        return self.hdl is not None and bool(self._mc.value)

    def __call__(self, flat):
        from . import _lib
        if self.hdl is None or flat.data_ptr() != self.buf.data_ptr():
            raise RuntimeError("PeerGradSync: the gradient buffer must be the symmetric one returned by buffer()")
        n, cap = flat.numel(), (flat.numel() + 3) // 4 * 4
        out = torch.empty(cap, dtype=torch.float32, device=flat.device)     # ordinary device memory: autograd keeps views of it
        stream = ctypes.c_void_p(torch.cuda.current_stream(flat.device).cuda_stream)
        scale = 1.0 / self.hdl.world_size if self.average else 1.0
        with torch.cuda.device(flat.device):
            _lib.check(_lib.lib.savi_allreduce_peers(self._bufs, self._pads, self._mc, self.hdl.rank, self.hdl.world_size,
                                                     ctypes.c_void_p(out.data_ptr()), cap, scale, self._pad_bytes, stream),
                       "savi_allreduce_peers")
        return out[:n]


def _peer_sync_possible(module, group):
    """One node, NCCL, CUDA parameters, <= 16 ranks: the conditions of the peer-memory kernel."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_backend(group) != "nccl":
        return False
    world = dist.get_world_size(group)
    if world < 2 or world > 16 or world > torch.cuda.device_count():
        return False
    p = next(module.parameters(), None)
    return p is not None and p.is_cuda


def attach_grad_sync(module, group=None, average=True, broadcast_parameters=True, peer="auto"):
    """Enable the fused flat-gradient all-reduce on `module`; rank 0's parameters are broadcast first (as DDP does).
    peer = "auto": the library's peer-memory kernel when the group is one node of NVLink-connected GPUs under NCCL (all ranks
    must take the same branch: the decision is agreed with one small all-reduce), else the NCCL / gloo all-reduce;
    True: require the peer kernel; False (or FOCUS_SAVI_PEER_SYNC=0): always the collective library."""
    if broadcast_parameters and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if os.environ.get("FOCUS_SAVI_PEER_SYNC") == "0":
        peer = False
    sync = None
    if peer and _peer_sync_possible(module, group):
        dev = next(module.parameters()).device
        ok = torch.ones(1, device=dev)
        try:                                                # allocate + rendezvous now, so a failure is seen by every rank here
            cand = PeerGradSync(group, average)
            cand.buffer(sum(p.numel() for p in module.parameters()), dev)
        except Exception as e:                              # no symmetric memory on this system: fall back, on all ranks
            if peer is True:
                raise
            warnings.warn("focus_b200: peer-memory gradient sync unavailable (%s: %s); using the %s all-reduce"
                          % (type(e).__name__, str(e)[:120], dist.get_backend(group)))
            ok.zero_()
            cand = None
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if ok.item() > 0:
            sync = cand
        elif peer is True:
            raise RuntimeError("peer-memory gradient sync is not available on every rank")
    elif peer is True:
        raise RuntimeError("peer-memory gradient sync needs an NCCL group of 2-16 CUDA ranks on one node")
    module.grad_sync = sync if sync is not None else GradSync(group, average)
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
