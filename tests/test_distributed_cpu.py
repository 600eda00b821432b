"""CPU, gloo, world_size 2: the data-parallel host logic (clip sharding + ONE flat gradient all-reduce)
reproduces single-process gradients on the concatenated batch.  The per-rank compute here is the
oracle's torch port (CPU); on the GPU box the same helper wraps the CUDA module (bench.py --gpus N)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from focus_b200.distributed import FlatGradAllReduce, GradSync, attach_grad_sync, shard_range


def test_shard_range_partitions_every_clip_once():
    for n in (1, 7, 64, 65):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from oracle import savi_numpy as O
    from oracle import savi_torch as OT
    K, D, Ds, M, blocks, heads, I, T, N, Bg = 4, 16, 16, 24, 1, 2, 2, 2, 40, 5
    P = {k: torch.from_numpy(v) for k, v in O.random_params(K, D, Ds, M, blocks, seed=1).items()}
    g = torch.Generator().manual_seed(0)
    x = torch.randn(Bg, T, N, D, generator=g, dtype=torch.float64)
    noise = torch.randn(Bg, K, Ds, generator=g, dtype=torch.float64)
    gs = torch.randn(Bg, T, K, Ds, generator=g, dtype=torch.float64)
    lo, hi = shard_range(Bg, rank, world)
    _, _, _, G = OT.forward_backward(P, x[lo:hi], noise[lo:hi], I, heads, gs[lo:hi] / (hi - lo))   # local batch-mean loss
    params = [torch.nn.Parameter(P[k].clone().float()) for k in P]
    for p, k in zip(params, P):
        p.grad = G[k].float().clone()
    FlatGradAllReduce(params)(weight=(hi - lo) / Bg)
    # the fused variant used by SlotAttentionVideo.backward: one in-place collective on the library's flat gradient buffer
    flat = torch.cat([G[k].float().reshape(-1) for k in P]) * ((hi - lo) * world / Bg)
    GradSync(average=True)(flat)
    flat_ref = torch.cat([p.grad.reshape(-1) for p in params])
    assert float((flat - flat_ref).abs().max()) <= 1e-6 * float(flat_ref.abs().max())
    # exchange selection: a gloo / CPU group cannot use the library's peer-memory kernel — "auto" falls back to the collective on
    # every rank (and broadcasts rank 0's parameters), peer=True refuses
    holder = torch.nn.Linear(3, 2)
    with torch.no_grad():
        holder.weight.fill_(float(rank + 1))
    assert isinstance(attach_grad_sync(holder, peer="auto"), GradSync) and float(holder.weight[0, 0]) == 1.0
    try:
        attach_grad_sync(holder, peer=True)
        raise AssertionError("peer=True must refuse a gloo group")
    except RuntimeError:
        pass
    if rank == 0:
        _, _, _, Gfull = OT.forward_backward(P, x, noise, I, heads, gs / Bg)
        worst = max(float((p.grad.double() - Gfull[k]).abs().max() / Gfull[k].abs().max().clamp_min(1e-30))
                    for p, k in zip(params, P) if Gfull[k].abs().max() > 1e-12)
        np.save(out, np.array([worst]))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_allreduce_equals_single_process_grads(tmp_path):
    out = str(tmp_path / "worst.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert float(np.load(out)[0]) < 1e-5
