"""2+ GPUs under torchrun: the fused flat gradient all-reduce (attach_grad_sync, NCCL AVG inside backward) must reproduce the
gradients of ONE process on the concatenated batch with a batch-mean loss.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_grad_sync.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from focus_b200 import SlotAttentionVideo
from focus_b200.distributed import attach_grad_sync, shard_range

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
I, K, D, T, N, Bg = 3, 24, 128, 3, 512, 4 * world
torch.manual_seed(100 + rank)                      # ranks start from DIFFERENT parameters: attach_grad_sync must broadcast rank 0's
m = SlotAttentionVideo(I, K, D, D, D, 1, 4, 0.0).cuda()
attach_grad_sync(m)
g = torch.Generator().manual_seed(7)
x = torch.randn(Bg, T, N, D, generator=g).bfloat16().cuda()
noise = torch.randn(Bg, K, D, generator=g).cuda()
gs = torch.randn(Bg, T, K, D, generator=g).cuda()
lo, hi = shard_range(Bg, rank, world)
s, a = m(x[lo:hi], noise=noise[lo:hi])
(s.float() * gs[lo:hi]).sum().div(hi - lo).backward()           # local batch-mean loss; AVG over ranks = global batch mean (equal shards)
got = {n: p.grad.clone() for n, p in m.named_parameters()}
if rank == 0:
    ref = SlotAttentionVideo(I, K, D, D, D, 1, 4, 0.0).cuda()
    ref.load_state_dict(m.state_dict())
    s2, _ = ref(x, noise=noise)
    (s2.float() * gs).sum().div(Bg).backward()
    gmax = max(float(p.grad.abs().max()) for p in ref.parameters())
    worst = max(float((got[n] - p.grad).abs().max()) / gmax for n, p in ref.named_parameters())
    print("world %d: fused flat all-reduce vs single process on the concatenated batch: max-normalised gradient difference %.2e" % (world, worst))
    assert worst < 1e-5
dist.barrier()
dist.destroy_process_group()
