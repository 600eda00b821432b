"""2+ GPUs under torchrun: the fused flat gradient all-reduce (attach_grad_sync, NCCL AVG inside backward) must reproduce the
gradients of ONE process on the concatenated batch with a batch-mean loss.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_grad_sync.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from focus_b200 import SlotAttentionVideo
from focus_b200.distributed import PeerGradSync, attach_grad_sync, shard_range

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
I, K, D, T, N, Bg = 3, 24, 128, 3, 512, 4 * world
torch.manual_seed(100 + rank)                      # ranks start from DIFFERENT parameters: attach_grad_sync must broadcast rank 0's
m = SlotAttentionVideo(I, K, D, D, D, 1, 4, 0.0).cuda()
g = torch.Generator().manual_seed(7)
x = torch.randn(Bg, T, N, D, generator=g).bfloat16().cuda()
noise = torch.randn(Bg, K, D, generator=g).cuda()
gs = torch.randn(Bg, T, K, D, generator=g).cuda()
lo, hi = shard_range(Bg, rank, world)
ref_grads = None
# the NCCL all-reduce, then the library's own peer-memory kernel with and without the NVSwitch multicast load
for mode in os.environ.get("GRAD_SYNC_MODES", "nccl,peer,peer-p2p").split(","):
    sync = attach_grad_sync(m, peer=(mode != "nccl"))
    if mode != "nccl" and not isinstance(sync, PeerGradSync):
        if rank == 0:
            print("world %d: %s: peer-memory sync unavailable on this box (fell back to %s)" % (world, mode, type(sync).__name__))
        continue
    if mode == "peer-p2p":
        sync.multicast = False; sync.buf = None     # re-rendezvous without the multicast mapping
    for step in range(2):                           # twice: the symmetric buffer and the signal pads are reused
        m.zero_grad(set_to_none=True)
        s, a = m(x[lo:hi], noise=noise[lo:hi])
        (s.float() * gs[lo:hi]).sum().div(hi - lo).backward()       # local batch-mean loss; AVG over ranks = global batch mean (equal shards)
    got = {n: p.grad.clone() for n, p in m.named_parameters()}
    if rank == 0:
        if ref_grads is None:
            ref = SlotAttentionVideo(I, K, D, D, D, 1, 4, 0.0).cuda()
            ref.load_state_dict(m.state_dict())
            s2, _ = ref(x, noise=noise)
            (s2.float() * gs).sum().div(Bg).backward()
            ref_grads = {n: p.grad for n, p in ref.named_parameters()}
        gmax = max(float(v.abs().max()) for v in ref_grads.values())
        worst = max(float((got[n] - v).abs().max()) / gmax for n, v in ref_grads.items())
        how = type(sync).__name__ + (" (multimem.ld_reduce)" if getattr(sync, "uses_multicast", False) else "")
        print("world %d: %s [%s] vs single process on the concatenated batch: max-normalised gradient difference %.2e" % (world, mode, how, worst))
        assert worst < 1e-5
    dist.barrier()
dist.destroy_process_group()
