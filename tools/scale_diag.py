"""Multi-GPU diagnosis under torchrun: per-rank step time WITHOUT any exchange (is one GPU of the box slower?) and the cost of
the gradient exchange alone (library peer kernel with / without the multicast load, NCCL), 100 calls back to back per rank.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/scale_diag.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
from focus_b200.distributed import GradSync, PeerGradSync

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
c = dict(bench.CONFIGS["c2"])
m = bench.make_params_like(c).to(dev)
g = torch.Generator().manual_seed(1 + rank)
x = torch.randn(c["B"], c["T"], c["N"], c["D"], generator=g).bfloat16().to(dev).requires_grad_(True)
noise = torch.randn(c["B"], c["K"], c["Ds"], generator=g).to(dev)
gs = torch.randn(c["B"], c["T"], c["K"], c["Ds"], generator=g).bfloat16().to(dev)
ga = torch.randn(c["B"], c["T"], c["N"], c["K"], generator=g).bfloat16().to(dev)


def step():
    m.zero_grad(set_to_none=True)
    s, a = m(x, noise=noise)
    torch.autograd.backward([s, a], [gs, ga]); x.grad = None


def timed(fn, n):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


out = {"rank": rank, "step_no_exchange_ms": timed(step, 30)}
n = sum(p.numel() for p in m.parameters())
for mode in ("peer-multicast", "peer-p2p", "nccl"):
    if mode == "nccl":
        flat = torch.randn(n, device=dev); sync = GradSync()
        out[mode + "_us"] = 1e3 * timed(lambda: sync(flat), 100)
    else:
        sync = PeerGradSync(multicast=(mode == "peer-multicast"))
        flat = sync.buffer(n, dev); flat.normal_()
        if mode == "peer-multicast" and not sync.uses_multicast:
            continue
        out[mode + "_us"] = 1e3 * timed(lambda: sync(flat), 100)
rows = [None] * world
dist.all_gather_object(rows, out)
if rank == 0:
    for r in rows:
        print(" ".join("%s=%s" % (k, ("%.3f" % v if isinstance(v, float) else v)) for k, v in r.items()))
dist.barrier()
dist.destroy_process_group()
