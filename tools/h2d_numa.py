"""Where does the end-to-end (host-resident inputs) step stop scaling at 4-8 ranks?  Every rank copies its 100.7 MB C2 input batch
from pinned host memory per step; this tool measures the AGGREGATE host->device bandwidth of all ranks copying concurrently,
for several ways of placing the pinned buffer, and prints the box's GPU / NUMA topology.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_numa.py"""
import ctypes, glob, os, subprocess, sys
import torch
import torch.distributed as dist

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 64 * 6 * 1024 * 128 * 2                     # bytes of one C2 input batch (bf16)
dev = torch.device("cuda", lr)
d = torch.empty(n, dtype=torch.uint8, device=dev)


def gpu_numa_node(i):
    try:
        bus = torch.cuda.get_device_properties(i).pci_bus_id
        dom = torch.cuda.get_device_properties(i).pci_domain_id
        devid = torch.cuda.get_device_properties(i).pci_device_id
        p = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, devid)
        return int(open(p).read())
    except Exception as e:
        return None


def node_cpus(node):
    try:
        txt = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        out = []
        for part in txt.split(","):
            a, _, b = part.partition("-")
            out += list(range(int(a), int(b or a) + 1))
        return out
    except Exception:
        return None


def measure(h, label):
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%-44s slowest rank %.3f ms -> aggregate %.1f GB/s (%.1f GB/s per rank)" % (label, t.item(), world * n / t.item() / 1e6, n / t.item() / 1e6), flush=True)


if rank == 0:
    print("world", world, "| cpus", os.cpu_count(), "| affinity of rank 0:", len(os.sched_getaffinity(0)), "cpus", flush=True)
    for cmd in (["nvidia-smi", "topo", "-m"], ["numactl", "-H"], ["lscpu"]):
        try:
            print(subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout[:3000], flush=True)
        except Exception as e:
            print(cmd, "unavailable:", e, flush=True)
    print("GPU -> NUMA node (sysfs):", [gpu_numa_node(i) for i in range(torch.cuda.device_count())], flush=True)
    print("nodes:", sorted(glob.glob("/sys/devices/system/node/node*")), flush=True)

# (a) torch pin_memory as bench.py does today (allocated wherever the process happens to run)
h = torch.empty(n, dtype=torch.uint8).pin_memory()
measure(h, "(a) torch pin_memory, default affinity")
del h
# (b) bind the process to the CPUs of the GPU's NUMA node first, then allocate + first-touch + pin
node = gpu_numa_node(lr)
cpus = node_cpus(node) if node is not None and node >= 0 else None
if cpus:
    try:
        os.sched_setaffinity(0, cpus)
    except Exception as e:
        if rank == 0:
            print("sched_setaffinity failed:", e)
h = torch.empty(n, dtype=torch.uint8)
h.fill_(1)                                       # first touch on the bound node
h = h.pin_memory()
measure(h, "(b) affinity = GPU's NUMA node, first touch")
del h
# (c) write-combined pinned memory (cudaHostAllocWriteCombined): no CPU cache snooping on the PCIe reads
try:
    rt = ctypes.CDLL("libcudart.so")
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(0x04))
    assert rc == 0, rc
    buf = (ctypes.c_uint8 * n).from_address(p.value)
    h = torch.frombuffer(buf, dtype=torch.uint8)
    measure(h, "(c) cudaHostAlloc write-combined" + (" + node affinity" if cpus else ""))
    del h, buf
    rt.cudaFreeHost(p)
except Exception as e:
    if rank == 0:
        print("(c) write-combined failed:", e)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
