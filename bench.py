#!/usr/bin/env python
"""bench.py — slot-attention fwd+bwd frames/sec on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA library)
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU arm (the unmodified reference module, host cores)

One "step" = one SlotAttentionVideo forward + backward over one batch of synthetic
MOVi-E-shaped clips (BASELINE.json configs[1], "C2": bf16, 64 clips/GPU, T=6, N=1024,
D=Ds=M=128, K=24, 3 iterations, 1 predictor block x 4 heads), upstream gradients for
both outputs, plus (N>1) the DDP gradient all-reduce.  One frame = one (clip, t) pair.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline     the dominant kernel against the measured HBM peak (MEASURED_PEAKS.json)
  step_roofline  whole-step k/v-bytes-x-iterations roofline of BASELINE.md §3
  cpu_baseline the unmodified reference module (oracle/_ref) timed on this box's host cores (N=1 only)
  kernels_ms   per-kernel CUDA-event times of the last timed step
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CONFIGS = {
    # name: B/GPU, T, N, D, Ds, M, K, I, blocks, heads, dtype
    "c1": dict(B=2, T=6, N=1024, D=128, Ds=128, M=128, K=24, I=3, blocks=1, heads=4, dtype="fp32"),
    "c2": dict(B=64, T=6, N=1024, D=128, Ds=128, M=128, K=24, I=3, blocks=1, heads=4, dtype="bf16"),
    "c3": dict(B=64, T=6, N=4096, D=192, Ds=192, M=192, K=24, I=3, blocks=1, heads=4, dtype="bf16"),
    "c4": dict(B=64, T=24, N=1024, D=128, Ds=128, M=128, K=11, I=2, blocks=1, heads=4, dtype="bf16"),
    # BASELINE configs[4] sweep corner in the throughput regime (tools/sweep.py; phase breakdowns with tools/phase_times.py n16k)
    "n16k": dict(B=64, T=2, N=16384, D=128, Ds=128, M=128, K=24, I=3, blocks=1, heads=4, dtype="bf16"),
}
METRIC = "slot-attn fwd+bwd frames/sec"
UNIT = "frames/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80)}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_params_like(model_cfg, seed=0):
    """Reference init (same RNG order) via our module's constructor, on CPU."""
    from focus_b200 import SlotAttentionVideo
    torch.manual_seed(seed)
    c = model_cfg
    return SlotAttentionVideo(c["I"], c["K"], c["D"], c["Ds"], c["M"], c["blocks"], c["heads"], 0.0)


# ---------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the UNMODIFIED reference module (oracle/_ref, made by oracle/make_ref.py from
# /root/reference; see oracle/_load_reference.py) on the host cores.  Nothing of focus_b200 is imported on this path:
# parameters come from the reference constructor under torch.manual_seed(0) (the same values make_params_like gives
# our module: the constructors are RNG-identical, tests/test_cabi_cpu.py).
# ---------------------------------------------------------------------------------------------
def _reference_module(c):
    from oracle import _load_reference as LR
    if not LR.reference_available():
        return None
    torch.manual_seed(0)
    return LR.reference_slot_attention_video(c["I"], c["K"], c["D"], c["Ds"], c["M"], c["blocks"], c["heads"], 0.0)


def _ref_step(ref, x, noise, gs, ga):
    """One forward + backward of the reference module with the slot noise injected (its single RNG draw, steve.py:56)."""
    x = x.detach().requires_grad_(True)
    orig = torch.Tensor.normal_
    torch.Tensor.normal_ = lambda self, *a, **k: self.copy_(noise)
    try:
        slots, attn = ref(x)
    finally:
        torch.Tensor.normal_ = orig
    for p in ref.parameters():
        p.grad = None
    torch.autograd.backward([slots, attn], [gs.to(slots.dtype), ga.to(attn.dtype)])
    return slots


def _ref_inputs(c, clips, device="cpu", dtype=torch.float32):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(clips, c["T"], c["N"], c["D"], generator=g).to(dtype).to(device)
    noise = torch.randn(clips, c["K"], c["Ds"], generator=g).to(device)
    gs = torch.randn(clips, c["T"], c["K"], c["Ds"], generator=g).to(device)
    ga = torch.randn(clips, c["T"], c["N"], c["K"], generator=g).to(dtype).to(device)
    return x, noise, gs, ga


def time_cpu_reference(cfg, sample_clips, steps, warmup):
    """frames/s of the reference's own CPU path (fp32, all host threads) on `sample_clips` clips of the workload."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    c = cfg
    ref = _reference_module(c)
    x, noise, gs, ga = _ref_inputs(c, sample_clips)
    if ref is not None:
        kind, what = "reference", "the unmodified reference module slowfast/models/STEVE/steve.py:SlotAttentionVideo (oracle/_ref)"
        run = lambda: _ref_step(ref, x, noise, gs, ga)
    else:
        # the reference sources did not travel (oracle/make_ref.py / __graft_entry__.build() not run where /root/reference exists):
        # time the oracle's torch port instead (same ATen operations in the reference's order) and SAY so
        from oracle import savi_numpy as O
        from oracle import savi_torch as OT
        kind = "port"
        what = "oracle/savi_torch.py (torch CPU ops in the reference's operation order; the reference sources were not found)"
        P = {k: torch.from_numpy(v).float() for k, v in O.random_params(c["K"], c["D"], c["Ds"], c["M"], c["blocks"], seed=0).items()}
        run = lambda: OT.forward_backward(P, x, noise, c["I"], c["heads"], gs, ga)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        run()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    t = statistics.median(times)
    return dict(value=sample_clips * c["T"] / t, unit=UNIT, cores=cores, kind=kind,
                sample="%d of %d clips/step, fp32 forward + backward (both upstream gradients), %s; median of %d steps after %d warm-up" %
                       (sample_clips, c["B"], what, steps, warmup)), t


def time_gpu_eager_reference(cfg, steps=10, warmup=3):
    """The reference module through PyTorch eager on the B200 (what FOCUS runs today): fp32 with TF32 off, and bf16 autocast.
    A reported baseline (SURVEY.md §8d "the real beat-this number"); never part of the product path."""
    out = {}
    dev = torch.device("cuda", 0)
    c = cfg
    for tag in ("fp32_tf32_off", "bf16_autocast"):
        ref = _reference_module(c)
        if ref is None:
            return None
        ref = ref.to(dev)
        x, noise, gs, ga = _ref_inputs(c, c["B"], dev)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if tag == "bf16_autocast" else torch.autocast("cuda", enabled=False)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        evs = []
        for i in range(warmup + steps):
            flush.fill_(i & 0xFF)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            with ctx:
                _ref_step(ref, x, noise, gs, ga)
            e1.record()
            if i >= warmup:
                evs.append((e0, e1))
        torch.cuda.synchronize(dev)
        ms = statistics.median(a.elapsed_time(b) for a, b in evs)
        out[tag] = {"ms_per_step": ms, "value": c["B"] * c["T"] / (ms * 1e-3), "unit": UNIT}
        del ref, x, noise, gs, ga, flush
        torch.cuda.empty_cache()
    out["what"] = ("unmodified reference module (oracle/_ref) via PyTorch eager on cuda:0, %d clips, forward + backward, CUDA events, "
                   "median of %d steps after %d warm-up, L2 flushed between steps" % (c["B"], steps, warmup))
    return out


def run_reference(args, cfg, rank):
    if rank != 0:
        return
    warm = max(0, args.warmup)
    cb, t = time_cpu_reference(cfg, cfg["B"], max(1, args.steps), warm)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": warm, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, cfg, args.gpus), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if args.gpu_eager and torch.cuda.is_available():
        line["gpu_eager_baseline"] = time_gpu_eager_reference(cfg)
    print(json.dumps(line), flush=True)


def config_dict(args, cfg, n):
    return {"workload": "BASELINE.json configs[1] (C2): STEVE MOVi-E 128x128 slot-attention fwd+bwd" if args.config == "c2"
            else "BASELINE.json config %s" % args.config,
            "clips_per_gpu": cfg["B"], "global_clips": cfg["B"] * n, "T": cfg["T"], "N": cfg["N"], "D": cfg["D"],
            "Ds": cfg["Ds"], "M": cfg["M"], "K": cfg["K"], "iters": cfg["I"], "predictor": "%d block x %d heads" % (cfg["blocks"], cfg["heads"]),
            "token_dtype": cfg["dtype"], "grad_attn": "dense N(0,1)",
            "parallelism": "dp%d" % n + ("" if n == 1 else " (torch DDP)" if args.ddp else " (gradient exchange inside backward: %s)" % getattr(args, "exchange", "one flat all-reduce")),
            "l2": "flushed between timed steps (256 MiB write outside the event brackets%s); step working set >> 126 MB L2" % ("" if n == 1 else "; ranks re-aligned on the device after the flush, before each start event")}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args, cfg, rank, world, local_rank):
    import torch.distributed as dist
    from focus_b200 import SlotAttentionVideo, _lib
    from focus_b200.slot_attention import _SaviFunction

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    # stdout carries exactly ONE JSON line: library chatter on fd 1 (e.g. NCCL's version banner) goes to stderr until then
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    c = cfg
    dt = torch.float32 if c["dtype"] == "fp32" else torch.bfloat16
    model = make_params_like(c).to(dev)
    ddp = model
    if world > 1 and args.ddp:       # stock DistributedDataParallel, what the reference wraps the model in (models/build.py:79-83)
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], output_device=local_rank,
                                                        gradient_as_bucket_view=True, bucket_cap_mb=64)
    elif world > 1:                  # the path's own exchange step: ONE all-reduce of the flat gradient buffer inside backward
        from focus_b200.distributed import PeerGradSync, attach_grad_sync
        sync = attach_grad_sync(model, peer=False if args.nccl_sync else "auto")
        args.exchange = ("library kernel over NVLink peer memory (savi_allreduce_peers%s)" % (", multimem.ld_reduce" if sync.uses_multicast else "")
                         if isinstance(sync, PeerGradSync) else "one flat NCCL all-reduce")
    g = torch.Generator(device="cpu").manual_seed(1 + rank)
    B, T, N, D, K, Ds = c["B"], c["T"], c["N"], c["D"], c["K"], c["Ds"]
    x_host = torch.randn(B, T, N, D, generator=g).to(dt).pin_memory()
    x = x_host.to(dev).requires_grad_(True)
    noise = torch.randn(B, K, Ds, generator=g).to(dev)
    gs = torch.randn(B, T, K, Ds, generator=g).to(dev)
    ga = torch.randn(B, T, N, K, generator=g).to(dt).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    gs_t = gs.to(dt)                                         # upstream gradient in the dtype the module returns

    def step(inp):
        model.zero_grad(set_to_none=True)                    # the trainer's optimizer.zero_grad() (tools/steve_train_net.py:110)
        slots, attn = ddp(inp, noise=noise)
        torch.autograd.backward([slots, attn], [gs_t, ga])
        return slots

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step(x)
        x.grad = None
    sync_all()

    # ---- timed region: value (inputs resident in HBM) ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    # N > 1: the L2 flush between timed steps (outside the event brackets) takes a different time on every rank, and a rank that
    # starts its step early then spends the difference waiting inside the step's gradient exchange.  A training loop has no flush
    # and stays in lock-step through the exchange, so the ranks are re-aligned ON THE DEVICE (a 4-byte all-reduce the stream
    # waits for, no host synchronisation) between the flush and the start event; the bracket still holds exactly one step.
    align = torch.zeros(1, device=dev) if world > 1 else None

    def align_ranks():
        if align is not None:
            dist.all_reduce(align)

    sync_all()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        align_ranks()
        evs[i][0].record()
        step(x)
        evs[i][1].record()
        x.grad = None
    sync_all()
    eager_ms = sum(a.elapsed_time(b) for a, b in evs)
    # ---- the same K steps replayed as ONE CUDA graph per step (stream capture of the module's forward + autograd backward,
    # incl. zero_grad and, for N > 1, the NCCL gradient all-reduce): same kernels, same work, no host launch gaps.
    graph_ms = None
    graph_err = None
    graph = None
    if not args.no_graph:
        try:
            cs = torch.cuda.Stream(dev)
            cs.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(cs):
                for _ in range(3):
                    step(x); x.grad = None
            torch.cuda.current_stream(dev).wait_stream(cs)
            sync_all()
            eager_slots = step(x).detach().clone(); x.grad = None
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=cs):
                g_slots = step(x)
            for _ in range(3):
                graph.replay()
            sync_all()
            if not torch.equal(g_slots, eager_slots):
                raise RuntimeError("graph replay does not reproduce the eager forward bit for bit")
            gevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            for i in range(args.steps):
                flush.fill_(i & 0xFF)
                align_ranks()
                gevs[i][0].record()
                graph.replay()
                gevs[i][1].record()
            sync_all()
            graph_ms = sum(a.elapsed_time(b) for a, b in gevs)
            x.grad = None
        except Exception as e:       # reported, never silent: the eager number stands
            graph_err = "%s: %s" % (type(e).__name__, str(e)[:200])
            graph_ms = None
            sync_all()
    clocks = sampler.stop()
    # ---- per-kernel CUDA-event times: the same step, 5 more times, with the library's event pairs around every launch
    # (on the launch stream).  A separate pass because an event record between two launches serialises them, and the
    # d_inputs kernel normally overlaps the backward clip kernel as its programmatic dependent launch.
    _lib.profile_enable(True)
    kern_ms = []
    for i in range(5):
        flush.fill_(i & 0xFF)
        step(x)
        x.grad = None
        kern_ms.append(_lib.profile_read())
    sync_all()
    _lib.profile_enable(False)
    launches_per_step = None
    # max over ranks; the graph number is used only when EVERY rank captured successfully
    t_local = torch.tensor([eager_ms, graph_ms if graph_ms is not None else float("inf")], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    eager_ms, graph_ms = float(t_local[0].item()), float(t_local[1].item())
    use_graph = graph_ms != float("inf") and graph_ms < eager_ms
    total_ms = graph_ms if use_graph else eager_ms
    ms_per_step = total_ms / args.steps
    frames = B * T * world
    value = frames / (ms_per_step * 1e-3)

    # ---- e2e: host (pinned) inputs -> H2D -> fwd+bwd -> D2H of the step's scalar result, every step ----
    # Every step's inputs are copied from pinned host memory inside the timed region; the copy of step i+1 is issued on a
    # second stream before step i's result is read back, so it overlaps step i's kernels (two device buffers).
    x_dev = [torch.empty_like(x_host, device=dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream(dev)
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    main_stream = torch.cuda.current_stream(dev)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(3, min(args.steps, 20))
    launches = [0, 0]

    def issue_copy(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i & 1])          # the buffer's previous user (step i-2) is done
            x_dev[i & 1].copy_(x_host, non_blocking=True)
            copied[i & 1].record(copy_stream)

    res_host = torch.zeros(64, dtype=torch.float32).pin_memory()
    res_done = [torch.cuda.Event() for _ in range(64)]

    def e2e_loop(nsteps):
        """Per step: H2D of that step's inputs (issued one step ahead on the copy stream), forward + backward through the
        module, D2H of the step's scalar result into pinned memory; the host reads result i-1 while step i runs, so the
        launch queue never drains (a training loop that logs its loss one step late)."""
        for ev in consumed:
            ev.record(main_stream)
        issue_copy(0)
        out = []
        for i in range(nsteps):
            if i + 1 < nsteps:
                issue_copy(i + 1)
            main_stream.wait_event(copied[i & 1])
            inp = x_dev[i & 1].detach().requires_grad_(True)
            model.zero_grad(set_to_none=True)
            slots, attn = ddp(inp, noise=noise)
            launches[0] = _SaviFunction.last_launches
            torch.autograd.backward([slots, attn], [gs_t, ga])
            launches[1] = _SaviFunction.last_launches
            consumed[i & 1].record(main_stream)
            res_host[i:i + 1].copy_((slots.detach().float() * gs).sum().reshape(1), non_blocking=True)     # D2H of the step's result
            res_done[i].record(main_stream)
            if i >= 1:
                res_done[i - 1].synchronize()
                out.append(float(res_host[i - 1]))
        res_done[nsteps - 1].synchronize()
        out.append(float(res_host[nsteps - 1]))
        return out

    e2e_loop(3)                                          # untimed: second stream, allocator blocks of the double buffer
    sync_all()
    e0.record()
    e2e_loop(e2e_steps)
    e1.record()
    sync_all()
    fwd_launches, bwd_launches = launches
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    t_local = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
    e2e_ms = float(t_local.item())
    # kernels the library launched per step, as counted by the library itself (savi_last_launch_count): parameter packing +
    # token LayerNorm + forward clip kernel, then backward clip kernel + weight gradients + d_inputs
    launches_per_step = fwd_launches + bwd_launches

    def shutdown():
        """Release the captured graph (it holds NCCL kernels when N > 1) BEFORE the process group goes away, and never let a
        stuck communicator teardown keep the job alive: the JSON line is already out, so a watchdog ends the process."""
        nonlocal graph
        if graph is not None:
            graph.reset()
            graph = None
        torch.cuda.synchronize(dev)
        if world > 1:
            sys.stdout.flush(); sys.stderr.flush()
            wd = threading.Timer(20.0, lambda: os._exit(0))
            wd.daemon = True
            wd.start()
            try:
                dist.barrier()
                dist.destroy_process_group()
            finally:
                wd.cancel()

    if rank != 0:
        shutdown()
        return
    # ---- roofline of the dominant kernel ----
    peak, peak_src = measured_peaks()
    e = 4 if c["dtype"] == "fp32" else 2
    km = {k: statistics.mean(d[k] for d in kern_ms) for k in kern_ms[0]}
    dom = max(("savi_fwd", "savi_bwd"), key=lambda k: km[k])
    algo_bytes_launch = B * T * c["I"] * 2 * N * Ds * e          # k/v bytes x iterations, one direction
    achieved = algo_bytes_launch / (km[dom] * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")      # dram__bytes_read + write per launch from the committed ncu --set full capture
    if os.path.exists(tp) and args.config == "c2" and not args.clips:
        with open(tp) as f:
            traffic = json.load(f).get(dom)
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes_launch,
                "kernel_ms": km[dom]}
    step_bytes = 4 * c["I"] * N * Ds * e * B * T
    step_roof = {"bytes_per_gpu_step": step_bytes, "roofline_ms": step_bytes / (peak * 1e9) * 1e3,
                 "frac": step_bytes / (peak * 1e9) * 1e3 / ms_per_step}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if c["dtype"] == "bf16" else "f32", "data": "synthetic", "config": config_dict(args, cfg, world),
            "clocks": clocks,
            "launch": {"mode": "cuda-graph replay (one graph = one step)" if use_graph else "eager (one Python call per step)",
                       "ms_per_step_eager": eager_ms / args.steps,
                       "ms_per_step_graph": None if graph_ms == float("inf") else graph_ms / args.steps, "graph_error": graph_err},
            "e2e": {"value": frames / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": x_host.numel() * x_host.element_size(), "d2h_bytes_per_step": 4},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "step_roofline": step_roof, "kernels_ms": km,
            "kernels_ms_note": "CUDA events around each launch in a separate 5-step pass right after the timed region (serialised: "
                               "in the timed steps d_inputs overlaps the backward clip kernel on the SMs it leaves idle)"}
    if world == 1 and not args.no_cpu_baseline:
        cb, _ = time_cpu_reference(cfg, min(cfg["B"], 16), 5, 1)      # bounded sample: ~10-20 s of host time
        line["cpu_baseline"] = cb
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--clips", type=int, default=0, help="override clips per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ddp", action="store_true", help="N>1: wrap the module in torch DDP instead of the fused flat all-reduce")
    ap.add_argument("--nccl-sync", action="store_true", help="N>1: NCCL all-reduce of the flat gradient buffer instead of the library's peer-memory kernel")
    ap.add_argument("--no-graph", action="store_true", help="do not try the CUDA-graph replay of the step")
    ap.add_argument("--gpu-eager", action="store_true", help="--impl reference only: also time the reference module through PyTorch eager on cuda:0")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.clips:
        cfg["B"] = args.clips
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    run_ours(args, cfg, rank, world, local_rank)


if __name__ == "__main__":
    main()
