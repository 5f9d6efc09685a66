#!/usr/bin/env python
"""Benchmark of the CSPN hot path (BASELINE.json metric: CSPN-24 forward Mpixel/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one forward of the hot path over one batch of synthetic input: at N=1 the workload is
BASELINE.json configs[1] (batch 8, NYU 304x228, 3x3, 24 iterations, fp32, forward only, mode CSPN_new).
With N>1 (launched by torchrun, one rank per GPU) every rank runs the same per-GPU batch on its own
batch slice of the global batch (weak scaling, no collective on the CSPN path); the time is the MAX over
ranks and `value` the whole-job throughput.

Timing: CUDA events on the launching stream around exactly K steps (torch's current stream is the stream
the C ABI launches on), barrier + synchronize on both sides; inputs rotate through enough independent
sets that their footprint exceeds the 126 MB L2, so every step reads its inputs from HBM.

The JSON line also carries:
  roofline      dominant (only) kernel: algorithmic bytes per launch / measured launch time vs measured HBM peak
  cpu_baseline  the reference's CPU path on the host cores: the UNMODIFIED reference module staged under baseline/_ref
                (baseline/fetch_ref.py; kind "reference") or, if that is missing, its op-for-op port oracle/torch_port.py
  e2e           the same metric through the host-buffer C-ABI entry points (pinned host buffers; H2D + kernel + D2H of every
                step inside the timed region; cspn_fwd_host_submit_* keeps three calls in flight)
  max_abs       largest deviation of the timed configuration's output from the C oracle (BASELINE.json: "max-abs vs ref")
  extra         the other BASELINE sizes (KITTI 1216x352 fp16 forward and forward+backward, 5x5 PAC variant, NYU
                forward+backward) on every rank - aggregated over ranks like `value` - and context rows (eager launch cost,
                the reference module itself on the B200)

`--impl reference` times the reference's own CPU implementation on the same config and prints the same line with
"impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "CSPN-24 forward Mpixel/s"
UNIT = "Mpixel/s"
NYU = dict(B=8, H=228, W=304, iters=24, ksize=3, mode=0, dtype="f32")          # BASELINE.json configs[1]
KITTI = dict(B=32, H=352, W=1216, iters=24, ksize=3, mode=0, dtype="f16")      # configs[2] (forward part)
PAC5 = dict(B=16, H=480, W=640, iters=12, ksize=5, mode=1, dtype="f32")        # configs[3]
L2_BYTES = 126e6
BYTES_PER_PX = {("f32", 3): 44.0, ("f16", 3): 22.0, ("f32", 5): 108.0, ("f16", 5): 54.0}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-extra", action="store_true", help="skip the context configs (KITTI / 5x5)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--eager", action="store_true", help="launch every step from Python instead of replaying CUDA graphs")
    return ap.parse_args()


def synth(cfg, seed, pinned=False):
    """Synthetic inputs of SURVEY.md 8(d): randn guidance, depth in [0,10), NYU sparse density 500/69312."""
    g = torch.Generator().manual_seed(seed)
    b, h, w = cfg["B"], cfg["H"], cfg["W"]
    cg = cfg["ksize"] ** 2 - 1
    guidance = torch.randn(b, cg, h, w, generator=g)
    depth = torch.rand(b, 1, h, w, generator=g) * 10.0
    mask = torch.rand(b, 1, h, w, generator=g) < (500.0 / 69312.0)
    sparse = mask * (torch.rand(b, 1, h, w, generator=g) * 10.0 + 0.1)
    dt = torch.float16 if cfg["dtype"] == "f16" else torch.float32
    out = [t.to(dt).contiguous() for t in (guidance, depth, sparse)]
    return [t.pin_memory() for t in out] if pinned else out


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.reasons.update(k for k, bit in names.items() if r & bit)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self.thread:
            self.thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def bind_to_gpu_numa_node(index):
    """Pin this process (and thereby its first-touch host allocations, e.g. the pinned staging arena) to the CPUs that
    sit on the same NUMA node as GPU `index`: with 8 ranks the H2D copies otherwise cross the socket interconnect.
    Best effort - returns a short description or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dev_dir = "/sys/bus/pci/devices/" + bus.lower()[-12:]
        with open(dev_dir + "/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        with open(dev_dir + "/numa_node") as f:
            node = f.read().strip()
        return f"numa node {node}, {len(cpus)} cpus"
    except Exception:
        return None


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload):
    """DRAM bytes per launch from the committed ncu --set full capture, if one exists for this workload."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


def run_module(cfg, sets):
    from cspn_monodepth_b200 import cspn_new, cspn_ours
    mod = cspn_new.AffinityPropagate(cfg["iters"], 3) if cfg["mode"] == 0 else cspn_ours.AffinityPropagate(cfg["iters"])

    def step(i):
        g, d, s = sets[i % len(sets)]
        return mod(g, d, s) if cfg["mode"] == 0 else mod(d, g, sparse_depth=s)
    step.inputs = sets
    return step


def n_input_sets(cfg):
    px_bytes = BYTES_PER_PX[(cfg["dtype"], cfg["ksize"])] * cfg["B"] * cfg["H"] * cfg["W"]
    nsets = max(2, int(np.ceil(1.5 * L2_BYTES / px_bytes)))     # rotating footprint >= 1.5 x L2
    if os.environ.get("CSPN_BENCH_WARM_L2"):                    # experiments only
        nsets = 1
    return nsets


def device_sets(cfg, dev, rank):
    nsets = n_input_sets(cfg)
    return [[t.to(dev) for t in synth(cfg, 1000 * rank + i)] for i in range(nsets)], nsets


def time_device(cfg, dev, rank, steps, warmup, dist=None, sampler=None, graph=True):
    """K timed steps with CUDA events; returns (ms_total_max_over_ranks, launches_per_step, nsets).

    The K steps are captured once into CUDA graphs (chunks of <= 512 steps) and the timed region replays them:
    the kernel takes ~20 us while an eager Python call costs more than that on the host, so eager launches
    would time the host, not the GPU.  The work on the device is identical.
    """
    from cspn_monodepth_b200 import _lib
    lib = _lib.load()
    sets, nsets = device_sets(cfg, dev, rank)
    step = run_module(cfg, sets)
    with torch.no_grad():
        for i in range(max(3, warmup)):
            step(i)
        launches = lib.cspn_last_launch_count()
        torch.cuda.synchronize(dev)
        graphs = []
        if graph:
            side = torch.cuda.Stream(dev)
            done = 0
            while done < steps:
                n = min(512, steps - done)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    with torch.cuda.graph(g, stream=side):
                        for i in range(done, done + n):
                            step(i)
                graphs.append(g)
                done += n
            for g in graphs[:1]:
                g.replay()                      # warm the graph itself
            torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx = sampler if sampler is not None else _Null()
        with ctx:
            e0.record()
            if graph:
                for g in graphs:
                    g.replay()
            else:
                for i in range(steps):
                    step(i)
            e1.record()
            torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, launches, nsets


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def time_e2e(cfg, dev, steps, dist=None):
    """The same metric through the host-buffer C-ABI call: pinned host inputs, H2D + kernel + D2H every step."""
    from cspn_monodepth_b200 import _lib
    lib = _lib.load()
    fn = lib.cspn_fwd_host_f32 if cfg["dtype"] == "f32" else lib.cspn_fwd_host_f16
    # Host buffers are carved out of ONE large pinned staging arena, the way a loader would stage batches: on these boxes
    # a 22 MiB pinned allocation of its own feeds the copy engine at 13-24 GB/s, a slice of a 256 MiB pinned arena at
    # 41-55 GB/s (tools/h2d_bw.py, profiles/r01_h2d_bandwidth.txt).
    plain = [synth(cfg, 77 + i) for i in range(2)]
    need = sum(t.numel() * t.element_size() + 4096 for s_ in plain for t in s_) + 2 * (plain[0][1].numel() * plain[0][1].element_size() + 4096)
    arena = torch.empty(max(need, 256 << 20), dtype=torch.uint8).pin_memory()
    cursor = [0]

    def carve(like):
        n = like.numel() * like.element_size()
        view = arena[cursor[0]:cursor[0] + n].view(like.dtype).view(like.shape)
        cursor[0] += (n + 4095) // 4096 * 4096
        return view
    sets = []
    for s_ in plain:
        views = [carve(t) for t in s_]
        for v, t in zip(views, s_):
            v.copy_(t)
        sets.append(views)
    outs = [carve(s_[1]) for s_ in sets]
    b, h, w = cfg["B"], cfg["H"], cfg["W"]
    cg = cfg["ksize"] ** 2 - 1
    stream = torch.cuda.current_stream(dev).cuda_stream

    depth = lib.cspn_host_pipeline_depth()
    sub = lib.cspn_fwd_host_submit_f32 if cfg["dtype"] == "f32" else lib.cspn_fwd_host_submit_f16
    outs += [carve(sets[0][1]) for _ in range(max(0, depth - len(outs)))]
    import ctypes
    tickets = [ctypes.c_int(0) for _ in range(depth)]

    def call(i):
        g, d, s = sets[i % 2]
        _lib.check(fn(g.data_ptr(), cg * h * w, d.data_ptr(), s.data_ptr(), 1, outs[i % depth].data_ptr(),
                      b, 1, h, w, cfg["iters"], cfg["ksize"], cfg["mode"], stream))

    def run_pipelined(n):
        """n calls with up to `depth` in flight; every result is waited for (it is in host memory) before its slot is reused."""
        for i in range(n):
            k = i % depth
            if i >= depth:
                _lib.check(lib.cspn_host_wait(tickets[k].value))
            g, d, s = sets[i % 2]
            _lib.check(sub(g.data_ptr(), cg * h * w, d.data_ptr(), s.data_ptr(), 1, outs[k].data_ptr(),
                           b, 1, h, w, cfg["iters"], cfg["ksize"], cfg["mode"], ctypes.byref(tickets[k])))
        for k in range(min(n, depth)):
            _lib.check(lib.cspn_host_wait(tickets[k].value))
    for i in range(3):
        call(i)
    run_pipelined(2 * depth)
    check = float(outs[0].float().mean())
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()

    def median_of_blocks(fn_block):
        blocks = []
        for _ in range(3):                  # median of three blocks of `steps` calls: host-side copies are noisy on shared boxes
            t0 = time.perf_counter()
            fn_block()
            blocks.append((time.perf_counter() - t0) * 1e3 / steps)
        ms = statistics.median(blocks)
        if dist is not None:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms
    # host wall clock: every call returns only when its result is in host memory (the timed region starts and ends with an
    # idle device and fully delivered results, so wall time = device + copy time)
    ms_sync = median_of_blocks(lambda: [call(i) for i in range(steps)])
    ms_pipe = median_of_blocks(lambda: run_pipelined(steps))
    esz = sets[0][0].element_size()
    h2d = (cg + 2) * b * h * w * esz
    d2h = b * h * w * esz
    return ms_pipe, ms_sync, h2d, d2h, check, depth


def time_fwd_bwd(cfg, dev, steps):
    """Forward + backward of one batch through the raw C ABI (cspn_fwd_* then cspn_bwd_*), graph-replayed, rotating inputs.
    Returns (ms per step, kernel launches per step)."""
    from cspn_monodepth_b200 import _lib
    lib = _lib.load()
    b, h, w, it, k, mode = cfg["B"], cfg["H"], cfg["W"], cfg["iters"], cfg["ksize"], cfg["mode"]
    cg = k * k - 1
    sfx = cfg["dtype"]
    px_bytes = BYTES_PER_PX[(cfg["dtype"], k)] * b * h * w
    nsets = max(2, int(np.ceil(1.5 * L2_BYTES / px_bytes)))
    sets = [[t.to(dev) for t in synth(cfg, 300 + i)] for i in range(nsets)]
    gout = torch.randn_like(sets[0][1])
    out, gg, gd = torch.empty_like(sets[0][1]), torch.empty_like(sets[0][0]), torch.empty_like(sets[0][1])
    nf, nb = lib.cspn_fwd_workspace_bytes(b, 1, h, w, it, k, mode), lib.cspn_bwd_workspace_bytes(b, 1, h, w, it, k, mode)
    wsf = torch.empty(max(nf, 16), dtype=torch.uint8, device=dev)
    wsb = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
    fwd_fn, bwd_fn = getattr(lib, "cspn_fwd_" + sfx), getattr(lib, "cspn_bwd_" + sfx)
    launches = [0]

    def step(i):
        g, d, s = sets[i % nsets]
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(fwd_fn(g.data_ptr(), cg * h * w, d.data_ptr(), s.data_ptr(), 1, out.data_ptr(), b, 1, h, w, it, k, mode, wsf.data_ptr(), nf, stream))
        n = lib.cspn_last_launch_count()
        _lib.check(bwd_fn(gout.data_ptr(), g.data_ptr(), cg * h * w, cg, d.data_ptr(), s.data_ptr(), 1, gg.data_ptr(), gd.data_ptr(),
                          b, 1, h, w, it, k, mode, wsb.data_ptr(), nb, stream))
        launches[0] = n + lib.cspn_last_launch_count()
    for i in range(3):
        step(i)
    torch.cuda.synchronize(dev)
    side = torch.cuda.Stream(dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for i in range(steps):
                step(i)
    graph.replay()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); graph.replay(); e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / steps, launches[0]


def reference_forward(cfg):
    """(callable(g, d, s) -> out, kind, description) of the reference's own implementation of the path: the UNMODIFIED module
    from baseline/_ref (staged by baseline/fetch_ref.py) when it is there, else the op-for-op port."""
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if cfg["mode"] == 0 and os.path.isdir(os.path.join(ref_root, "network")):
        if ref_root not in sys.path:
            sys.path.insert(0, ref_root)
        from network.libs.post_process import CSPN_new as ref_new          # the reference's own file, unmodified
        mod = ref_new.AffinityPropagate(cfg["iters"], 3)
        return (lambda g, d, s: mod(g, d, s)), "reference", "baseline/_ref/network/libs/post_process/CSPN_new.py (the unmodified reference module)"
    from oracle import torch_port
    if cfg["mode"] == 0:
        return (lambda g, d, s: torch_port.mode_a_forward(g, d, s, cfg["iters"])), "port", "oracle/torch_port.py (op-for-op PyTorch restatement of CSPN_new.py:26-128)"
    return (lambda g, d, s: torch_port.mode_b_forward(d, g, s, cfg["iters"])), "port", "oracle/torch_port.py (op-for-op PyTorch restatement of CSPN_ours.py:24-54)"


def cpu_reference(cfg, min_seconds=4.0, max_steps=8, min_steps=5, batch=None, threads=None):
    """Reference CPU path on a bounded sample; returns (dict for `cpu_baseline`, list of step times)."""
    b = batch or cfg["B"]
    small = dict(cfg, B=b, dtype="f32")                 # the reference is fp32-only (CSPN_new.py:122)
    g, d, s = synth(small, 5)
    fn, kind, what = reference_forward(cfg)
    fwd = lambda: fn(g, d, s)
    cores = os.cpu_count() or 1
    best = None
    with torch.no_grad():
        for t in ([threads] if threads else sorted({1, max(1, cores // 2), cores})):
            torch.set_num_threads(t)
            fwd()
            t0 = time.perf_counter(); fwd(); dt = time.perf_counter() - t0
            if best is None or dt < best[1]:
                best = (t, dt)
        torch.set_num_threads(best[0])
        times = []
        t_start = time.perf_counter()
        while len(times) < max_steps and (len(times) < min_steps or time.perf_counter() - t_start < min_seconds):
            t0 = time.perf_counter(); fwd(); times.append(time.perf_counter() - t0)
    px = b * cfg["H"] * cfg["W"]
    med = statistics.median(times)
    return {"value": px / med / 1e6, "unit": UNIT, "cores": best[0], "host_cores": cores, "kind": kind,
            "sample": f"{len(times)} forwards of batch {b} x {cfg['W']}x{cfg['H']} fp32, {cfg['iters']} iterations, {what}, torch {torch.__version__}, "
                      + (f"{best[0]} threads" if threads else f"{best[0]} threads (fastest of 1/{max(1, cores // 2)}/{cores})"),
            "ms_per_step": med * 1e3}, times


def cpu_c_oracle(cfg):
    """Stronger CPU baseline for context: the fused C oracle with OpenMP over images."""
    from oracle import c_oracle
    g, d, s = (t.numpy() for t in synth(dict(cfg, dtype="f32"), 6))
    c_oracle.forward(g, d, s, cfg["iters"], cfg["ksize"], cfg["mode"], threads=0)
    t0 = time.perf_counter(); c_oracle.forward(g, d, s, cfg["iters"], cfg["ksize"], cfg["mode"], threads=0); dt = time.perf_counter() - t0
    return {"value": cfg["B"] * cfg["H"] * cfg["W"] / dt / 1e6, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": "1 forward, oracle/cspn_oracle.c (fused C restatement, OpenMP over images)"}


def main_reference(args, rank):
    """The reference's own CPU implementation on the bench configuration: W warm-up + K timed forwards (K >= 5 so that the
    median is one), each a bounded sample = one batch of the workload; rank 0 only."""
    if rank != 0:
        return
    cfg = NYU
    steps, warm = max(5, args.steps if args.steps < 2000 else 5), max(0, min(args.warmup, 3))
    steps = min(steps, 20)                               # ~0.25 s per forward: the whole arm stays within a minute
    res, times = cpu_reference(cfg, min_seconds=0.0, max_steps=steps + warm, min_steps=steps + warm)
    times = times[warm:]
    ms = statistics.median(times) * 1e3
    value = cfg["B"] * cfg["H"] * cfg["W"] / (ms * 1e-3) / 1e6
    res["value"] = value
    res["ms_per_step"] = ms
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times), "warmup": warm,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, n_input_sets(cfg)), "cpu_baseline": res,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(cfg, nsets):
    es = 4 if cfg["dtype"] == "f32" else 2
    mb = BYTES_PER_PX[(cfg["dtype"], cfg["ksize"])] * cfg["B"] * cfg["H"] * cfg["W"] / 1e6
    return {"workload": f"batch {cfg['B']} x {cfg['W']}x{cfg['H']}, {cfg['ksize']}x{cfg['ksize']}, {cfg['iters']} iterations, fp{8 * es}, forward only, "
                        f"mode {'CSPN_new' if cfg['mode'] == 0 else 'CSPN_ours'}, per GPU (BASELINE.json configs[1])",
            "per_gpu_batch": cfg["B"], "height": cfg["H"], "width": cfg["W"], "iters": cfg["iters"],
            "l2_policy": f"inputs rotate over {nsets} independent sets ({nsets} x {mb:.1f} MB > 126 MB L2)", "sharding": "independent batch slices, no collective",
            "launch": "timed steps replayed from CUDA graphs (one kernel launch per step)"}


def max_abs_vs_oracle(cfg, dev):
    """Largest |ours - oracle| over the whole output of the bench configuration (one input set), and the gradient errors of
    forward + backward relative to the largest entry (BASELINE.json metric: "max-abs vs ref"; target 1e-4)."""
    from oracle import c_oracle
    g, d, s = synth(cfg, 1000)
    step = run_module(cfg, [[t.to(dev).requires_grad_(i < 2) for i, t in enumerate((g, d, s))]])
    y = step(0)
    go = torch.randn(d.shape, generator=torch.Generator().manual_seed(9)).to(d.dtype)
    y.backward(go.to(dev))
    f32 = [t.float().numpy() for t in (g, d, s, go)]
    ref = c_oracle.forward(f32[0], f32[1], f32[2], cfg["iters"], cfg["ksize"], cfg["mode"], threads=0)
    out = {"max_abs": float(np.abs(y.detach().float().cpu().numpy() - ref).max()), "oracle": "oracle/cspn_oracle.c on the same inputs (pinned to reference-run goldens)"}
    if cfg["ksize"] == 3:
        gg, gd = c_oracle.backward(f32[0], f32[1], f32[2], f32[3], cfg["iters"], 3, cfg["mode"], threads=0)
        tg, td = step.inputs[0][0].grad, step.inputs[0][1].grad
        out["grad_depth_rel"] = float(np.abs(td.float().cpu().numpy() - gd).max() / max(1.0, np.abs(gd).max()))
        out["grad_guidance_rel"] = float(np.abs(tg.float().cpu().numpy() - gg).max() / max(1.0, np.abs(gg).max()))
    return out


def time_eager(cfg, dev, steps=200):
    """Per-call time when every step is launched from Python (no CUDA graph): what an eager training loop pays."""
    sets, _ = device_sets(cfg, dev, 0)
    step = run_module(cfg, sets)
    with torch.no_grad():
        for i in range(10):
            step(i)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(i)
        e1.record()
        torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / steps


def time_reference_on_gpu(cfg, dev, steps=5):
    """Context row: the UNMODIFIED reference module (baseline/_ref) run on the B200 itself (`.cuda()` tensors, ~500 ATen
    launches per forward) - the only apples-to-apples GPU comparison.  None when the reference is not staged."""
    fn, kind, what = reference_forward(cfg)
    if kind != "reference":
        return None
    g, d, s = [t.to(dev) for t in synth(dict(cfg, dtype="f32"), 1000)]
    with torch.no_grad():
        for _ in range(2):
            y = fn(g, d, s)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            y = fn(g, d, s)
        e1.record()
        torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    ours = run_module(cfg, [[g, d, s]])(0)
    return {"value": cfg["B"] * cfg["H"] * cfg["W"] / (ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms, "what": what + " on cuda:0, eager, L2-warm inputs",
            "max_abs_vs_ours": float((y - ours).abs().max())}


def extra_rows(dev, rank, world, dist, peak):
    """The other BASELINE sizes on every rank at once (weak scaling: each rank its own batch), aggregated over ranks."""
    extra = {}

    def agg(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    for name, c, st in (("kitti_b32_1216x352_f16_fwd", KITTI, 30), ("pac5x5_b16_640x480_f32_fwd", PAC5, 10)):
        try:
            ms, ln, ns = time_device(c, dev, rank, st, 3, dist)
            px = c["B"] * c["H"] * c["W"]
            gbs = BYTES_PER_PX[(c["dtype"], c["ksize"])] * px / (ms / st * 1e-3) / 1e9
            extra[name] = {"value": world * px / (ms / st * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms / st, "launches_per_step": ln,
                           "roofline_frac": gbs / peak, "input_sets": ns, "n_gpus": world, "per_gpu_batch": c["B"]}
        except Exception as exc:        # context numbers must not take the headline down
            extra[name] = {"error": repr(exc)[:200]}
        torch.cuda.empty_cache()
    # forward + backward (BASELINE.json configs[2] is forward+backward): algorithmic bytes 11 + 20 elements per pixel
    # (5x5: 27 forward + 52 backward elements: 24 + 1 + 1 + 1 read, 24 + 1 written)
    for name, c, st, elems in (("kitti_b32_1216x352_f16_fwd_bwd", KITTI, 10, 31), ("nyu_b8_304x228_f32_fwd_bwd", NYU, 100, 31),
                               ("pac5x5_b16_640x480_f32_fwd_bwd", PAC5, 5, 79)):
        try:
            ms, ln = time_fwd_bwd(c, dev, st)
            ms = agg(ms)
            px = c["B"] * c["H"] * c["W"]
            es = 4 if c["dtype"] == "f32" else 2
            extra[name] = {"value": world * px / (ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms, "launches_per_step": ln,
                           "roofline_frac": elems * es * px / (ms * 1e-3) / 1e9 / peak, "n_gpus": world, "per_gpu_batch": c["B"]}
        except Exception as exc:
            extra[name] = {"error": repr(exc)[:200]}
        torch.cuda.empty_cache()
    if world == 1:
        try:
            extra.update(neighbour_rows(dev))
        except Exception as exc:
            extra["neighbour_rows_error"] = repr(exc)[:200]
        torch.cuda.empty_cache()
    try:
        extra["unet_train_step"] = unet_train_step(dev, rank, world, dist)
    except Exception as exc:
        extra["unet_train_step"] = {"error": repr(exc)[:200]}
    torch.cuda.empty_cache()
    return extra


def _event_ms(fn, reps, dev):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / reps


def neighbour_rows(dev):
    """what sits either side of the module, at the headline batch (8 x 304x228 fp32), eager launches timed with
    CUDA events: the two output heads (upstream), masked-L1 loss and metrics (downstream), the legacy max-of-8 CSPN, and the
    whole tail of unet_cspn_nyu's forward + backward (heads -> CSPN-24 -> loss) with the reference's formulation beside it."""
    import torch.nn.functional as F
    from cspn_monodepth_b200 import criteria, cspn_legacy, cspn_new, heads
    rows = {}
    b, cin, h, w, H, W, ng = NYU["B"], 64, 114, 152, NYU["H"], NYU["W"], 12
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(b, cin, h, w, generator=gen).to(dev).requires_grad_(True)
    wd = (torch.randn(1, cin, 3, 3, generator=gen) / 24).to(dev).requires_grad_(True)
    wg = (torch.randn(ng, cin, 3, 3, generator=gen) / 24).to(dev).requires_grad_(True)
    sparse = ((torch.rand(b, 1, H, W, generator=gen) < 500.0 / 69312.0) * (torch.rand(b, 1, H, W, generator=gen) * 10 + 0.1)).to(dev)
    target = (torch.rand(b, 1, H, W, generator=gen) * 9.5 + 0.5).to(dev)
    px = b * H * W

    def ref_heads():
        k = torch.zeros(cin, 1, 2, 2, device=dev)
        k[:, :, 0, 0] = 1
        u = F.conv_transpose2d(x, k, stride=2, groups=cin)[:, :, :H, :W]      # unet_ours.py:138-150 (the NYU file builds its mask in a Python loop)
        return F.conv2d(u, wd, padding=1), F.conv2d(u, wg, padding=1)

    with torch.no_grad():
        ms = _event_ms(lambda: heads.guidance_depth_heads(x, wd, wg, H, W), 50, dev)
        tf32 = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        ms_ref = _event_ms(ref_heads, 10, dev)
        d_ref, g_ref = ref_heads()
        torch.backends.cudnn.allow_tf32 = tf32
        d, g = heads.guidance_depth_heads(x, wd, wg, H, W)
    rows["heads_nyu_b8_f32_fwd"] = {"value": px / (ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms, "launches_per_step": 1,
                                    "useful_tflops": 2.0 * b * h * w * 9 * cin * (1 + ng) / (ms * 1e-3) / 1e12,
                                    "roofline_frac_hbm": 4.0 * (b * cin * h * w + (1 + ng) * px) / (ms * 1e-3) / 1e9 / hbm_peak()[0],
                                    "torch_cudnn_fp32_ms": ms_ref, "max_abs_vs_torch_fp32": max(float((d - d_ref).abs().max()), float((g - g_ref).abs().max())),
                                    "what": "both heads (64 -> 1 + 12 channels) in one launch of csrc/cspn_heads.cu vs conv_transpose2d + 2 x conv2d (cuDNN, TF32 off)"}
    god, gog = torch.randn_like(d), torch.randn_like(g)

    def heads_train():
        dd, gg = heads.guidance_depth_heads(x, wd, wg, H, W)
        torch.autograd.backward([dd, gg], [god, gog])
    rows["heads_nyu_b8_f32_fwd_bwd"] = {"ms_per_step": _event_ms(heads_train, 20, dev), "launches_per_step": 4}
    # legacy max-of-8 CSPN, 16 steps (CSPN.py)
    with torch.no_grad():
        gd_ = g[:, :8].contiguous()
        ms = _event_ms(lambda: cspn_legacy.legacy_propagate(gd_, d, sparse, 16), 50, dev)
    rows["legacy_cspn16_nyu_b8_f32_fwd"] = {"value": px / (ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms, "launches_per_step": 4,
                                            "roofline_frac": 44.0 * px / (ms * 1e-3) / 1e9 / hbm_peak()[0]}
    # masked L1 loss (fwd + bwd) and the 10 metrics: 8 / 12 / 8 bytes per pixel
    pred = (target + 0.3 * torch.randn_like(target)).abs().add(0.05).requires_grad_(True)
    crit = criteria.MaskedL1Loss()

    def loss_step():
        pred.grad = None
        crit(pred, target).backward()
    ms_l = _event_ms(loss_step, 50, dev)
    with torch.no_grad():
        ms_m = _event_ms(lambda: criteria.evaluate_device(pred, target), 50, dev)
    rows["masked_l1_fwd_bwd_and_metrics_nyu_b8"] = {"loss_fwd_bwd_ms": ms_l, "metrics_ms": ms_m, "launches": "1 + 1, 1",
                                                    "what": "eager Python calls: the kernels stream 4.4 MB, the time is launch + autograd overhead"}
    # in-place activated batch norm (network/libs/inplace_abn) on the decoder's last 64-channel activation, training mode
    from cspn_monodepth_b200 import abn
    bn = abn.InPlaceABN(cin).to(dev)
    bn_ref = torch.nn.Sequential(torch.nn.BatchNorm2d(cin), torch.nn.LeakyReLU(0.01)).to(dev)
    act_in = torch.randn(b, cin, h, w, generator=gen).to(dev).requires_grad_(True)
    gz = torch.randn(b, cin, h, w, generator=gen).to(dev)
    n_el = b * cin * h * w

    def abn_step(mod):
        def run():
            act_in.grad = None
            mod(act_in * 1.0).backward(gz)
        return run
    with torch.no_grad():
        buf = torch.empty_like(act_in)
        ms_f = _event_ms(lambda: bn(buf.copy_(act_in)), 50, dev) - _event_ms(lambda: buf.copy_(act_in), 50, dev)
    ms_fb, ms_fb_ref = _event_ms(abn_step(bn), 30, dev), _event_ms(abn_step(bn_ref), 30, dev)
    rows["inplace_abn_b8x64x114x152_f32"] = {"fwd_ms": ms_f, "fwd_roofline_frac": 12.0 * n_el / (ms_f * 1e-3) / 1e9 / hbm_peak()[0],
                                             "fwd_bwd_ms": ms_fb, "torch_batchnorm_leakyrelu_fwd_bwd_ms": ms_fb_ref, "launches": "3 + 1 forward, 2 + 1 backward",
                                             "what": "InPlaceABN(64) training step on 8x64x114x152 (35.5 MB): forward = 12 B/element (statistics read + "
                                                     "normalise in place), backward = 20 B/element; both timings include the x*1.0 copy that feeds the module"}
    # the tail of unet_cspn_nyu.ResNet.forward (:383-386) + criterion, forward + backward
    prop = cspn_new.AffinityPropagate(24, 3)

    def tail_ours():
        x.grad = wd.grad = wg.grad = None
        dd, gg = heads.guidance_depth_heads(x, wd, wg, H, W)
        crit(prop(gg, dd, sparse), target).backward()
    rows["unet_tail_nyu_b8_f32_train"] = {"ms_per_step": _event_ms(tail_ours, 20, dev),
                                          "what": "heads -> CSPN-24 -> masked L1, forward + backward: 1 + 1 + 1 forward launches, 1 + 1 + 4 backward"}
    fn, kind, _ = reference_forward(NYU)
    if kind == "reference":
        def tail_ref():
            x.grad = wd.grad = wg.grad = None
            dd, gg = ref_heads()
            y = fn(gg, dd, sparse)
            valid = target > 0
            (target - y)[valid].abs().mean().backward()
        try:
            rows["unet_tail_nyu_b8_f32_train"]["reference_formulation_ms"] = _event_ms(tail_ref, 3, dev)
        except Exception as exc:
            rows["unet_tail_nyu_b8_f32_train"]["reference_error"] = repr(exc)[:160]
    return rows


def unet_train_step(dev, rank, world, dist):
    """BASELINE.json configs[4] "NCCL grad all-reduce only": one training step of the reference's own unet_cspn_nyu (staged under
    baseline/_ref, ResNet-50 encoder, 256 M parameters - resnet18 does not build in the reference) with the B200 heads + CSPN module dropped in, batch 8 per GPU; with N > 1 the model is
    wrapped in DistributedDataParallel - the only collective is its gradient all-reduce, the CSPN path itself issues none."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "network")):
        return {"unavailable": "baseline/_ref is not staged"}
    if ref not in sys.path:
        sys.path.insert(0, ref)
    from cspn_monodepth_b200 import criteria, dropin
    if torch.cuda.device_count() > 1:
        dropin.install_inplace_abn()       # unet_cspn_nyu.py:19-25 imports InPlaceABNSync on multi-GPU boxes; the reference's needs a torch-0.4 cffi build
    from network import unet_cspn_nyu
    torch.manual_seed(11)
    model = unet_cspn_nyu.resnet50(pretrained=False).to(dev).train()
    dropin.patch_model(model, heads=True)
    ddp = model
    if dist is not None:
        from torch.nn.parallel import DistributedDataParallel
        ddp = DistributedDataParallel(model, device_ids=[dev.index], find_unused_parameters=True)      # the reference model carries layers its forward never calls
    opt = torch.optim.SGD(ddp.parameters(), lr=1e-3, momentum=0.9)
    crit = criteria.MaskedL1Loss()
    gen = torch.Generator().manual_seed(100 + rank)
    rgb = torch.rand(8, 3, 228, 304, generator=gen)
    dense = torch.rand(8, 1, 228, 304, generator=gen) * 9 + 0.5
    mask = torch.rand(8, 1, 228, 304, generator=gen) < 500.0 / 69312.0
    xin, target = torch.cat([rgb, dense * mask], dim=1).to(dev), dense.to(dev)
    losses = []

    def step():
        opt.zero_grad(set_to_none=True)
        loss = crit(ddp(xin), target)
        loss.backward()
        opt.step()
        losses.append(loss.detach())
    ms = _event_ms(step, 3, dev)
    out = {"ms_per_step": ms, "images_per_s": world * 8 / (ms * 1e-3), "n_gpus": world, "per_gpu_batch": 8,
           "loss_first_last": [float(losses[0]), float(losses[-1])],
           "what": "reference unet_cspn_nyu.resnet50 (its decoder blocks still build their unpooling masks in Python loops) + B200 heads / CSPN-24 / masked-L1 kernels, SGD step" + (", DDP gradient all-reduce over NCCL" if dist is not None else "")}
    if dist is not None:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["ms_per_step"] = float(t.item())
        out["images_per_s"] = world * 8 / (out["ms_per_step"] * 1e-3)
        p0 = next(model.parameters()).detach().flatten()[:1024].clone()
        lo, hi = p0.clone(), p0.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        out["replicas_in_sync"] = bool(torch.equal(lo, hi))
    return out


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        main_reference(args, rank)
        return
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None     # host buffers next to their GPU
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    cfg = dict(NYU, B=int(os.environ.get("CSPN_BENCH_B", NYU["B"])))      # B override: experiments only
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_total, launches, nsets = time_device(cfg, dev, rank, args.steps, args.warmup, dist, sampler, graph=not args.eager)
    ms_step = ms_total / args.steps
    px_step = cfg["B"] * cfg["H"] * cfg["W"]
    value = world * px_step / (ms_step * 1e-3) / 1e6
    # end-to-end leg on every rank at once (each GPU has its own PCIe link); the slowest rank sets the time
    e_ms, e_sync_ms, h2d, d2h, _, depth = time_e2e(cfg, dev, min(args.steps, 50), dist)
    peak, peak_src = hbm_peak()
    extra = None if args.no_extra else extra_rows(dev, rank, world, dist, peak)      # every rank runs them (aggregated over ranks)
    if rank == 0:
        alg_bytes = BYTES_PER_PX[(cfg["dtype"], cfg["ksize"])] * px_step
        achieved = alg_bytes / (ms_step * 1e-3) / 1e9
        from cspn_monodepth_b200 import _lib
        plan = _lib.forward_plan(cfg["B"], 1, cfg["H"], cfg["W"], cfg["iters"], cfg["ksize"], cfg["mode"])
        kernel = {1: "cspn::fused3x3_kernel<float,10,8,CSPN_new> (single 64x80 tile per CTA)", 2: "cspn::dual3x3_kernel<float,10,CSPN_new> (two 64x40 tiles per CTA)"}.get(plan["kernel"], "?")
        if plan.get("transport"):
            kernel += {"cluster": ", hardware clusters (DSMEM halo ring)", "stream": ", stream transport (halo ring through global inboxes)",
                       "hybrid": ", hybrid transport (row clusters: DSMEM left / right, global inboxes up / down)"}[plan["transport"]]
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(cfg, nsets), "gpu_launches": launches * args.steps,
                "clocks": sampler.summary(),
                "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": ncu_traffic("nyu_b8_f32"), "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": alg_bytes, "launch_us": ms_step * 1e3,
                             "fma_floor_us": px_step * cfg["iters"] * 8 / (148 * 128 * 1.965e9) * 1e6}}
        line["e2e"] = {"value": world * px_step / (e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": e_ms,
                       "api": f"cspn_fwd_host_submit_f32 + cspn_host_wait (C ABI, host buffers in one pinned staging arena), {depth} calls in flight per rank: "
                              "the H2D copy of a call overlaps the kernel and the D2H copy of the previous one; every result is waited for in host memory",
                       "timing": "host wall clock around blocks of calls that start and end with an idle device (median of 3 blocks, max over ranks)",
                       "sync_call": {"value": world * px_step / (e_sync_ms * 1e-3) / 1e6, "ms_per_step": e_sync_ms, "api": "cspn_fwd_host_f32, one blocking call per step"},
                       "n_gpus": world, "host_binding": numa}
        try:
            line["max_abs"] = max_abs_vs_oracle(cfg, dev)
        except Exception as exc:
            line["max_abs"] = {"error": repr(exc)[:200]}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"], _ = cpu_reference(cfg)
            line["cpu_baseline_fused_c"] = cpu_c_oracle(cfg)
            # BASELINE.json configs[0]: the reference's own CPU-runnable case (B = 1), one thread and all cores
            cores = os.cpu_count() or 1
            line["cpu_cfg1"] = {f"{t}_threads": cpu_reference(cfg, min_seconds=1.0, max_steps=10, min_steps=5, batch=1, threads=t)[0] for t in sorted({1, cores})}
        if extra is not None:
            if world == 1:
                try:
                    extra["eager_launch_us_per_call"] = {"value": time_eager(cfg, dev) * 1e3, "what": "nn.Module forward launched from Python every step (no CUDA graph), same workload"}
                    extra["reference_module_on_b200"] = time_reference_on_gpu(cfg, dev)
                except Exception as exc:
                    extra["context_error"] = repr(exc)[:200]
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
