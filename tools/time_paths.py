"""A/B timing of the CUDA paths through the raw C ABI (experiments; bench.py is the judged harness).

    python tools/time_paths.py [fwd|bwd|all] [--cfg nyu|kitti|kitti4|pac5] [--path auto|generic]

Every measurement captures `reps` calls into one CUDA graph (inputs rotate over enough sets to defeat L2) and
times the replay with CUDA events.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cspn_monodepth_b200 import _lib  # noqa: E402

CFGS = {
    "nyu": dict(B=8, H=228, W=304, iters=24, ksize=3, mode=0, dtype=torch.float32),
    "nyu16": dict(B=8, H=228, W=304, iters=24, ksize=3, mode=0, dtype=torch.float16),
    "kitti": dict(B=32, H=352, W=1216, iters=24, ksize=3, mode=0, dtype=torch.float16),
    "kitti32": dict(B=32, H=352, W=1216, iters=24, ksize=3, mode=0, dtype=torch.float32),
    "kitti1": dict(B=1, H=352, W=1216, iters=24, ksize=3, mode=0, dtype=torch.float16),
    "nyu7": dict(B=7, H=228, W=304, iters=24, ksize=3, mode=0, dtype=torch.float32),
    "nyu9": dict(B=9, H=228, W=304, iters=24, ksize=3, mode=0, dtype=torch.float32),
    "kitti4": dict(B=4, H=352, W=1216, iters=24, ksize=3, mode=0, dtype=torch.float16),
    "pac3": dict(B=8, H=228, W=304, iters=24, ksize=3, mode=1, dtype=torch.float32),
    "pac5": dict(B=16, H=480, W=640, iters=12, ksize=5, mode=1, dtype=torch.float32),
}


def make(cfg, seed, dev):
    g = torch.Generator().manual_seed(seed)
    b, h, w = cfg["B"], cfg["H"], cfg["W"]
    cg = cfg["ksize"] ** 2 - 1
    guidance = torch.randn(b, cg, h, w, generator=g)
    depth = torch.rand(b, 1, h, w, generator=g) * 10
    sparse = (torch.rand(b, 1, h, w, generator=g) < 0.0072) * (torch.rand(b, 1, h, w, generator=g) * 10 + 0.1)
    gout = torch.randn(b, 1, h, w, generator=g)
    return [t.to(cfg["dtype"]).to(dev).contiguous() for t in (guidance, depth, sparse, gout)]


def timed(fn, reps, dev):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize(dev)
    side = torch.cuda.Stream(dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for i in range(reps):
                fn(i)
    graph.replay()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="all")
    ap.add_argument("--cfg", default="nyu,kitti")
    ap.add_argument("--path", default="auto")
    ap.add_argument("--reps", type=int, default=0)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = _lib.load()
    lib.cspn_set_path({"auto": 0, "generic": 1, "fused": 2}[args.path])
    for name in args.cfg.split(","):
        cfg = CFGS[name]
        b, h, w, it, k, mode = cfg["B"], cfg["H"], cfg["W"], cfg["iters"], cfg["ksize"], cfg["mode"]
        cg = k * k - 1
        es = 4 if cfg["dtype"] == torch.float32 else 2
        px = b * h * w
        nsets = max(2, int(1.5 * 126e6 / ((cg + 3) * es * px)) + 1)
        sets = [make(cfg, i, dev) for i in range(nsets)]
        sfx = "f32" if es == 4 else "f16"
        out = torch.empty_like(sets[0][1])
        gg = torch.empty_like(sets[0][0])
        gd = torch.empty_like(sets[0][1])
        nf = lib.cspn_fwd_workspace_bytes(b, 1, h, w, it, k, mode)
        nb = lib.cspn_bwd_workspace_bytes(b, 1, h, w, it, k, mode)
        wsf = torch.empty(max(nf, 16), dtype=torch.uint8, device=dev)
        wsb = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
        fwd_fn, bwd_fn = getattr(lib, "cspn_fwd_" + sfx), getattr(lib, "cspn_bwd_" + sfx)

        def fwd(i):
            g, d, s, _ = sets[i % nsets]
            _lib.check(fwd_fn(g.data_ptr(), cg * h * w, d.data_ptr(), s.data_ptr(), 1, out.data_ptr(), b, 1, h, w, it, k, mode,
                              wsf.data_ptr(), nf, torch.cuda.current_stream().cuda_stream))

        def bwd(i):
            g, d, s, go = sets[i % nsets]
            _lib.check(bwd_fn(go.data_ptr(), g.data_ptr(), cg * h * w, cg, d.data_ptr(), s.data_ptr(), 1, gg.data_ptr(), gd.data_ptr(),
                              b, 1, h, w, it, k, mode, wsb.data_ptr(), nb, torch.cuda.current_stream().cuda_stream))

        reps = args.reps or (200 if px < 2e6 else 20)
        if args.what in ("fwd", "all"):
            ms = timed(fwd, reps, dev)
            print(f"{name:8s} fwd path={args.path:7s} {ms * 1e3:9.1f} us  {px / ms / 1e3:9.0f} Mpx/s  launches {lib.cspn_last_launch_count()}  "
                  f"alg {(cg + 3) * es * px / ms / 1e6:7.0f} GB/s", flush=True)
        if args.what in ("bwd", "all"):
            ms = timed(bwd, max(2, reps // 2), dev)
            print(f"{name:8s} bwd path={args.path:7s} {ms * 1e3:9.1f} us  {px / ms / 1e3:9.0f} Mpx/s  launches {lib.cspn_last_launch_count()}  "
                  f"alg {(2 * cg + 4) * es * px / ms / 1e6:7.0f} GB/s  ws {nb / 1e6:.0f} MB", flush=True)


if __name__ == "__main__":
    main()
