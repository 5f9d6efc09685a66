"""Forward (and optionally backward) time over batch sizes, to check the planner's choice of halo transport:
    for m in auto global dsmem; do CSPN_EXCHANGE=$m python tools/plan_sweep.py [bwd]; done     (auto = variable unset)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cspn_monodepth_b200 import _lib
import tools.time_paths as tp

what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
dev = torch.device("cuda:0")
lib = _lib.load()
mode = os.environ.get("CSPN_EXCHANGE", "auto")
rows = []
for name, h, w, dt, batches in (("nyu", 228, 304, torch.float32, (1, 2, 4, 7, 8, 12, 16, 32, 64)), ("kitti", 352, 1216, torch.float16, (1, 2, 4, 8, 16, 32))):
    for b in batches:
        cfg = dict(B=b, H=h, W=w, iters=24, ksize=3, mode=0, dtype=dt)
        sets = [tp.make(cfg, i, dev) for i in range(3)]
        sfx = "f32" if dt == torch.float32 else "f16"
        out, gg, gd = torch.empty_like(sets[0][1]), torch.empty_like(sets[0][0]), torch.empty_like(sets[0][1])
        nf, nb = lib.cspn_fwd_workspace_bytes(b, 1, h, w, 24, 3, 0), lib.cspn_bwd_workspace_bytes(b, 1, h, w, 24, 3, 0)
        wsf = torch.empty(max(nf, 16), dtype=torch.uint8, device=dev)
        wsb = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
        ff, bf = getattr(lib, "cspn_fwd_" + sfx), getattr(lib, "cspn_bwd_" + sfx)

        def fwd(i):
            g, d, s, _ = sets[i % 3]
            _lib.check(ff(g.data_ptr(), 8 * h * w, d.data_ptr(), s.data_ptr(), 1, out.data_ptr(), b, 1, h, w, 24, 3, 0, wsf.data_ptr(), nf, torch.cuda.current_stream().cuda_stream))

        def bwd(i):
            g, d, s, go = sets[i % 3]
            _lib.check(bf(go.data_ptr(), g.data_ptr(), 8 * h * w, 8, d.data_ptr(), s.data_ptr(), 1, gg.data_ptr(), gd.data_ptr(), b, 1, h, w, 24, 3, 0,
                          wsb.data_ptr(), nb, torch.cuda.current_stream().cuda_stream))
        ms = tp.timed(fwd if what == "fwd" else bwd, 20 if b * h * w < 4e6 else 6, dev)
        rows.append(f"{name} B={b:3d} {what} {mode:6s} {ms * 1e3:9.1f} us  {b * h * w / ms / 1e3:8.0f} Mpx/s  inbox {'yes' if (nf if what == 'fwd' else nb - 192 * 24 * 64 * 64 * 4) > 0 else 'no '}")
        print(rows[-1], flush=True)
