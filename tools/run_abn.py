"""A few training steps of InPlaceABN(64) on the decoder-sized activation (for ncu captures): python tools/run_abn.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cspn_monodepth_b200 import abn  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
m = abn.InPlaceABN(64).to(dev)
x = torch.randn(8, 64, 114, 152, device=dev, requires_grad=True)
gz = torch.randn(8, 64, 114, 152, device=dev)
for _ in range(reps):
    x.grad = None
    m(x * 1.0).backward(gz)
torch.cuda.synchronize()
print("abn", float(x.grad.abs().mean()))
if len(sys.argv) > 2:      # timing: python tools/run_abn.py 3 time
    import bench
    with torch.no_grad():
        buf = torch.empty_like(x)
        f = bench._event_ms(lambda: m(buf.copy_(x)), 200, dev) - bench._event_ms(lambda: buf.copy_(x), 200, dev)

    def step():
        x.grad = None
        m(x * 1.0).backward(gz)
    print("forward %.1f us (12 B/element -> %.0f GB/s), forward+backward %.1f us" % (f * 1e3, 12.0 * x.numel() / f / 1e6, bench._event_ms(step, 100, dev) * 1e3))
