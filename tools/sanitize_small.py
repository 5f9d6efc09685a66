"""Small forward / backward / 5x5 problems for compute-sanitizer (memcheck, racecheck, synccheck):
    CSPN_EXCHANGE=global compute-sanitizer --tool racecheck python tools/sanitize_small.py
Several tiles per image so that both halo transports, the history scratch and the shared-memory tiles are exercised."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cspn_monodepth_b200 import _lib, cspn_new, cspn_ours
from oracle import c_oracle
from tests.util import make_inputs

dev = "cuda:0"
def cu(a): return torch.from_numpy(np.ascontiguousarray(a)).to(dev)
for mode, shape, iters, cg in ((0, (2, 97, 131), 6, 12), (1, (1, 150, 70), 5, 8)):
    b, h, w = shape
    g, d, s = make_inputs(h, b, cg, 1, h, w, density=0.05)
    tg, td, ts = cu(g).requires_grad_(True), cu(d).requires_grad_(True), cu(s)
    mod = cspn_new.AffinityPropagate(iters, 3) if mode == 0 else cspn_ours.AffinityPropagate(iters)
    y = mod(tg, td, ts) if mode == 0 else mod(td, tg, sparse_depth=ts)
    go = np.random.default_rng(1).standard_normal(d.shape).astype(np.float32)
    y.backward(cu(go))
    torch.cuda.synchronize()
    ref = c_oracle.forward(g, d, s, iters, 3, mode)
    gg, gd = c_oracle.backward(g, d, s, go, iters, 3, mode)
    print("mode", mode, shape, "fwd err %.2e" % np.abs(y.detach().cpu().numpy() - ref).max(),
          "gd err %.2e" % np.abs(td.grad.cpu().numpy() - gd).max(), "gg err %.2e" % np.abs(tg.grad.cpu().numpy() - gg).max(), flush=True)
g, d, s = make_inputs(3, 1, 24, 1, 40, 70, density=0.05)
y = cspn_ours.AffinityPropagate(6)(cu(d), cu(g), sparse_depth=cu(s))
torch.cuda.synchronize()
print("5x5 err %.2e" % np.abs(y.cpu().numpy() - c_oracle.forward(g, d, s, 6, 5, 1)).max(), flush=True)

# the neighbours of the module: heads forward + backward, masked L1 + metrics, in-place ABN forward + backward, legacy CSPN
from cspn_monodepth_b200 import abn, criteria, cspn_legacy, heads
rng = np.random.default_rng(9)
x = cu(rng.standard_normal((2, 20, 19, 27)).astype(np.float32)).requires_grad_(True)
wd = cu((rng.standard_normal((1, 20, 3, 3)) / 8).astype(np.float32)).requires_grad_(True)
wg = cu((rng.standard_normal((12, 20, 3, 3)) / 8).astype(np.float32)).requires_grad_(True)
dd, gg_ = heads.guidance_depth_heads(x, wd, wg, 37, 53)
sp = cu(((rng.random((2, 1, 37, 53)) < 0.1) * 3.0).astype(np.float32))
tgt = cu((rng.random((2, 1, 37, 53)) * 9 + 0.5).astype(np.float32))
out = cspn_new.AffinityPropagate(4, 3)(gg_, dd, sp)
loss = criteria.MaskedL1Loss()(out, tgt)
loss.backward()
res = criteria.Result(); res.evaluate(out.detach().abs() + 0.1, tgt)
m = abn.InPlaceABN(20).to(dev)
z = m(x * 1.0)
z.backward(torch.ones_like(z))
with torch.no_grad():
    leg = cspn_legacy.legacy_propagate(gg_[:, :8].contiguous(), dd, sp, 5)
torch.cuda.synchronize()
print("neighbours: loss %.4f rmse %.4f abn mean %.3e legacy mean %.3f gx %.3e" % (float(loss.detach()), res.rmse, float(z.mean()), float(leg.mean()), float(x.grad.abs().mean())), flush=True)
