"""The legacy max-of-8 CSPN kernel (csrc/cspn_legacy.cu through cspn_monodepth_b200/cspn_legacy.py) against the reference-run
vectors and the numpy oracle.  Tolerance: 1e-4 at depth scale 10 (scaled with the value range otherwise); NaN patterns exact."""
import os

import numpy as np
import pytest
import torch

from cspn_monodepth_b200 import _lib, cspn_legacy
from oracle import legacy_oracle
from tests.util import assert_close_nan, make_inputs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"


def _cu(a, dtype=torch.float32):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(DEV).to(dtype)


def _atol(ref):
    return 1e-4 * max(1.0, float(np.nanmax(np.abs(ref))) / 10.0) if np.isfinite(ref).any() else 1e-4


def test_matches_reference_vectors():
    z = np.load(os.path.join(ROOT, "tests", "golden", "legacy_golden.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    assert len(names) == 7
    for n in names:
        g, d, s = z[n + "/guidance"], z[n + "/depth"], z[n + "/sparse"]
        with torch.no_grad():
            y = cspn_legacy.AffinityPropagate()(_cu(g), _cu(d), _cu(s)).cpu().numpy()
            y2 = cspn_legacy.AffinityPropagate_prediction()(_cu(g), _cu(d)).cpu().numpy()
        assert_close_nan(y, z[n + "/out"], _atol(z[n + "/out"]), n)
        assert_close_nan(y2, z[n + "/out_prediction"], _atol(z[n + "/out_prediction"]), n + " prediction")


@pytest.mark.parametrize("b,cg,h,w,dtype", [(8, 8, 228, 304, torch.float32), (2, 12, 352, 1216, torch.float32), (2, 8, 352, 1216, torch.float16),
                                            (3, 8, 97, 131, torch.float32), (1, 8, 3, 1000, torch.float32), (1, 8, 500, 5, torch.float16)])
def test_full_size_vs_oracle(b, cg, h, w, dtype):
    g, d, s = make_inputs(11, b, cg, 1, h, w, density=0.02)
    tg, td, ts = _cu(g, dtype), _cu(d, dtype), _cu(s, dtype)
    ref = legacy_oracle.forward(tg.float().cpu().numpy(), td.float().cpu().numpy(), ts.float().cpu().numpy())
    with torch.no_grad():
        y = cspn_legacy.AffinityPropagate()(tg, td, ts)
    lib = _lib.load()
    assert lib.cspn_last_launch_count() == 4
    tol = _atol(ref) + (float(np.abs(ref).max()) * 2.0 ** -10 if dtype == torch.float16 else 0.0)    # + 1 fp16 ulp of the output
    assert_close_nan(y.float().cpu().numpy(), ref, tol, f"{b}x{h}x{w}")
    hit = s > 0
    assert np.array_equal(y.float().cpu().numpy()[hit], ts.float().cpu().numpy()[hit])                 # samples re-injected exactly


@pytest.mark.parametrize("iters", [1, 3, 4, 5, 8, 9, 13])
def test_iteration_counts_and_strided_guidance(iters):
    g, d, s = make_inputs(5, 2, 12, 1, 45, 83, density=0.05)
    wide = _cu(g)
    ref = legacy_oracle.forward(g[:, 2:10], d, s, iters)
    with torch.no_grad():
        y = cspn_legacy.legacy_propagate(wide[:, 2:10], _cu(d), _cu(s), iters)                         # channel-narrowed view: batch stride 12 planes
    assert_close_nan(y.cpu().numpy(), ref, _atol(ref), f"T={iters}")


def test_errors_and_graph_capture():
    g, d, s = make_inputs(1, 1, 8, 1, 40, 64)
    tg, td, ts = _cu(g), _cu(d), _cu(s)
    with pytest.raises(RuntimeError):
        cspn_legacy.AffinityPropagate()(tg.cpu(), td.cpu(), ts.cpu())
    with pytest.raises(RuntimeError):
        cspn_legacy.AffinityPropagate()(tg[:, :7], td, ts)
    with pytest.raises(RuntimeError):
        cspn_legacy.AffinityPropagate()(tg.requires_grad_(True), td, ts)
    tg = tg.detach()
    ref = legacy_oracle.forward(g, d, s)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st), torch.no_grad():
        cspn_legacy.AffinityPropagate()(tg, td, ts)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            y = cspn_legacy.AffinityPropagate()(tg, td, ts)
        gr.replay()
        st.synchronize()
    assert_close_nan(y.cpu().numpy(), ref, _atol(ref), "graph replay")
