"""In-place activated batch normalisation kernels (csrc/cspn_abn.cu, cspn_monodepth_b200/abn.py) against the BatchNorm-autograd
vectors (tests/golden/abn_golden.npz) and the float64 oracle (oracle/abn_oracle.py).

Tolerances (fp32 kernels, statistics accumulated in double): outputs 4 ulp of the value + 1e-6; input gradients 2e-5 relative
to the largest gradient (y is recovered from the saved OUTPUT by a division, as in the reference); parameter gradients 1e-5
relative; running statistics 1e-6."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cspn_monodepth_b200 import abn
from oracle import abn_oracle
from tests.test_abn_oracle import ACTS, case_args, load_abn_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cu(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(DEV)


def _module(c, w, b, rm, rv, act, slope, sync=False):
    m = (abn.InPlaceABNSync if sync else abn.InPlaceABN)(c, affine=w is not None, activation=act, slope=slope).to(DEV)
    with torch.no_grad():
        if w is not None:
            m.weight.copy_(_cu(w)), m.bias.copy_(_cu(b))
        m.running_mean.copy_(_cu(rm)), m.running_var.copy_(_cu(rv))
    return m


def _run(m, x, dz):
    xin = _cu(x).requires_grad_(True)
    h = xin * 1.0                                   # a non-leaf tensor the module may overwrite, like a conv output
    z = m(h)
    assert z.data_ptr() == h.data_ptr()             # in place
    z.backward(_cu(dz))
    return z.detach().cpu().numpy(), xin.grad.cpu().numpy()


def _close(a, ref, rel, what):
    ref = np.asarray(ref, np.float64)
    err = np.abs(np.asarray(a, np.float64) - ref).max()
    assert err <= rel * max(1.0, np.abs(ref).max()), (what, err)


def test_matches_batchnorm_autograd_vectors():
    for name, c in load_abn_golden().items():
        training, act, slope, w, b = case_args(c)
        m = _module(c["x"].shape[1], w, b, c["rm0"], c["rv0"], act, slope).train(training)
        z, dx = _run(m, c["x"], c["dz"])
        _close(z, c["z"], 1e-6, name + " z")
        _close(dx, c["dx"], 2e-5, name + " dx")
        _close(m.running_mean.cpu().numpy(), c["running_mean"], 1e-6, name + " running_mean")
        _close(m.running_var.cpu().numpy(), c["running_var"], 1e-6, name + " running_var")
        if w is not None and training:
            _close(m.weight.grad.cpu().numpy(), c["dweight"], 1e-5, name + " dweight")
            _close(m.bias.grad.cpu().numpy(), c["dbias"], 1e-5, name + " dbias")
        if w is not None and not training:
            assert not m.weight.grad.any() and not m.bias.grad.any()          # functions.py:147-150


@pytest.mark.parametrize("shape,act", [((8, 64, 114, 152), "leaky_relu"), ((8, 64, 57, 76), "elu"), ((4, 256, 29, 38), "none"),
                                        ((2, 3, 5, 1031), "leaky_relu"), ((1, 1, 1, 2), "leaky_relu"), ((32, 64, 88, 304), "leaky_relu")])
def test_full_size_vs_oracle(shape, act):
    rng = np.random.default_rng(5)
    c = shape[1]
    x = (rng.standard_normal(shape) * 1.3 + rng.standard_normal((1, c, 1, 1))).astype(np.float32)
    dz = rng.standard_normal(shape).astype(np.float32)
    w, b = (rng.standard_normal(c) + 0.3).astype(np.float32), rng.standard_normal(c).astype(np.float32)
    rm, rv = np.zeros(c, np.float32), np.ones(c, np.float32)
    m = _module(c, w, b, rm, rv, act, 0.01)
    z, dx = _run(m, x, dz)
    zr, mean, var, rmr, rvr = abn_oracle.abn_forward(x, w, b, rm, rv, True, 0.1, 1e-5, act, 0.01)
    _close(z, zr, 2e-6, "z")
    dxr, dwr, dbr = abn_oracle.abn_backward(z, dz, var, w, b, True, 1e-5, act, 0.01)      # from the kernel's own fp32 output, like the kernel
    _close(dx, dxr, 2e-5, "dx")
    _close(m.weight.grad.cpu().numpy(), dwr, 1e-5, "dweight")
    _close(m.bias.grad.cpu().numpy(), dbr, 1e-5, "dbias")
    _close(m.running_mean.cpu().numpy(), rmr, 1e-6, "running_mean")
    _close(m.running_var.cpu().numpy(), rvr, 1e-6, "running_var")
    z2, dx2 = _run(_module(c, w, b, rm, rv, act, 0.01), x, dz)
    assert np.array_equal(z, z2) and np.array_equal(dx, dx2)                  # deterministic reductions: bitwise reproducible


def test_errors_like_the_reference():
    m = abn.InPlaceABN(4).to(DEV)
    x = torch.randn(2, 4, 6, 6, device=DEV)
    with pytest.raises(ValueError, match="Non-contiguous"):                     # functions.py:65-67
        m(x.transpose(2, 3))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        m(torch.randn(2, 4, 3, 3))
    with pytest.raises(RuntimeError, match="float32"):
        m(x.half())
    assert "activation=leaky_relu" in repr(m) and "slope=0.01" in repr(m)
    assert [k for k, _ in m.state_dict().items()] == ["weight", "bias", "running_mean", "running_var"]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _sync_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)        # both replicas share cuda:0 here; NCCL needs one GPU per rank
    try:
        rng = np.random.default_rng(11)
        c = 6
        xs = [(rng.standard_normal((2, c, 9, 13)) * (1 + k) + k).astype(np.float32) for k in range(world)]
        dzs = [rng.standard_normal((2, c, 9, 13)).astype(np.float32) for _ in range(world)]
        w, b = (rng.standard_normal(c) + 0.3).astype(np.float32), rng.standard_normal(c).astype(np.float32)
        m = _module(c, w, b, np.zeros(c), np.ones(c), "leaky_relu", 0.01, sync=True)
        z, dx = _run(m, xs[rank], dzs[rank])
        zr, mean, var, rmr, rvr = abn_oracle.abn_forward(xs[rank], w, b, np.zeros(c), np.ones(c), True, 0.1, 1e-5, "leaky_relu", 0.01, world_x=xs)
        zall = [abn_oracle.abn_forward(xs[k], w, b, np.zeros(c), np.ones(c), True, 0.1, 1e-5, "leaky_relu", 0.01, world_x=xs)[0] for k in range(world)]
        dxr, dwr, dbr = abn_oracle.abn_backward(zall[rank], dzs[rank], var, w, b, True, 1e-5, "leaky_relu", 0.01, world=list(zip(zall, dzs)))
        _close(z, zr, 2e-6, "z")
        _close(dx, dxr, 2e-5, "dx")
        _close(m.weight.grad.cpu().numpy(), dwr, 1e-5, "dweight")
        _close(m.bias.grad.cpu().numpy(), dbr, 1e-5, "dbias")
        _close(m.running_mean.cpu().numpy(), rmr, 1e-6, "rm")
        _close(m.running_var.cpu().numpy(), rvr, 1e-6, "rv")
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_sync_variant_two_replicas():
    """InPlaceABNSync over two replicas (functions.py:166-297): statistics and gradient means over both, one all-reduce each way."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sync_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)


def test_sync_variant_over_nccl():
    """The same check with one process per GPU and NCCL (needs two GPUs; the two-replica gloo test above covers single-GPU boxes)."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), os.path.join(root, "tests", "abn_nccl_worker.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok abn nccl 2" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
