"""InPlaceABNSync over NCCL, one process per GPU (launched by torchrun with 2+ ranks on 2+ GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 tests/abn_nccl_worker.py

Every rank normalises its own slice; forward output, running statistics, input and parameter gradients must equal the float64
oracle evaluated on the concatenation of all slices (functions.py:166-297: statistics and gradient means over every replica)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cspn_monodepth_b200 import abn  # noqa: E402
from oracle import abn_oracle  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
rng = np.random.default_rng(11)
c = 16
xs = [(rng.standard_normal((3, c, 19, 23)) * (1 + k) + k).astype(np.float32) for k in range(world)]
dzs = [rng.standard_normal((3, c, 19, 23)).astype(np.float32) for _ in range(world)]
w, b = (rng.standard_normal(c) + 0.3).astype(np.float32), rng.standard_normal(c).astype(np.float32)
m = abn.InPlaceABNSync(c, activation="leaky_relu", slope=0.01).to(dev)
with torch.no_grad():
    m.weight.copy_(torch.from_numpy(w)); m.bias.copy_(torch.from_numpy(b))
x = torch.from_numpy(xs[rank]).to(dev).requires_grad_(True)
z = m(x * 1.0)
z.backward(torch.from_numpy(dzs[rank]).to(dev))
zr, mean, var, rmr, rvr = abn_oracle.abn_forward(xs[rank], w, b, np.zeros(c), np.ones(c), True, 0.1, 1e-5, "leaky_relu", 0.01, world_x=xs)
zall = [abn_oracle.abn_forward(xs[k], w, b, np.zeros(c), np.ones(c), True, 0.1, 1e-5, "leaky_relu", 0.01, world_x=xs)[0] for k in range(world)]
dxr, dwr, dbr = abn_oracle.abn_backward(zall[rank], dzs[rank], var, w, b, True, 1e-5, "leaky_relu", 0.01, world=list(zip(zall, dzs)))


def close(a, ref, rel, what):
    err = np.abs(np.asarray(a, np.float64) - ref).max()
    assert err <= rel * max(1.0, np.abs(ref).max()), (what, err)


close(z.detach().cpu().numpy(), zr, 2e-6, "z")
close(x.grad.cpu().numpy(), dxr, 2e-5, "dx")
close(m.weight.grad.cpu().numpy(), dwr, 1e-5, "dweight")
close(m.bias.grad.cpu().numpy(), dbr, 1e-5, "dbias")
close(m.running_mean.cpu().numpy(), rmr, 1e-6, "running_mean")
close(m.running_var.cpu().numpy(), rvr, 1e-6, "running_var")
dist.barrier()
if rank == 0:
    print("ok abn nccl", world)
dist.destroy_process_group()
