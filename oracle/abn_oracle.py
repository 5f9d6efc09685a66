"""CPU restatement (numpy, float64) of the reference's in-place activated batch normalisation.  TEST INFRASTRUCTURE ONLY: imported
by tests/ and bench.py's checks - never by the product.

  abn_forward(x, weight, bias, running_mean, running_var, training, momentum, eps, activation, slope)
      network/libs/inplace_abn/functions.py:70-113 + src/bn.cu:125-166 (mean_var_kernel, forward_kernel) + :299-335 (activations)
  abn_backward(z, dz, var, weight, bias, training, eps, activation, slope, world_sums=None)
      functions.py:115-163 + src/bn.cu:168-232 (edz_eydz_kernel, backward_kernel) + :312-377 (activation undo)
  sync variants: pass the per-replica tensors as a list to abn_forward_sync / abn_backward_sync (functions.py:166-297).

Pinning: the reference's native extension (cffi over THC, torch 0.4) cannot be built in this image, so InPlaceABN itself cannot be
run here.  The restatement is pinned to tests/golden/abn_golden.npz instead, produced by running the class the reference itself
substitutes for it on one GPU - nn.BatchNorm2d + activation with autograd (unet_cspn_nyu.py:25-29, bn.py:23-44 `ABN`) - with the
BatchNorm weight set to |w| + eps, which is the only place where InPlaceABN's arithmetic departs from it (bn.cu:153).
"""
import numpy as np


def _bshape(x):
    return (1, -1) + (1,) * (x.ndim - 2)


def _act(z, activation, slope):
    if activation == "leaky_relu":
        return np.where(z < 0, z * slope, z)
    if activation == "elu":
        return np.where(z < 0, np.expm1(np.minimum(z, 0)), z)
    return z


def _act_undo(z, dz, activation, slope):
    """functions.py:55-63: gradient routed through the activation, then the activation inverted on the saved output."""
    if activation == "leaky_relu":
        return np.where(z < 0, z / slope, z), np.where(z < 0, dz * slope, dz)
    if activation == "elu":
        return np.where(z < 0, np.log1p(np.minimum(z, 0)), z), np.where(z < 0, dz * (z + 1), dz)
    return z, dz


def _gamma_beta(weight, bias, eps, c):
    gamma = np.abs(np.asarray(weight, np.float64)) + eps if weight is not None else np.ones(c)        # bn.cu:153
    beta = np.asarray(bias, np.float64) if bias is not None else np.zeros(c)
    return gamma, beta


def batch_sums(x):
    x = np.asarray(x, np.float64)
    axes = (0,) + tuple(range(2, x.ndim))
    return x.sum(axes), (x * x).sum(axes)


def abn_forward(x, weight, bias, running_mean, running_var, training=True, momentum=0.1, eps=1e-5, activation="leaky_relu", slope=0.01, world_x=None):
    """Returns z, mean, var, new_running_mean, new_running_var.  world_x: every replica's x (sync variant), else [x]."""
    x = np.asarray(x, np.float64)
    c = x.shape[1]
    if training:
        parts = [np.asarray(p, np.float64) for p in (world_x if world_x is not None else [x])]
        n = sum(p.size // c for p in parts)
        s1 = sum(batch_sums(p)[0] for p in parts)
        s2 = sum(batch_sums(p)[1] for p in parts)
        mean = s1 / n
        var = np.maximum(s2 / n - mean * mean, 0.0)
        rm = (1 - momentum) * np.asarray(running_mean, np.float64) + momentum * mean              # functions.py:90-92
        rv = (1 - momentum) * np.asarray(running_var, np.float64) + momentum * var * n / (n - 1)
    else:
        mean, var = np.asarray(running_mean, np.float64), np.asarray(running_var, np.float64)
        rm, rv = mean, var
    gamma, beta = _gamma_beta(weight, bias, eps, c)
    inv = np.where((var != 0) | (eps != 0), 1.0 / np.sqrt(var + eps), 0.0)                         # bn.cu:148-151
    sh = _bshape(x)
    z = (x - mean.reshape(sh)) * inv.reshape(sh) * gamma.reshape(sh) + beta.reshape(sh)
    return _act(z, activation, slope), mean, var, rm, rv


def abn_backward(z, dz, var, weight, bias, training=True, eps=1e-5, activation="leaky_relu", slope=0.01, world=None):
    """Returns dx, dweight, dbias.  world: list of (z, dz) of every replica for the sync variant (this replica included)."""
    z, dz = np.asarray(z, np.float64), np.asarray(dz, np.float64)
    c = z.shape[1]
    gamma, beta = _gamma_beta(weight, bias, eps, c)
    sh = _bshape(z)

    def sums(zz, dd):
        zz, dd = _act_undo(np.asarray(zz, np.float64), np.asarray(dd, np.float64), activation, slope)
        y = (zz - beta.reshape(sh)) / gamma.reshape(sh)
        axes = (0,) + tuple(range(2, zz.ndim))
        return dd.sum(axes), (y * dd).sum(axes), zz.size // c

    zu, du = _act_undo(z, dz, activation, slope)
    y = (zu - beta.reshape(sh)) / gamma.reshape(sh)
    n_local = z.size // c
    if training:
        parts = [sums(a, b) for a, b in (world if world is not None else [(z, dz)])]
        n = sum(p[2] for p in parts)
        edz, eydz = sum(p[0] for p in parts) / n, sum(p[1] for p in parts) / n                    # bn.cu:168-186
    else:
        edz, eydz = np.zeros(c), np.zeros(c)                                                      # functions.py:147-150
    var = np.asarray(var, np.float64)
    inv = np.where((var != 0) | (eps != 0), 1.0 / np.sqrt(var + eps), 0.0)
    dx = (du - edz.reshape(sh) - y * eydz.reshape(sh)) * (gamma * inv).reshape(sh)                # bn.cu:204-211
    dweight = dbias = None
    if weight is not None:
        dweight = np.sign(np.asarray(weight, np.float64)) * eydz * n_local                        # bn.cu:217-224
        dbias = edz * n_local                                                                     # :227-231
    return dx, dweight, dbias
