"""Helpers shared by the tests: load committed golden fixtures and shipped-checkpoint weights."""
import glob
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "case_*.npz")))


def load_npz(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {}
    for k in z.files:
        v = z[k]
        if v.dtype.kind in "fiub" and v.ndim > 0:
            out[k] = torch.from_numpy(v)
        elif v.dtype.kind in "US":
            out[k] = str(v)
        else:
            out[k] = v.item()
    return out


def load_weights(net_name):
    z = np.load(os.path.join(GOLDEN, "weights", net_name + ".npz"))
    W = [torch.from_numpy(z[f"W{i}"]) for i in range(5)]
    b = [torch.from_numpy(z[f"b{i}"]) for i in range(5)]
    return W, b


def full_policy(c, key, N):
    """Golden files store only the live kernel columns; rebuild the (N, 50, ..) tensors."""
    t = c[key]
    nk = int(c["nk"])
    shape = (N, 50) + tuple(t.shape[2:])
    out = torch.zeros(shape)
    if nk > 0:
        out[:, :nk] = t[:, :nk]
    return out


def rel_err(a, b, floor=1e-6):
    a, b = a.double(), b.double()
    return ((a - b).abs() / b.abs().clamp_min(floor)).max().item()


def frac_within(a, b, rtol, atol):
    a, b = a.double(), b.double()
    ok = (a - b).abs() <= atol + rtol * b.abs()
    return ok.double().mean().item()
