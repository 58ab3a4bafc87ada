"""Times MPPI.propagate for every tile shape of the fp32 network kernels (rows per thread x features per thread,
csrc/exact_mlp.cu) on a few batch sizes, to pick the shape heuristics (pick_shape).  GPU box only.

    python tools/exact_shape_sweep.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import problem, load_net_arrays, seeded_policy  # noqa: E402
from optimalmodulationds_b200 import MPPI, LinDS  # noqa: E402
from optimalmodulationds_b200.sdf.robot_sdf import RobotSdfCollisionNet  # noqa: E402

dev = torch.device("cuda", 0)


def build(name, N, H=None):
    p = problem(name)
    p["N"] = N
    if H:
        p["H"] = H
    W, b, _ = load_net_arrays(p["net"])
    net = RobotSdfCollisionNet(in_channels=p["dof"] + 3, out_channels=p["out"], layers=[256] * 4, skips=[])
    net.load_arrays(W, b)
    t = lambda x: x.to(dev)  # noqa: E731
    m = MPPI(t(p["q0"]), t(p["qf"]), t(p["dh"]), t(p["obs"]), p["dt"], p["H"], N, [LinDS(t(p["qf"]))], t(p["dh_a"]), net,
             p["K"])
    m.dst_thr, m.ker_thr, m.ignored_links = p["dst_thr"], p["ker_thr"], list(p["ignored"])
    mu_c, sigma_c, alpha_c, mu_tmp, sigma_tmp, alpha_tmp = seeded_policy(p, N, 1)
    P = m.Policy
    P.n_kernels = p["nk"]
    P.mu_c.copy_(t(mu_c)); P.sigma_c.copy_(t(sigma_c)); P.alpha_c.copy_(t(alpha_c))
    P.mu_tmp.copy_(t(mu_tmp)); P.sigma_tmp.copy_(t(sigma_tmp)); P.alpha_tmp.copy_(t(alpha_tmp))
    return m


def time_propagate(m, reps=5):
    for _ in range(2):
        m.propagate()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        m.propagate()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / reps


