"""Latency of the two Franka control loops of the reference (ds_mppi/frankaIntegrator.py:101-121: N=1, H=2 every
control tick; ds_mppi/frankaPlanner.py:99-190: N=40, H=10) through the drop-in API, CPU tensors in/out as those
scripts use, and with CUDA tensors.  Prints loop frequency; the reference logs ~500 Hz for the integrator
(experiment_logs/my_*.txt) on its authors' CPU.

    python tools/latency_bench.py [--obstacles 28] [--iters 300]
"""
import argparse
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_net_arrays, shelf  # noqa: E402
from optimalmodulationds_b200 import MPPI, LinDS  # noqa: E402
from optimalmodulationds_b200.sdf.robot_sdf import RobotSdfCollisionNet  # noqa: E402


def build(device, N, H, obs, K=5):
    pi = math.pi
    t = lambda x: x.to(device)  # noqa: E731
    dh_a = torch.tensor([0, 0, 0, 0.0825, -0.0825, 0, 0.088, 0])
    dh_d = torch.tensor([0.333, 0, 0.316, 0, 0.384, 0, 0, 0.107])
    dh_alpha = torch.tensor([0, -pi / 2, pi / 2, pi / 2, -pi / 2, pi / 2, pi / 2, 0])
    dh = torch.vstack((dh_d, dh_a * 0, dh_a, dh_alpha)).T.contiguous()
    q0 = torch.tensor([-0.88, 0.38, 0.5, -1, 0.45, 1.9, 0.31])
    qf = torch.tensor([-1.24, 1.53, 1.22, -1.21, -0.21, 1.55, 0.08])
    W, b, _ = load_net_arrays("franka")
    net = RobotSdfCollisionNet(in_channels=10, out_channels=9, layers=[256] * 4, skips=[])
    net.load_arrays(W, b)
    m = MPPI(t(q0), t(qf), t(dh), t(obs), 0.01 if N == 1 else 0.5, H, N, [LinDS(t(qf)), LinDS(t(q0))], t(dh_a), net, K)
    m.dst_thr, m.ker_thr, m.ignored_links = 0.01, 0.1, [0, 1, 2]
    P = m.Policy
    P.alpha_s, P.sigma_c_nominal = 3.0, 1.0
    for k in range(5):                                  # a few kernels so the policy blend does work
        P.add_kernel(t(q0) + 0.1 * k, 0.1, torch.eye(7, device=device))
    P.alpha_c[:5] = 0.3
    return m


def loop(m, iters, planner):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        m.Policy.sample_policy()
        m.propagate()
        if planner:
            m.get_cost()
            m.shift_policy_means()
        else:
            m.q_cur = m.q_cur + m.qdot[0, :] * 0.001      # frankaIntegrator.py:121 (reads qdot on the host side)
    torch.cuda.synchronize()
    return iters / (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--obstacles", type=int, default=28)
    ap.add_argument("--iters", type=int, default=300)
    args = ap.parse_args()
    obs = shelf(12)[:args.obstacles] if args.obstacles < 294 else shelf(12)
    for device in ("cpu", "cuda"):
        for name, N, H, planner in (("integrator N=1 H=2", 1, 2, False), ("planner N=40 H=10", 40, 10, True)):
            m = build(device, N, H, obs)
            loop(m, 20, planner)
            hz = loop(m, args.iters if not planner else max(20, args.iters // 5), planner)
            print(f"{name:22s} caller tensors on {device:4s} M={obs.shape[0]:4d}: {hz:8.1f} Hz "
                  f"({1e3 / hz:.3f} ms per loop, {N * H * hz:.0f} state-steps/s)")


if __name__ == "__main__":
    main()
