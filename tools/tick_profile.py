"""Where a graphed control tick spends its time (host side), phase by phase -- run on the GPU box.

    python tools/tick_profile.py
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from optimalmodulationds_b200 import MPPI, LinDS  # noqa: E402
from optimalmodulationds_b200.sdf.robot_sdf import RobotSdfCollisionNet  # noqa: E402

p = bench.problem("franka_shelf_294")
obs = p["obs"][:28].clone()
W, b, _ = bench.load_net_arrays("franka")
net = RobotSdfCollisionNet(in_channels=10, out_channels=9, layers=[256] * 4, skips=[])
net.load_arrays(W, b)
m = MPPI(p["q0"].clone(), p["qf"].clone(), p["dh"], obs, 0.01, 2, 1, [LinDS(p["qf"].clone()), LinDS(p["q0"].clone())],
         p["dh_a"], net, p["K"])
m.dst_thr = 0.03
P = m.Policy
P.alpha_s = 0.0
for k in range(5):
    P.add_kernel(p["q0"] + 0.1 * k, 0.1, torch.eye(7))
for _ in range(30):
    P.sample_policy(); m.propagate()

import cProfile
import pstats
pr = cProfile.Profile()
n = 2000
t0 = time.perf_counter()
pr.enable()
for _ in range(n):
    m.update_obstacles(obs)
    P.sample_policy()
    m.propagate()
    m.q_cur = torch.clamp(m.q_cur + m.qdot[0, :] * 0.01, m.Cost.q_min, m.Cost.q_max)
pr.disable()
dt = (time.perf_counter() - t0) / n
print(f"tick {dt * 1e6:.1f} us under cProfile")
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
