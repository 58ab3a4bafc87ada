"""Runs a few propagate() calls of a small workload (for ncu captures of the whole-horizon kernel).
    python tools/run_small.py [workload] [N] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from exact_shape_sweep_lib import build  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "planar2"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
m = build(name, N)
for _ in range(reps):
    m.propagate()
torch.cuda.synchronize()
print("done", name, N, m.dt_H, m.n_obs)
