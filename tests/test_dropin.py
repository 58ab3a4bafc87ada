"""The drop-in seam: the reference's scripts import `MPPI`, `LinDS`, `policy`, `cost`, `fk_num`, `plots` and
`sdf.robot_sdf` by those bare names (standalonePlanar2d.py:7-12, frankaPlanner.py:3-9); with
optimalmodulationds_b200/dropin first on PYTHONPATH they must resolve to this package and expose the
star-import namespace the scripts rely on."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "optimalmodulationds_b200", "dropin")

PROBE = r"""
import sys
sys.path.append('../functions/')            # what the reference scripts do; must lose against PYTHONPATH
from MPPI import *
sys.path.append('../../mlp_learn/')
from sdf.robot_sdf import RobotSdfCollisionNet
from LinDS import *
import MPPI as M, policy, cost, fk_num, plots, LinDS as L
need = ['torch', 'time', 'np', 'plt', 'pi', 'profile', 'record_function', 'ProfilerActivity', 'numeric_fk_model',
        'numeric_fk_model_vec', 'generalized_sigmoid', 'TensorPolicyMPPI', 'eval_rbf', 'Cost', 'MPPI', 'LinDS',
        'init_robot_plot', 'init_kernel_means', 'upd_r_h', 'plot_obs_init', 'plot_obs_update', 'init_jpos_plot',
        'upd_jpos_plot', 'dh_fk']
missing = [n for n in need if n not in globals()]
assert not missing, missing
for mod in (M, policy, cost, fk_num, plots, L):
    assert 'optimalmodulationds_b200' in mod.__file__, mod.__file__
net = RobotSdfCollisionNet(in_channels=10, out_channels=7, layers=[256] * 4, skips=[])
assert net.in_channels == 10 and net.out_channels == 7 and hasattr(net, 'model')
ds = LinDS(torch.zeros(7))
assert ds.get_velocity(torch.ones(3, 7)).shape == (3, 7)
print('ok')
"""


def test_reference_import_lines_resolve_to_dropin():
    env = dict(os.environ, PYTHONPATH=DROPIN + os.pathsep + os.environ.get("PYTHONPATH", ""))
    out = subprocess.run([sys.executable, "-c", PROBE], env=env, cwd="/tmp", capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().endswith("ok")


@pytest.mark.gpu
def test_script_style_planner_loop_runs_through_dropin():
    """examples/planar7_loop.py mirrors standalonePlanar7d.py's loop (CPU tensors, attribute pokes, second 1x1
    MPPI object, kernel adding): it must run headless and drive the arm towards the goal without collisions."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "planar7_loop.py"), "--iters", "150"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][-1]
    kv = dict(tok.split("=") for tok in line.split()[1:])
    assert int(kv["iterations"]) == 150 or float(kv["final_dist"]) <= 0.1
    assert float(kv["final_dist"]) < float(kv["start_dist"]) - 0.2, line
    print(line)
