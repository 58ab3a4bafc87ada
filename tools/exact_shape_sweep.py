"""Times MPPI.propagate for every tile shape of the fp32 network kernels (see exact_shape_sweep_lib.py).  GPU box only.

    python tools/exact_shape_sweep.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from exact_shape_sweep_lib import build, time_propagate  # noqa: E402

CASES = [("planar2", 100, None), ("planar7", 250, None), ("planar7", 1000, None), ("planar7", 2000, None),
         ("planar7", 4000, None), ("planar7", 16000, 10), ("franka_shelf_294", 40, 10), ("franka_shelf_2064", 4096, 10)]
SHAPES = [(None, None), (8, 8), (4, 8), (8, 4), (4, 4), (8, 2), (4, 2)]
for name, N, H in CASES:
    m = build(name, N, H)
    row = []
    for rpt, ft in SHAPES:
        for k, v in (("DSMPPI_EXACT_RPT", rpt), ("DSMPPI_EXACT_FT", ft)):
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)
        row.append(f"{'auto' if rpt is None else f'{rpt}x{ft}'}: {time_propagate(m):8.3f}")
    print(f"{name:18s} N={N:6d} H={m.dt_H:3d} M={m.n_obs:5d} ms/propagate  " + "  ".join(row), flush=True)
