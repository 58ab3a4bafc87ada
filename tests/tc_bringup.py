"""Bring-up script for the tcgen05 prefilter (run on the GPU box): compares the tensor-core pass-1
distances with the fp32 ones on the Franka shelf, for the layout-debug flag given in DSMPPI_TC_FLAGS."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.golden_util import load_npz  # noqa: E402
from tests.mppi_factory import make_mppi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
c = load_npz("case_franka_shelf")
m = make_mppi(c, device="cuda", N=32, H=1)
torch.manual_seed(0)
q = c["q0"].cuda() + 0.4 * torch.randn(n, 7, device="cuda")
ex = m.debug_pass1(q, "exact")
torch.cuda.synchronize()
for mode in ("tc_f16", "tc_bf16"):
    t0 = time.time()
    tc = m.debug_pass1(q, mode)
    torch.cuda.synchronize()
    dt = time.time() - t0
    err = (tc - ex).abs()
    print(f"flags={os.environ.get('DSMPPI_TC_FLAGS', '0')} {mode}: n={n} M={ex.shape[1]} max|err|={err.max().item():.4e} "
          f"mean|err|={err.mean().item():.4e} exact range [{ex.min().item():.3f},{ex.max().item():.3f}] "
          f"tc range [{tc.min().item():.3f},{tc.max().item():.3f}] nan={int(torch.isnan(tc).sum())} ({dt*1e3:.1f} ms)")
    if mode == "tc_f16":
        print("   sample exact:", [f"{v:.4f}" for v in ex[0, :6].tolist()])
        print("   sample tc   :", [f"{v:.4f}" for v in tc[0, :6].tolist()])
