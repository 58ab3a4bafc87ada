"""Bring-up / timing of the tensor-core scoring kernel (tc_exact.cu) against the FFMA kernel (exact_mlp.cu) through the
public API: distance_repulsion_nn on random states, every row-source mode (dense pairs, selected pairs, candidate list),
plus the forward-only pass (debug_pass1).  Run on the GPU box."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.golden_util import load_npz  # noqa: E402
from tests.mppi_factory import make_mppi  # noqa: E402


def stats(name, a, b):
    a, b = a.double().cpu(), b.double().cpu()
    e = (a - b).abs()
    scale = b.pow(2).mean().sqrt().item()
    rel = e / b.abs().clamp_min(1e-3 * scale)
    print(f"    {name:10s} max|diff|/rms {e.max().item() / scale:9.2e}  rms|diff|/rms {e.pow(2).mean().sqrt().item() / scale:9.2e}"
          f"  frac(rel<1e-5) {(rel < 1e-5).double().mean().item():.4f}  nan {int(torch.isnan(a).sum())}")


def timed(f, n=5):
    f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    torch.manual_seed(1)
    for case, n, M, pass1 in (("planar2", 1000, None, "exact"), ("planar7", 1000, None, "exact"),
                              ("franka_shelf", 2048, 40, "exact"), ("franka_shelf", 4096, 300, "tc_f16"),
                              ("franka_shelf", 4096, 300, "exact")):
        c = load_npz(f"case_{case}")
        if M is not None:
            obs = torch.rand(M, 4) * 1.2 - 0.6
            obs[:, 3] = 0.03
            c["obs"] = obs
        m = make_mppi(c, device="cuda", pass1=pass1)
        d = c["q0"].shape[0]
        q = ((torch.rand(n, d) * 2 - 1) * 2.5).cuda()
        print(f"{case}: n={n} M={m.obs.shape[0]} K={m.n_closest_obs} pass1={pass1}")
        out = {}
        for mode in ("ffma", "tc_split", "tc_split:4"):       # :4 = DSMPPI_TCX_DEBUG=4, no truncation compensation
            os.environ["DSMPPI_TCX_DEBUG"] = mode.split(":")[1] if ":" in mode else "0"
            m.set_score_mode(mode.split(":")[0])
            dist, grad = m.distance_repulsion_nn(q)
            p1 = m.debug_pass1(q, "exact")
            torch.cuda.synchronize()
            out[mode] = (dist.clone(), grad.clone(), p1.clone())
            ms = timed(lambda: m.distance_repulsion_nn(q))
            print(f"  {mode:9s} distance_repulsion_nn {ms:8.3f} ms")
        for mode in out:
            if mode != "ffma":
                print("  ", mode)
                stats("dist", out[mode][0], out["ffma"][0])
                stats("grad", out[mode][1], out["ffma"][1])
                stats("pass1", out[mode][2], out["ffma"][2])


if __name__ == "__main__":
    main()
