"""Bitwise regression harness for the tensor-core scoring kernel (tc_exact.cu): a restructuring of its pipeline must not
change a single bit, because every product sum keeps its k-order and its three MMAs.  `run TAG` (on the GPU box) writes
the kernel's outputs on fixed seeded inputs to gpurun_out/tcx_regress_TAG.pt and prints timings; `cmp A B` (anywhere)
compares two such files bit for bit."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(f, n=5):
    f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def run(tag):
    from tests.golden_util import load_npz
    from tests.mppi_factory import make_mppi
    out = {}
    for case, n, M, pass1 in (("planar2", 1000, None, "exact"), ("planar7", 1000, None, "exact"),
                              ("planar2", 300000, None, "exact"),
                              ("franka_shelf", 2048, 40, "exact"), ("franka_shelf", 4096, 300, "tc_f16")):
        torch.manual_seed(1)
        c = load_npz(f"case_{case}")
        if M is not None:
            obs = torch.rand(M, 4) * 1.2 - 0.6
            obs[:, 3] = 0.03
            c["obs"] = obs
        m = make_mppi(c, device="cuda", pass1=pass1)
        m.set_score_mode("tc_split")
        d = c["q0"].shape[0]
        q = ((torch.rand(n, d) * 2 - 1) * 2.5).cuda()
        dist, grad = m.distance_repulsion_nn(q)
        key = f"{case}_{n}_{M}_{pass1}"
        out[key + "_dist"] = dist.cpu().clone()
        out[key + "_grad"] = grad.cpu().clone()
        if n <= 4096:
            out[key + "_p1"] = m.debug_pass1(q, "exact").cpu().clone()
        ms = timed(lambda: m.distance_repulsion_nn(q))
        print(f"{key:36s} distance_repulsion_nn {ms:8.3f} ms")
    # whole-horizon kernel (MODE 2) and the per-step sequence through a full MPPI iteration
    for case in ("planar7", "planar2", "franka_shelf"):
        c = load_npz(f"case_{case}")
        m = make_mppi(c, device="cuda", pass1="auto")
        m.set_score_mode("tc_split")
        traj, dist, kv, dots, acts = m.propagate()
        cost = m.get_cost()
        out[f"roll_{case}_traj"] = traj.cpu().clone()
        out[f"roll_{case}_dist"] = dist.cpu().clone()
        out[f"roll_{case}_dots"] = dots.cpu().clone()
        out[f"roll_{case}_cost"] = cost.cpu().clone()
        ms = timed(lambda: m.propagate())
        print(f"rollout {case:20s} propagate {ms:8.3f} ms")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    torch.save(out, os.path.join(ROOT, "gpurun_out", f"tcx_regress_{tag}.pt"))


def cmp(a, b):
    A = torch.load(os.path.join(ROOT, "gpurun_out", f"tcx_regress_{a}.pt"))
    B = torch.load(os.path.join(ROOT, "gpurun_out", f"tcx_regress_{b}.pt"))
    bad = 0
    for k in A:
        x, y = A[k], B[k]
        same = x.shape == y.shape and bool((x.view(torch.int32) == y.view(torch.int32)).all())
        nd = int((x.view(torch.int32) != y.view(torch.int32)).sum()) if x.shape == y.shape else -1
        md = float((x.double() - y.double()).abs().max()) if x.shape == y.shape else float("nan")
        print(f"{k:40s} {'bitwise equal' if same else f'DIFFERENT ({nd} of {x.numel()} elements, max |diff| {md:.3e})'}")
        bad += not same
    print("ALL BITWISE EQUAL" if bad == 0 else f"{bad} tensors differ")
    return bad


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(sys.argv[2])
    else:
        sys.exit(1 if cmp(sys.argv[2], sys.argv[3]) else 0)
