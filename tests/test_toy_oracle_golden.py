"""CPU tests of the toy variant: the oracle with the MPPI_toy.py constant set (oracle.toy_params, matrix DS,
cost_toy, un-masked policy update) against outputs of the UNMODIFIED reference class
ds_mppi/functions/MPPI_toy.py (tests/golden/toycase_*.npz, produced by tests/golden/make_golden_toy.py)."""
import glob
import os

import pytest
import torch

from oracle import mppi_oracle as orc
from tests.golden_util import GOLDEN, frac_within, full_policy, load_npz, load_weights

TOY_CASES = sorted(os.path.basename(p)[8:-4] for p in glob.glob(os.path.join(GOLDEN, "toycase_*.npz")))


def check(a, b, rtol, atol, name, min_frac=0.99, loose=20):
    f = frac_within(a, b, rtol, atol)
    min_frac = min(min_frac, 1.0 - 1.0 / max(a.numel(), 1)) if min_frac < 1.0 else 1.0
    assert f >= min_frac, f"{name}: only {f:.4f} within rtol={rtol}"
    assert frac_within(a, b, loose * rtol, loose * atol) == 1.0, f"{name}: outliers beyond {loose}x tolerance"


def test_cases_present():
    assert {"toy2", "toy2_near", "toy2_rot", "toy2_step"} <= set(TOY_CASES)


@pytest.mark.parametrize("tag", TOY_CASES)
def test_toy_one_step_map_teacher_forced(tag):
    c = load_npz(f"toycase_{tag}")
    W, b = load_weights("toy2")
    net = orc.Net(W, b)
    assert net.n_in == 4 and net.n_out == 1
    N, H, nk, dt = int(c["N"]), int(c["H"]), int(c["nk"]), float(c["dt"])
    mu, sg, al = (full_policy(c, k, N) for k in ("mu_tmp", "sigma_tmp", "alpha_tmp"))
    prm = orc.toy_params(dt, 1, int(c["K"]), c["A"], dst_thr=float(c["dst_thr"]), p=float(c["p"]))
    for t in range(H):
        q = c["all_traj"][:, t, :]
        o = orc.rollout(net, q, c["qf"], c["obs"], mu, sg, al, nk, prm, N)
        check(o.closest_dist_all[:, 0], c["closest_dist_all"][:, t], 1e-5, 2e-6, f"dist[{t}]")
        check(o.dot_products[:, 0], c["dot_products"][:, t], 1e-5, 1e-5, f"dot[{t}]")
        check(o.norm_basis[:, 0], c["norm_basis"][:, t], 1e-4, 1e-5, f"basis[{t}]")
        if nk > 0:   # MPPI_toy.py:178-179: stored kernel values carry the activation
            check(o.kernel_val_all[:, 0, :nk], c["kernel_val_all"][:, t, :nk], 1e-4, 1e-6, f"kval[{t}]")
        if t == 0:
            check(o.qdot, c["qdot"], 1e-5, 1e-5, "qdot")
        if t + 1 < H:
            check(q + dt * o.qdot, c["all_traj"][:, t + 1, :], 1e-5, 1e-5, f"traj[{t + 1}]")


@pytest.mark.parametrize("tag", TOY_CASES)
def test_toy_cost_and_policy_update(tag):
    c = load_npz(f"toycase_{tag}")
    nk, N, H = int(c["nk"]), int(c["N"]), int(c["H"])
    cost = orc.evaluate_costs_toy(c["all_traj"], c["closest_dist_all"], c["qf"])
    if not torch.isfinite(c["cost"]).all():
        assert torch.equal(torch.isfinite(cost), torch.isfinite(c["cost"]))
        return
    check(cost, c["cost"], 1e-5, 1e-4, "cost")
    mu, sg, al = (full_policy(c, k, N) for k in ("mu_tmp", "sigma_tmp", "alpha_tmp"))
    kv = torch.zeros(N, H, 50)
    kv[:, :, :max(nk, 1)] = c["kernel_val_all"]
    mu1, sg1, al1, n_upd, _ = orc.policy_update(c["cost"], kv, torch.zeros(N, H), mu, sg, al, c["mu_c0"],
                                                c["sigma_c0"], c["alpha_c0"], nk, float(c["ker_thr"]),
                                                upd_rate=float(c["upd_rate"]), toy=True)
    check(mu1, c["mu_c1"], 1e-5, 1e-6, "mu_c")
    check(sg1, c["sigma_c1"], 1e-5, 1e-6, "sigma_c")
    check(al1, c["alpha_c1"], 1e-5, 1e-6, "alpha_c")
    assert n_upd >= int(c["n_changed"])
