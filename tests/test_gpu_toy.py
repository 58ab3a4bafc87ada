"""GPU parity tests of the toy variant (run with -m gpu on a B200): optimalmodulationds_b200.MPPI_toy.MPPI --
the same CUDA kernels with the MPPI_toy.py constant set, matrix nominal DS, 2-coordinate obstacles, three-term
cost and un-masked policy update -- against the committed outputs of the UNMODIFIED reference class
(tests/golden/toycase_*.npz) and against the oracle on fresh inputs.  Tolerances as in test_gpu_parity.py."""
import glob
import os

import pytest
import torch

from oracle import mppi_oracle as orc
from tests.golden_util import GOLDEN, frac_within, full_policy, load_npz, load_weights

pytestmark = pytest.mark.gpu

TOY_CASES = sorted(os.path.basename(p)[8:-4] for p in glob.glob(os.path.join(GOLDEN, "toycase_*.npz")))


def check(a, b, rtol, atol, name, min_frac=0.99, loose=20):
    a, b = a.detach().cpu(), b.detach().cpu()
    assert a.shape == b.shape, f"{name}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    f = frac_within(a, b, rtol, atol)
    min_frac = min(min_frac, 1.0 - 1.0 / max(a.numel(), 1)) if min_frac < 1.0 else 1.0
    assert f >= min_frac, f"{name}: only {f:.4f} within rtol={rtol} (max abs diff {(a - b).abs().max():.3e})"
    assert frac_within(a, b, loose * rtol, loose * atol) == 1.0, \
        f"{name}: outliers beyond {loose}x tolerance (max abs diff {(a - b).abs().max():.3e})"


@pytest.fixture(scope="module")
def factory():
    from tests import mppi_factory
    return mppi_factory


@pytest.mark.parametrize("tag", TOY_CASES)
def test_toy_one_step_map_teacher_forced_vs_reference(tag, factory):
    c = load_npz(f"toycase_{tag}")
    N, H, nk, dt = int(c["N"]), int(c["H"]), int(c["nk"]), float(c["dt"])
    m = factory.make_toy_mppi(c, device="cpu", H=1)
    for t in range(H):
        q = c["all_traj"][:, t, :]
        m.q_cur = q
        out = m.propagate()
        assert len(out) == 4                                      # MPPI_toy.py:207: no kernel_activations
        traj, dist, kv, dots = out
        assert traj.device.type == "cpu" and traj.shape == (N, 1, 2)
        check(dist[:, 0], c["closest_dist_all"][:, t], 1e-5, 2e-6, f"dist[{t}]")
        check(dots[:, 0], c["dot_products"][:, t], 1e-5, 1e-5, f"dot[{t}]", min_frac=0.9)
        check(m.norm_basis[:, 0], c["norm_basis"][:, t], 1e-4, 1e-5, f"basis[{t}]")
        if nk > 0:
            check(kv[:, 0, :], c["kernel_val_all"][:, t, :nk], 1e-4, 1e-6, f"kval[{t}]")
        if t == 0:
            check(m.qdot, c["qdot"], 1e-5, 1e-5, "qdot")
        if t + 1 < H:
            check(q + dt * m.qdot, c["all_traj"][:, t + 1, :], 1e-5, 1e-5, f"traj[{t + 1}]")


@pytest.mark.parametrize("tag", TOY_CASES)
@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_toy_iteration_vs_reference(tag, device, factory):
    """propagate -> get_cost -> shift_policy_means, free-running over the whole horizon."""
    c = load_npz(f"toycase_{tag}")
    nk = int(c["nk"])
    m = factory.make_toy_mppi(c, device=device)
    traj, dist, kv, dots = m.propagate()
    assert traj.device.type == device
    # free-running over the horizon: first step at the one-step tolerance (1e-5), the rest at the full-horizon
    # tolerance of BASELINE.json's north_star (1e-4) -- per-step differences of ~1e-7 accumulate along a rollout
    check(dist[:, 0], c["closest_dist_all"][:, 0], 1e-5, 2e-6, "closest_dist_all[0]")
    check(dist, c["closest_dist_all"], 1e-4, 1e-5, "closest_dist_all")
    check(dots, c["dot_products"], 1e-4, 1e-5, "dot_products", min_frac=0.9)
    check(m.qdot, c["qdot"], 1e-5, 1e-5, "qdot")
    check(traj, c["all_traj"], 1e-4, 1e-5, "all_traj")
    if nk > 0:
        check(kv, c["kernel_val_all"][:, :, :nk], 1e-4, 1e-6, "kernel_val_all")
    check(m.norm_basis, c["norm_basis"], 1e-4, 1e-5, "norm_basis")
    cost = m.get_cost()
    if torch.isfinite(c["cost"]).all():
        check(cost, c["cost"], 1e-4, 1e-3, "cost")
    else:
        assert torch.equal(torch.isfinite(cost.cpu()), torch.isfinite(c["cost"]))
        return
    assert m.shift_policy_means() == 0                             # MPPI_toy.py:324
    check(m.Policy.mu_c, c["mu_c1"], 1e-4, 1e-5, "mu_c")
    check(m.Policy.sigma_c, c["sigma_c1"], 1e-4, 1e-5, "sigma_c")
    check(m.Policy.alpha_c, c["alpha_c1"], 1e-4, 1e-5, "alpha_c")


def test_toy_update_given_reference_trajectories(factory):
    """Cost and policy update on the reference's own trajectories / kernel values (no rollout in between)."""
    c = load_npz("toycase_toy2_near")
    N, H, nk = int(c["N"]), int(c["H"]), int(c["nk"])
    m = factory.make_toy_mppi(c, device="cpu")
    m.propagate()
    m.all_traj, m.closest_dist_all = c["all_traj"].clone(), c["closest_dist_all"].clone()
    kv = torch.zeros(N, H, 50)
    kv[:, :, :nk] = c["kernel_val_all"]
    m.kernel_val_all = kv
    cost = m.get_cost()
    check(cost, c["cost"], 1e-5, 1e-4, "cost")
    m.shift_policy_means()
    check(m.Policy.mu_c, c["mu_c1"], 1e-5, 1e-6, "mu_c")
    check(m.Policy.alpha_c, c["alpha_c1"], 1e-5, 1e-6, "alpha_c")


def test_toy_matches_oracle_on_fresh_inputs_and_tensor_core_mode(factory):
    """Ragged sizes, many obstacles (tensor-core prefilter on the 2-coordinate net) vs the oracle; tc mode must equal
    the all-fp32 mode bit for bit."""
    c = load_npz("toycase_toy2")
    g = torch.Generator().manual_seed(5)
    M, N, H, nk = 97, 37, 5, 3
    ang = torch.rand(M, generator=g) * 6.283
    rad = 3 + 2 * torch.rand(M, generator=g)
    obs = torch.stack((rad * torch.cos(ang), rad * torch.sin(ang), 0.2 + 0.3 * torch.rand(M, generator=g)), 1)
    c = dict(c, obs=obs, nk=nk, K=3)      # N / H go through the factory arguments (fresh policy samples)
    c["q_cur"] = c["q0"] = torch.tensor([-0.3, 0.2])
    outs = {}
    for mode in ("exact", "tc_f16"):
        m = factory.make_toy_mppi(c, device="cuda", N=N, H=H, pass1=mode)
        torch.manual_seed(11)
        m.Policy.sample_policy()
        mu, sg, al = m.Policy.mu_tmp.cpu().clone(), m.Policy.sigma_tmp.cpu().clone(), m.Policy.alpha_tmp.cpu().clone()
        outs[mode] = [x.cpu().clone() for x in m.propagate()] + [m.qdot.cpu().clone()]
        if mode == "tc_f16":
            st = m.pass1_stats()
            assert st["mode"] == 1 and st["band_overflows"] == 0
    for a, b in zip(outs["exact"], outs["tc_f16"]):
        assert torch.equal(a, b)
    W, b = load_weights("toy2")
    prm = orc.toy_params(float(c["dt"]), H, 3, c["A"], dst_thr=float(c["dst_thr"]), p=float(c["p"]), with_basis=False)
    o = orc.rollout(orc.Net(W, b), c["q_cur"], c["qf"], obs, mu, sg, al, nk, prm, N)
    traj, dist, kv, dots, qdot = outs["exact"]
    check(dist[:, :2], o.closest_dist_all[:, :2], 1e-5, 2e-6, "dist")
    check(qdot, o.qdot, 1e-5, 1e-5, "qdot")
    check(traj[:, :2], o.all_traj[:, :2], 1e-4, 1e-5, "traj")
    check(kv[:, :2], o.kernel_val_all[:, :2, :nk], 1e-4, 1e-6, "kval")


def test_toy_dropin_namespace():
    """`from MPPI_toy import *` must hand the script the same names as the reference module does."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); from MPPI_toy import *; "
            "print(MPPI.__module__, Cost.__module__, callable(generalized_sigmoid), callable(init_toy_plot), "
            "callable(eval_rbf), TensorPolicyMPPI.__name__, pi > 3, torch.__name__, np.__name__)"
            % os.path.join(root, "optimalmodulationds_b200", "dropin"))
    out = subprocess.check_output([sys.executable, "-c", code], text=True).split()
    assert out[0] == "optimalmodulationds_b200.MPPI_toy" and out[1] == "optimalmodulationds_b200.cost_toy"
    assert out[2:] == ["True", "True", "True", "TensorPolicyMPPI", "True", "torch", "numpy"]
