"""GPU parity tests of the SEDS nominal DS inside the CUDA rollout (run with -m gpu on a B200): an MPPI object
whose DS_ARRAY holds a SEDS (frankaIntegrator.py:70-71) against the committed output of the UNMODIFIED reference
(tests/golden/sedscase_*.npz; N_traj = 1, the only batch size the reference's SEDS supports) and, for real
batches, against the oracle (pinned on the same reference outputs by tests/test_seds_oracle_golden.py)."""
import math

import pytest
import torch

from oracle import mppi_oracle as orc
from tests.golden_util import frac_within, load_npz, load_weights

pytestmark = pytest.mark.gpu


def check(a, b, rtol, atol, name, min_frac=0.99, loose=20):
    a, b = a.detach().cpu(), b.detach().cpu()
    assert a.shape == b.shape, f"{name}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    f = frac_within(a, b, rtol, atol)
    min_frac = min(min_frac, 1.0 - 1.0 / max(a.numel(), 1)) if min_frac < 1.0 else 1.0
    assert f >= min_frac, f"{name}: only {f:.4f} within rtol={rtol} (max abs diff {(a - b).abs().max():.3e})"
    assert frac_within(a, b, loose * rtol, loose * atol) == 1.0, \
        f"{name}: outliers beyond {loose}x tolerance (max abs diff {(a - b).abs().max():.3e})"


def franka_dh():
    pi = math.pi
    dh_a = torch.tensor([0, 0, 0, 0.0825, -0.0825, 0, 0.088, 0])
    dh_d = torch.tensor([0.333, 0, 0.316, 0, 0.384, 0, 0, 0.107])
    dh_alpha = torch.tensor([0, -pi / 2, pi / 2, pi / 2, -pi / 2, pi / 2, pi / 2, 0])
    return torch.vstack((dh_d, dh_a * 0, dh_a, dh_alpha)).T.contiguous(), dh_a


def test_seds_rollout_vs_reference_single_sample():
    from optimalmodulationds_b200 import MPPI, SEDS
    from tests.mppi_factory import make_net
    c = load_npz("sedscase_franka_left10")
    mdl = load_npz(c["model"])
    ds = SEDS.from_arrays(mdl["Mu"], mdl["Sigma"], mdl["Priors"], mdl["xT"])
    dh, dh_a = franka_dh()
    H, dt = int(c["H"]), float(c["dt"])
    m = MPPI(c["q0"], c["qf"], dh, c["obs"], dt, H, 1, [ds], dh_a, make_net("franka"), 5)
    m.dst_thr, m.ker_thr, m.ignored_links = 0.01, 0.1, [0, 1, 2]
    m.Policy.sample_policy()
    traj, dist, kv, dots, acts = m.propagate()
    check(traj, c["all_traj"], 1e-4, 1e-5, "all_traj", min_frac=1.0)
    check(dist, c["closest_dist_all"], 1e-4, 1e-5, "closest_dist_all", min_frac=1.0)
    check(m.qdot, c["qdot"], 1e-5, 1e-5, "qdot", min_frac=1.0)
    check(dots, c["dot_products"], 1e-4, 1e-5, "dot_products", min_frac=1.0)


@pytest.mark.parametrize("model,netname", [("seds_left10", "franka"), ("seds_sine", "franka"), ("seds_2d", "planar2")])
def test_seds_batched_rollout_vs_oracle(model, netname):
    """Whole batches (which the reference's SEDS cannot evaluate): per-sample start states on and off the
    demonstrated region, H = 3, against the oracle restatement."""
    from optimalmodulationds_b200 import MPPI, SEDS
    from tests.mppi_factory import make_net
    mdl = load_npz(model)
    ds = SEDS.from_arrays(mdl["Mu"], mdl["Sigma"], mdl["Priors"], mdl["xT"])
    d = ds.dof
    x = mdl["x"]
    N, H, dt = x.shape[0], 3, 0.02
    goal = mdl["xT"].reshape(-1)
    if d == 7:
        dh, dh_a = franka_dh()
        obs = torch.tensor([[5.0, 5.0, 5.0, 0.03], [4.0, 5.0, 5.0, 0.03]])
        ignored, K = [0, 1, 2], 2
    else:
        dh_a = torch.zeros(3); dh_a[1:] = 3
        dh = torch.vstack((dh_a * 0, dh_a * 0, dh_a, dh_a * 0)).T.contiguous()
        obs = torch.tensor([[60.0, 0, 0, .5], [0.0, 45, 0, .5]])
        ignored, K = [], 1
    m = MPPI(x[0].clone(), goal, dh, obs, dt, H, N, [ds], dh_a, make_net(netname), K)
    m.dst_thr, m.ignored_links = 0.01, ignored
    m.q_cur = x.clone()
    m.Policy.sample_policy()
    traj, dist, kv, dots, acts = m.propagate()
    W, b = load_weights(netname)
    sp = orc.SedsParams(mdl["Mu"], mdl["Sigma"], mdl["Priors"])
    prm = orc.RolloutParams(dt=dt, dt_H=H, n_closest_obs=K, dst_thr=0.01, ignored_links=ignored, seds=sp,
                            with_basis=False)
    z = torch.zeros(N, 50, d)
    o = orc.rollout(orc.Net(W, b), x, goal, obs, z, torch.zeros(N, 50), z, 0, prm, N)
    # the nominal velocity itself: obstacles are far away, so the first modulated velocity is v (normalised when
    # |v| > 0.5, MPPI.py:211-213) -- rows whose GMR output is a cancelling sum get a magnitude-relative tolerance
    scale = o.qdot.norm(dim=1, keepdim=True).clamp_min(1e-3)
    err = ((m.qdot.cpu() - o.qdot).abs() / scale).max(dim=1)[0]
    assert (err < 1e-4).float().mean() >= 0.97 and (err < 5e-3).all(), err.max()
    check(traj, o.all_traj, 1e-4, 1e-5, "all_traj")


def test_switching_between_seds_and_linds_objects():
    """DS objects are read at call time (switch_DS_idx, frankaIntegratorSwitching.py): the kernel must follow."""
    from optimalmodulationds_b200 import MPPI, SEDS, LinDS
    from tests.mppi_factory import make_net
    mdl = load_npz("seds_left10")
    ds = SEDS.from_arrays(mdl["Mu"], mdl["Sigma"], mdl["Priors"], mdl["xT"])
    goal = mdl["xT"].reshape(-1)
    dh, dh_a = franka_dh()
    obs = torch.tensor([[5.0, 5.0, 5.0, 0.03]])
    q0 = goal + mdl["Mu"][:7, 0]
    m = MPPI(q0, goal, dh, obs, 0.02, 2, 4, [ds, LinDS(goal)], dh_a, make_net("franka"), 1)
    m.Policy.sample_policy()
    m.propagate()
    v_seds = m.qdot.clone()
    m.switch_DS_idx(1)
    m.propagate()
    v_lin = m.qdot.clone()
    m.switch_DS_idx(0)
    m.propagate()
    assert torch.equal(m.qdot, v_seds) and not torch.allclose(v_seds, v_lin)
    lin = -(q0 - goal) / (q0 - goal).norm()
    assert torch.allclose(v_lin[0], lin, atol=1e-5)
