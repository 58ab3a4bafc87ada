"""CPU tests of the SEDS nominal DS: the oracle's restatement (oracle.seds_velocity) and the host-side mirror class
(optimalmodulationds_b200.SEDS) against outputs of the UNMODIFIED reference class (tests/golden/seds_*.npz,
sedscase_*.npz from tests/golden/make_golden_seds.py)."""
import pytest
import torch

from oracle import mppi_oracle as orc
from tests.golden_util import frac_within, load_npz, load_weights

MODELS = ["seds_left10", "seds_2d", "seds_sine"]


def _check(v, ref, name):
    # GMR outputs are sums of G terms with cancellation near the goal: relative to the row's magnitude
    scale = ref.norm(dim=1, keepdim=True).clamp_min(1e-3)
    err = ((v - ref).abs() / scale).max(dim=1)[0]
    assert (err < 1e-4).float().mean() >= 0.97, f"{name}: {err.max():.3e}"
    assert (err < 5e-3).all(), f"{name}: {err.max():.3e}"


@pytest.mark.parametrize("model", MODELS)
def test_oracle_seds_velocity(model):
    c = load_npz(model)
    sp = orc.SedsParams(c["Mu"], c["Sigma"], c["Priors"])
    _check(orc.seds_velocity(c["x"], c["xT"], sp), c["velocity"], model)


@pytest.mark.parametrize("model", MODELS)
def test_host_mirror_class(model):
    from optimalmodulationds_b200.SEDS import SEDS
    c = load_npz(model)
    ds = SEDS.from_arrays(c["Mu"], c["Sigma"], c["Priors"], c["xT"])
    assert ds.dof == c["x"].shape[1] and ds.n_gaussians == c["Sigma"].shape[2]
    assert ds.Sigma_inv.shape == (ds.dof, ds.dof, ds.n_gaussians) and ds.det.shape == (ds.n_gaussians,)
    _check(ds.get_velocity(c["x"]), c["velocity"], model)
    # one state at a time, the only call shape the reference's indexing supports (SEDS.py:70-71)
    _check(torch.cat([ds.get_velocity(c["x"][i:i + 1]) for i in range(8)]), c["velocity"][:8], model)


def test_oracle_rollout_with_seds_vs_reference():
    c = load_npz("sedscase_franka_left10")
    m = load_npz(c["model"])
    W, b = load_weights("franka")
    H, dt = int(c["H"]), float(c["dt"])
    sp = orc.SedsParams(m["Mu"], m["Sigma"], m["Priors"])
    z = torch.zeros(1, 50, 7)
    prm = orc.RolloutParams(dt=dt, dt_H=H, n_closest_obs=5, dst_thr=0.01, ignored_links=[0, 1, 2], seds=sp,
                            with_basis=False)
    o = orc.rollout(orc.Net(W, b), c["q0"], c["qf"], c["obs"], z, torch.zeros(1, 50), z, 0, prm, 1)
    assert frac_within(o.all_traj, c["all_traj"], 1e-4, 1e-5) == 1.0
    assert frac_within(o.closest_dist_all, c["closest_dist_all"], 1e-4, 1e-5) == 1.0
    assert frac_within(o.qdot, c["qdot"], 1e-4, 1e-5) == 1.0
