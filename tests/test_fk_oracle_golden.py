"""CPU tests of the true-distance (FK) provider: the oracle's restatement of MPPI.distance_repulsion_fk against
outputs of the UNMODIFIED reference (tests/golden/fkdist_*.npz, fkcase_*.npz from tests/golden/make_golden_fk.py)."""
import pytest
import torch

from oracle import mppi_oracle as orc
from tests.golden_util import frac_within, full_policy, load_npz, load_weights


@pytest.mark.parametrize("tag", ["planar7", "planar7_many"])
def test_fk_distance_and_gradient(tag):
    c = load_npz(f"fkdist_{tag}")
    dist, grad, idx = orc.distance_repulsion_fk(c["q"], c["obs"], c["dh_params"], 10, return_idx=True)
    assert torch.equal(idx, c["idx"])                          # [obstacle, link, point] of the minimum
    assert frac_within(dist, c["distance"], 1e-5, 2e-6) == 1.0
    ref = c["rep_vec"]
    assert ref.shape[1] == 7
    assert frac_within(grad, ref[:, :7], 1e-4, 2e-5) >= 0.99
    assert frac_within(grad, ref[:, :7], 2e-3, 4e-4) == 1.0


def test_fk_rollout_one_step_map_teacher_forced():
    c = load_npz("fkcase_planar7")
    W, b = load_weights("planar7")
    N, H, nk, dt = int(c["N"]), int(c["H"]), int(c["nk"]), float(c["dt"])
    mu, sg, al = (full_policy(c, k, N) for k in ("mu_tmp", "sigma_tmp", "alpha_tmp"))
    prm = orc.RolloutParams(dt=dt, dt_H=1, n_closest_obs=1, dst_thr=float(c["dst_thr"]), ignored_links=[],
                            fk_dh_params=c["dh_params"])
    for t in range(H):
        q = c["all_traj"][:, t, :]
        o = orc.rollout(orc.Net(W, b), q, c["qf"], c["obs"], mu, sg, al, nk, prm, N)
        assert frac_within(o.closest_dist_all[:, 0], c["closest_dist_all"][:, t], 1e-5, 2e-6) >= 0.99
        assert frac_within(o.dot_products[:, 0], c["dot_products"][:, t], 1e-4, 1e-5) >= 0.95
        assert frac_within(o.kernel_activations[:, 0], c["kernel_activations"][:, t], 1e-4, 1e-5) >= 0.95
        if t + 1 < H:
            assert frac_within(q + dt * o.qdot, c["all_traj"][:, t + 1, :], 1e-4, 1e-5) >= 0.99
