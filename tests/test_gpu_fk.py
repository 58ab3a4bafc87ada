"""GPU parity tests of the true-distance (FK) provider (run with -m gpu on a B200): MPPI.distance_repulsion_fk
and a rollout with `distance_provider = 'fk'` through the drop-in object / C ABI, against the committed outputs
of the UNMODIFIED reference (tests/golden/fkdist_*.npz, fkcase_planar7.npz) and the oracle on fresh inputs."""
import math

import pytest
import torch

from oracle import mppi_oracle as orc
from tests.golden_util import frac_within, full_policy, load_npz

pytestmark = pytest.mark.gpu


def check(a, b, rtol, atol, name, min_frac=0.99, loose=20):
    a, b = a.detach().cpu(), b.detach().cpu()
    assert a.shape == b.shape, f"{name}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    f = frac_within(a, b, rtol, atol)
    min_frac = min(min_frac, 1.0 - 1.0 / max(a.numel(), 1)) if min_frac < 1.0 else 1.0
    assert f >= min_frac, f"{name}: only {f:.4f} within rtol={rtol} (max abs diff {(a - b).abs().max():.3e})"
    assert frac_within(a, b, loose * rtol, loose * atol) == 1.0, \
        f"{name}: outliers beyond {loose}x tolerance (max abs diff {(a - b).abs().max():.3e})"


def planar7_mppi(obs, N, H, dt, q0, qf, device="cpu"):
    from tests.mppi_factory import make_net
    from optimalmodulationds_b200 import MPPI, LinDS
    dev = torch.device(device)
    dh_a = torch.zeros(8); dh_a[1:] = 1
    dh = torch.vstack((dh_a * 0, dh_a * 0, dh_a, dh_a * 0)).T.contiguous()
    t = lambda x: x.to(dev)  # noqa: E731
    m = MPPI(t(q0), t(qf), t(dh), t(obs), dt, H, N, [LinDS(t(qf)), LinDS(t(q0))], t(dh_a), make_net("planar7"), 1)
    return m, dh


@pytest.mark.parametrize("tag", ["planar7", "planar7_many"])
def test_distance_repulsion_fk_vs_reference(tag):
    c = load_npz(f"fkdist_{tag}")
    n = c["q"].shape[0]
    m, _ = planar7_mppi(c["obs"], n, 1, 0.1, torch.zeros(7), torch.ones(7))
    dist, rep, idx = m.distance_repulsion_fk(c["q"], return_indices=True)
    assert rep.shape == (n, 7) and m.nn_grad.shape == (n, 7)
    assert torch.equal(idx.long(), c["idx"])                   # same (obstacle, link, point) as dist_tens
    check(dist, c["distance"], 1e-5, 2e-6, "distance", min_frac=1.0)
    check(rep, c["rep_vec"], 1e-4, 2e-5, "rep_vec")


def test_fk_rollout_teacher_forced_and_free_running_vs_reference():
    c = load_npz("fkcase_planar7")
    N, H, nk, dt = int(c["N"]), int(c["H"]), int(c["nk"]), float(c["dt"])
    for H_run in (1, H):
        m, _ = planar7_mppi(c["obs"], N, H_run, dt, c["q0"], c["qf"])
        m.distance_provider = "fk"
        m.dst_thr, m.ker_thr, m.ignored_links = float(c["dst_thr"]), float(c["ker_thr"]), []
        P = m.Policy
        P.n_kernels = nk
        P.mu_tmp.copy_(full_policy(c, "mu_tmp", N)); P.sigma_tmp.copy_(full_policy(c, "sigma_tmp", N))
        P.alpha_tmp.copy_(full_policy(c, "alpha_tmp", N))
        if H_run == 1:
            for t in range(H):
                q = c["all_traj"][:, t, :]
                m.q_cur = q
                traj, dist, kv, dots, acts = m.propagate()
                check(dist[:, 0], c["closest_dist_all"][:, t], 1e-5, 2e-6, f"dist[{t}]")
                check(dots[:, 0], c["dot_products"][:, t], 1e-4, 1e-5, f"dot[{t}]", min_frac=0.9)
                check(acts[:, 0], c["kernel_activations"][:, t], 1e-4, 1e-5, f"act[{t}]", min_frac=0.95)
                check(kv[:, 0], c["kernel_val_all"][:, t, :nk], 1e-4, 1e-6, f"kval[{t}]")
                if t + 1 < H:
                    check(q + dt * m.qdot, c["all_traj"][:, t + 1, :], 1e-4, 1e-5, f"traj[{t + 1}]")
        else:
            m.q_cur = c["q_cur"]
            traj, dist, kv, dots, acts = m.propagate()
            check(dist[:, :2], c["closest_dist_all"][:, :2], 1e-4, 1e-5, "closest_dist_all[:2]")
            check(traj[:, :3], c["all_traj"][:, :3], 1e-4, 1e-5, "all_traj[:3]")
            check(m.qdot, c["qdot"], 1e-4, 1e-5, "qdot")


def test_fk_provider_on_franka_chain_matches_oracle_and_finite_differences():
    """Beyond the reference ("not implemented for Franka", MPPI.py:115): the geometric-Jacobian gradient works for
    any modified-DH arm.  Checked against the oracle and against central differences of the oracle's distance."""
    from tests.mppi_factory import make_net
    from optimalmodulationds_b200 import MPPI, LinDS
    pi = math.pi
    dh_a = torch.tensor([0, 0, 0, 0.0825, -0.0825, 0, 0.088, 0])
    dh_d = torch.tensor([0.333, 0, 0.316, 0, 0.384, 0, 0, 0.107])
    dh_alpha = torch.tensor([0, -pi / 2, pi / 2, pi / 2, -pi / 2, pi / 2, pi / 2, 0])
    dh = torch.vstack((dh_d, dh_a * 0, dh_a, dh_alpha)).T.contiguous()
    g = torch.Generator().manual_seed(3)
    obs = torch.cat(((torch.rand(70, 3, generator=g) - 0.5) * 1.6, 0.03 + 0.05 * torch.rand(70, 1, generator=g)), 1)
    q = (torch.rand(200, 7, generator=g) * 2 - 1) * 1.5
    q0 = torch.zeros(7)
    m = MPPI(q0, q0 + 1, dh, obs, 0.1, 1, 200, [LinDS(q0 + 1)], dh_a, make_net("franka"), 5)
    dist, rep, idx = m.distance_repulsion_fk(q, return_indices=True)
    od, og, oi = orc.distance_repulsion_fk(q, obs, dh, 10, return_idx=True)
    same = (idx.long() == oi).all(1)
    assert same.float().mean() > 0.98                          # near-ties between neighbouring points may flip
    check(dist, od, 1e-5, 2e-6, "distance")
    check(rep[same, :7], og[same], 1e-4, 2e-5, "gradient")
    eps = 1e-3
    num = torch.zeros(200, 7, dtype=torch.float64)
    for j in range(7):
        e = torch.zeros(7, dtype=torch.float64); e[j] = eps
        num[:, j] = (orc.distance_repulsion_fk((q.double() + e).float(), obs, dh)[0].double()
                     - orc.distance_repulsion_fk((q.double() - e).float(), obs, dh)[0].double()) / (2 * eps)
    ok = ((rep[:, :7].double() - num).abs() <= 5e-3 + 5e-3 * num.abs()).all(1)
    assert ok.float().mean() > 0.95                            # the min over points has kinks: most rows are smooth
