"""Parity AT THE SHAPES bench.py TIMES (BASELINE.json configs[0..2]).

Goldens: tests/golden/case_c1_planar2.npz, bench_c2_planar7.npz, bench_c3_franka2064.npz -- outputs of the unmodified
reference on bench.py's own problem definitions (tests/golden/make_golden_bench.py).
  * CPU (`-m "not gpu"`): the oracle restatement against those goldens (teacher-forced one-step map + cost + update),
    so the oracle is pinned at the benchmarked shapes too;
  * GPU (`-m gpu`): the CUDA path at the FULL bench configuration (e.g. Franka shelf: 4096 samples x 50 steps x 2064
    spheres) -- the golden samples against the reference, a fixed 64-sample subset teacher-forced through the oracle
    at every step, cost and policy update of the whole batch against the oracle, and the exactness counters of the
    prefilter (no capacity retry, no fp32 fallback, no dropped row).
Tolerances as in tests/test_gpu_parity.py (north_star: distances / gradients / velocities 1e-5 relative, trajectories /
cost / policy 1e-4); the absolute terms of the tensor-core split mode are written at `ATOL`."""
import pytest
import torch

import bench
from oracle import mppi_oracle as orc
from tests.golden_util import frac_within, full_policy, load_npz, load_weights

# absolute terms per scoring arithmetic (metres / unit vectors): see tests/test_gpu_parity.py for their derivation
ATOL = {"ffma": dict(dist=2e-6, dot=1e-5), "tc_split": dict(dist=5e-6, dot=3e-5)}


def check(a, b, rtol, atol, name, min_frac=0.99, loose=20, kink_samples=0):
    """>= min_frac within tolerance, nothing beyond loose x; `kink_samples` leading-index samples may be set aside for
    quantities that follow the gradient direction (ReLU-kink flips)."""
    a, b = a.detach().cpu(), b.detach().cpu()
    assert a.shape == b.shape, f"{name}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    if kink_samples and a.shape[0] > 1:
        viol = ((a.double() - b.double()).abs() / (atol + rtol * b.double().abs())).reshape(a.shape[0], -1)
        worst = viol.max(dim=1)[0]
        drop = torch.argsort(worst, descending=True)[:kink_samples]
        keep = torch.ones(a.shape[0], dtype=torch.bool)
        keep[drop[worst[drop] > 1.0]] = False
        a, b = a[keep], b[keep]
    f = frac_within(a, b, rtol, atol)
    min_frac = min(min_frac, 1.0 - 1.0 / max(a.numel(), 1)) if min_frac < 1.0 else 1.0
    assert f >= min_frac, f"{name}: only {f:.4f} within rtol={rtol} (max abs diff {(a - b).abs().max():.3e})"
    assert frac_within(a, b, loose * rtol, loose * atol) == 1.0, \
        f"{name}: outliers beyond {loose}x tolerance (max abs diff {(a - b).abs().max():.3e})"


def _oracle_params(c, H=1):
    return orc.RolloutParams(dt=float(c["dt"]), dt_H=H, n_closest_obs=int(c["K"]), dst_thr=float(c["dst_thr"]),
                             ignored_links=c["ignored_links"].tolist(), p=float(c["p"]), with_basis=False)


def _kept_policy(c):
    keep = c["keep"]
    n_all = c["mu_tmp"].shape[0]
    return tuple(full_policy(c, k, n_all)[keep] for k in ("mu_tmp", "sigma_tmp", "alpha_tmp"))


# ------------------------------------------------------------------------------------------ CPU: oracle vs reference
@pytest.mark.parametrize("tag,stride", [("bench_c2_planar7", 1), ("bench_c3_franka2064", 7)])
def test_oracle_one_step_map_at_bench_shapes(tag, stride):
    """Teacher-forced: the reference's own states of the kept samples in, one oracle step out (every step for C2,
    every 7th for the 2064-sphere shelf to keep the CPU suite short)."""
    c = load_npz(tag)
    W, b = load_weights(c["net"])
    net = orc.Net(W, b)
    H, nk, dt = int(c["H"]), int(c["nk"]), float(c["dt"])
    mu, sg, al = _kept_policy(c)
    n = mu.shape[0]
    prm = _oracle_params(c)
    for t in range(0, H, stride):
        q = c["all_traj"][:, t, :]
        o = orc.rollout(net, q, c["qf"], c["obs"], mu, sg, al, nk, prm, n)
        check(o.closest_dist_all[:, 0], c["closest_dist_all"][:, t], 1e-5, 2e-6, f"dist[{t}]")
        check(o.dot_products[:, 0], c["dot_products"][:, t], 1e-5, 1e-5, f"dot[{t}]", kink_samples=1)
        check(o.kernel_activations[:, 0], c["kernel_activations"][:, t], 1e-4, 1e-5, f"act[{t}]", kink_samples=1)
        check(o.kernel_val_all[:, 0, :nk], c["kernel_val_all"][:, t, :nk], 1e-4, 1e-6, f"kval[{t}]")
        if t + 1 < H:
            check(q + dt * o.qdot, c["all_traj"][:, t + 1, :], 1e-5, 1e-5, f"traj[{t + 1}]", kink_samples=1)


@pytest.mark.parametrize("tag", ["bench_c2_planar7", "bench_c3_franka2064"])
def test_oracle_cost_and_qdot_at_bench_shapes(tag):
    c = load_npz(tag)
    keep = c["keep"]
    cost = orc.evaluate_costs(c["all_traj"], c["closest_dist_all"], c["qf"], c["dh_params"], c["q_min"], c["q_max"])
    check(cost, c["cost"][keep], 1e-5, 1e-4, "cost")
    # get_qdot (MPPI.py:319-329) needs the velocities of ALL samples; C3 stores every sample
    if keep.numel() == int(c["N"]):
        best = c["qdot"][torch.argmin(c["cost"])]
        beta = c["cost"].mean() / 50
        w = torch.exp(-c["cost"] / beta)
        w = w / w.sum()
        torch.testing.assert_close(best, c["qdot_best"], rtol=0, atol=0)
        torch.testing.assert_close((w[:, None] * c["qdot"]).sum(0), c["qdot_weighted"], rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------ GPU: the benched shapes
@pytest.fixture(scope="module")
def factory():
    from tests import mppi_factory
    return mppi_factory


def _bench_mppi(workload, score, golden=None, device="cuda"):
    """The MPPI object exactly as bench.py builds it (bench.build_mppi), with the golden's policy draw."""
    import argparse
    p = bench.problem(workload)
    if golden is not None and "q0" in golden:
        p["q0"] = golden["q0"]
    args = argparse.Namespace(pass1="auto", score=score, per_step=False)
    seed = int(golden["policy_seed"]) if golden is not None else 100
    mppi, pol, _ = bench.build_mppi(p, p["N"], p["H"], torch.device(device), args, seed, 1)
    return p, mppi, pol


def _teacher_force_against_oracle(p, mppi, pol, out, subset, atol, steps):
    traj, dist, kv, dots, acts = out
    W, b = load_weights(p["net"])
    net = orc.Net(W, b)
    nk, dt = p["nk"], p["dt"]
    prm = orc.RolloutParams(dt=dt, dt_H=1, n_closest_obs=p["K"], dst_thr=p["dst_thr"], ignored_links=list(p["ignored"]),
                            p=2.0, with_basis=False)
    mu, sg, al = (x[subset] for x in pol[3:])
    n = len(subset)
    kink = max(1, n // 100)
    for t in steps:
        q = traj[subset, t, :].cpu()
        o = orc.rollout(net, q, p["qf"], p["obs"], mu, sg, al, nk, prm, n)
        check(dist[subset, t], o.closest_dist_all[:, 0], 1e-5, atol["dist"], f"dist[{t}]")
        check(dots[subset, t], o.dot_products[:, 0], 1e-5, atol["dot"], f"dot[{t}]", min_frac=0.9, kink_samples=kink)
        check(acts[subset, t], o.kernel_activations[:, 0], 1e-4, 1e-5, f"act[{t}]", kink_samples=kink)
        check(kv[subset, t, :], o.kernel_val_all[:, 0, :nk], 1e-4, 1e-6, f"kval[{t}]")
        if t + 1 < traj.shape[1]:
            check(traj[subset, t + 1, :], q + dt * o.qdot, 1e-5, 1e-5, f"traj[{t + 1}]", kink_samples=kink)


def _cost_and_update_against_oracle(p, mppi, pol, out):
    traj, dist, kv, dots, acts = out
    P = mppi.Policy
    mu0, sg0, al0 = P.mu_c.clone().cpu(), P.sigma_c.clone().cpu(), P.alpha_c.clone().cpu()
    cost = mppi.get_cost()
    ocost = orc.evaluate_costs(traj.cpu(), dist.cpu(), p["qf"], p["dh"], p["qlim"][0], p["qlim"][1])
    check(cost, ocost, 1e-5, 1e-4, "cost")
    _, n_upd = mppi.shift_policy_means()
    mu1, sg1, al1, on, _ = orc.policy_update(cost.cpu(), mppi.kernel_val_all.cpu(), acts.cpu(), pol[3], pol[4], pol[5],
                                             mu0, sg0, al0, p["nk"], p["ker_thr"])
    assert int(n_upd) == on
    check(P.mu_c, mu1, 1e-4, 1e-6, "mu_c")
    check(P.sigma_c, sg1, 1e-4, 1e-6, "sigma_c")
    check(P.alpha_c, al1, 1e-4, 1e-6, "alpha_c")
    return cost


@pytest.mark.gpu
@pytest.mark.parametrize("score", ["tc_split", "ffma"])
def test_c3_franka_shelf_2064_at_the_benched_shape(score):
    """BENCH default: 4096 samples x 50 steps x 2064 spheres, K = 5 (what BENCH_rNN.json times)."""
    g = load_npz("bench_c3_franka2064")
    p, mppi, pol = _bench_mppi("franka_shelf_2064", score, g)
    assert (p["N"], p["H"], p["obs"].shape[0]) == (4096, 50, 2064)
    n_g = int(g["N"])
    assert torch.equal(pol[5][:n_g, :p["nk"]], g["alpha_tmp"]), "bench.seeded_policy no longer matches the golden draw"
    out = mppi.propagate()
    traj, dist, kv, dots, acts = out
    a = ATOL[score]
    # (1) the reference's own samples: free-running is comparable while no sample has met a kink -- the first steps
    kink = 1
    check(dist[:n_g, :3], g["closest_dist_all"][:, :3], 1e-4, 10 * a["dist"], "golden closest_dist (3 steps)",
          kink_samples=kink)
    check(traj[:n_g, :3], g["all_traj"][:, :3], 1e-4, 1e-5, "golden all_traj (3 steps)", kink_samples=kink)
    check(mppi.qdot[:n_g], g["qdot"], 1e-5, 1e-5, "golden qdot", kink_samples=kink)
    # (2) ... and teacher-forced over the whole horizon: golden states in, one GPU step out (N = 64, M = 2064, H = 1)
    m1, _, _ = bench.build_mppi(dict(p, N=n_g, H=1), n_g, 1, torch.device("cuda"),
                                __import__("argparse").Namespace(pass1="auto", score=score, per_step=False), 100, 1)
    bench.load_policy(m1, p, tuple(pol[:3]) + tuple(x[:n_g] for x in pol[3:]), torch.device("cuda"))
    for t in range(0, 50):
        m1.q_cur = g["all_traj"][:, t, :].cuda()
        _, d1, k1, o1, a1 = m1.propagate()
        check(d1[:, 0], g["closest_dist_all"][:, t], 1e-5, a["dist"], f"tf dist[{t}]")
        check(o1[:, 0], g["dot_products"][:, t], 1e-5, a["dot"], f"tf dot[{t}]", min_frac=0.9, kink_samples=kink)
        check(a1[:, 0], g["kernel_activations"][:, t], 1e-4, 1e-5, f"tf act[{t}]", kink_samples=kink)
        check(k1[:, 0, :], g["kernel_val_all"][:, t, :p["nk"]], 1e-4, 1e-6, f"tf kval[{t}]")
        if t + 1 < 50:
            check(g["all_traj"][:, t, :] + p["dt"] * m1.qdot.cpu(), g["all_traj"][:, t + 1, :], 1e-5, 1e-5,
                  f"tf traj[{t + 1}]", kink_samples=kink)
    assert m1.pass1_stats()["mode"] == 1
    # (3) a fixed 64-sample subset of the 4096 through the ORACLE at every 5th step of the GPU's own rollout
    subset = torch.arange(0, 4096, 64)
    _teacher_force_against_oracle(p, mppi, pol, out, subset, a, range(0, 50, 5))
    # (4) exactness of the prefilter at this shape: nothing repeated, nothing dropped
    st, xs, ss = mppi.pass1_stats(), mppi.exactness_stats(), mppi.score_stats()
    print(f"franka_shelf_2064 [{score}]: {st}, {xs}, {ss}")
    assert st["mode"] == 1, "the benched shape must run the tensor-core prefilter"
    assert xs["capacity_retries"] == 0 and xs["exact_fallbacks"] == 0 and ss["dropped_rows"] == 0
    assert 5.0 <= st["rescored_pairs"] / (4096 * 50) <= 16.0
    # (5) cost and policy update of the whole batch
    _cost_and_update_against_oracle(p, mppi, pol, out)


@pytest.mark.gpu
@pytest.mark.parametrize("score", ["tc_split", "ffma"])
def test_c2_planar7_at_the_benched_shape(score):
    """configs[1]: 1000 samples x 30 steps, 4 spheres, K = 1 -- the whole-horizon single-launch path."""
    g = load_npz("bench_c2_planar7")
    p, mppi, pol = _bench_mppi("planar7", score, g)
    assert (p["N"], p["H"]) == (1000, 30)
    keep = g["keep"]
    assert torch.equal(pol[5][keep][:, :p["nk"]], g["alpha_tmp"][keep])
    out = mppi.propagate()
    traj, dist, kv, dots, acts = out
    a = ATOL[score]
    kink = max(1, len(keep) // 100)
    check(dist[keep, :2], g["closest_dist_all"][:, :2], 1e-4, 10 * a["dist"], "golden closest_dist (2 steps)",
          kink_samples=kink)
    check(traj[keep, :2], g["all_traj"][:, :2], 1e-4, 1e-5, "golden all_traj (2 steps)", kink_samples=kink)
    check(mppi.qdot[keep], g["qdot"], 1e-5, 1e-5, "golden qdot", kink_samples=kink)
    # teacher-forced over the whole horizon on the reference's states (H = 1 objects take per-sample start states)
    import argparse
    nk_ = len(keep)
    m1, _, _ = bench.build_mppi(dict(p, N=nk_, H=1), nk_, 1, torch.device("cuda"),
                                argparse.Namespace(pass1="auto", score=score, per_step=False), 100, 1)
    bench.load_policy(m1, p, tuple(pol[:3]) + tuple(x[keep] for x in pol[3:]), torch.device("cuda"))
    for t in range(30):
        m1.q_cur = g["all_traj"][:, t, :].cuda()
        _, d1, k1, o1, a1 = m1.propagate()
        check(d1[:, 0], g["closest_dist_all"][:, t], 1e-5, a["dist"], f"tf dist[{t}]")
        check(o1[:, 0], g["dot_products"][:, t], 1e-5, a["dot"], f"tf dot[{t}]", min_frac=0.9, kink_samples=kink)
        check(a1[:, 0], g["kernel_activations"][:, t], 1e-4, 1e-5, f"tf act[{t}]", kink_samples=kink)
        check(k1[:, 0, :], g["kernel_val_all"][:, t, :p["nk"]], 1e-4, 1e-6, f"tf kval[{t}]")
        if t + 1 < 30:
            check(g["all_traj"][:, t, :] + p["dt"] * m1.qdot.cpu(), g["all_traj"][:, t + 1, :], 1e-5, 1e-5,
                  f"tf traj[{t + 1}]", kink_samples=kink)
    # the GPU's own 1000 x 30 rollout through the oracle, all samples, every 3rd step
    _teacher_force_against_oracle(p, mppi, pol, out, torch.arange(1000), a, range(0, 30, 3))
    _cost_and_update_against_oracle(p, mppi, pol, out)
    # cost / update kernels on the reference's own trajectories (kept samples)
    cost = mppi.Cost.evaluate_costs(g["all_traj"].cuda(), g["closest_dist_all"].cuda())
    check(cost, g["cost"][keep], 1e-5, 1e-4, "cost on reference trajectories")


@pytest.mark.gpu
@pytest.mark.parametrize("score", ["tc_split", "ffma"])
def test_c1_planar2_at_the_benched_shape(score):
    """configs[0] with a non-degenerate start (collisions, activations, all 10 kernels updated): 100 x 10."""
    g = load_npz("case_c1_planar2")
    assert int((g["closest_dist_all"] < 0).sum()) > 0 and int((g["kernel_activations"] > 0).sum()) > 0
    assert int(g["n_updated"]) == 10
    p, mppi, pol = _bench_mppi("planar2", score, g)
    out = mppi.propagate()
    a = ATOL[score]
    _teacher_force_against_oracle(p, mppi, pol, out, torch.arange(100), a, range(10))
    _cost_and_update_against_oracle(p, mppi, pol, out)
    # the update on the reference's own rollout outputs equals the reference's update
    m2 = _bench_mppi("planar2", score, g)[1]
    m2.cur_cost = g["cost"].cuda()
    kvf = torch.zeros(100, 10, 50)
    kvf[:, :, :10] = g["kernel_val_all"]
    m2.kernel_val_all, m2.kernel_activations = kvf.cuda(), g["kernel_activations"].cuda()
    _, n_upd = m2.shift_policy_means()
    assert int(n_upd) == 10
    check(m2.Policy.mu_c, g["mu_c1"], 1e-5, 1e-6, "mu_c")
    check(m2.Policy.alpha_c, g["alpha_c1"], 1e-5, 1e-6, "alpha_c")
