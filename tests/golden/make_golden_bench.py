"""Golden fixtures AT THE BENCHMARKED SHAPES, produced by executing the unmodified reference (build container only).

    python tests/golden/make_golden_bench.py

bench.py times BASELINE.json's configurations; the fixtures of make_golden.py are smaller than those (64 x 12 ...).
This script drives the reference through oracle/ref_harness.py on bench.py's own problem definitions
(bench.problem / bench.seeded_policy, so the GPU tests rebuild exactly these inputs) and stores
  case_c1_planar2.npz      configs[0]: planar 2-DoF, 100 samples x 10 steps, started next to the obstacle so that
                           collisions, kernel activations and a non-empty policy update all occur
  bench_c2_planar7.npz     configs[1]: planar 7-DoF, 1000 samples x 30 steps (outputs of every 8th sample; cost and
                           the updated policy of all 1000)
  bench_c3_franka2064.npz  configs[2]: Franka shelf, 2064 spheres, 50 steps: the first 64 samples of the bench's
                           4096-sample draw (seed 100), every output of every step
  edges_*.npz              get_qdot('best' | 'weighted') and update_kernel_normal_bases (MPPI.py:284-304,319-329)
Also writes nothing else; never imported by the product, the GPU tests or bench.py.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

torch.set_num_threads(8)


def run(tag, workload, N, H, policy_seed, n_policy, keep_every=1, q0=None, nk=None, store_basis=False):
    """The first N samples of bench.seeded_policy(p, n_policy, policy_seed) through the reference."""
    p = bench.problem(workload)
    if q0 is not None:
        p["q0"] = q0
    if nk is not None:
        p["nk"] = nk
    pol = bench.seeded_policy(p, n_policy, policy_seed)
    pol = pol[:3] + tuple(t[:N].clone() for t in pol[3:])
    it = rh.ReferenceIteration(p, N, H, seed=0, policy=pol)
    m = it.mppi
    P = m.Policy
    mu_c0, sigma_c0, alpha_c0 = P.mu_c.clone(), P.sigma_c.clone(), P.alpha_c.clone()
    with contextlib.redirect_stdout(io.StringIO()):
        (traj, cdist, kval, dots, acts), cost, (_, n_upd) = it.step()
        qd_best = m.get_qdot('best').clone()
        qd_weighted = m.get_qdot('weighted').clone()
    keep = torch.arange(0, N, keep_every)
    nkk = max(p["nk"], 1)
    out = dict(
        net=np.array(p["net"]), workload=np.array(workload), q0=p["q0"], qf=p["qf"], dh_params=p["dh"], dh_a=p["dh_a"],
        obs=p["obs"], dt=np.float64(p["dt"]), H=np.int64(H), N=np.int64(N), K=np.int64(p["K"]), nk=np.int64(p["nk"]),
        dst_thr=np.float64(p["dst_thr"]), ker_thr=np.float64(p["ker_thr"]), p=np.float64(2.0),
        ignored_links=np.array(list(p["ignored"]), dtype=np.int64), q_min=p["qlim"][0], q_max=p["qlim"][1],
        q_cur=p["q0"], policy_seed=np.int64(policy_seed), n_policy=np.int64(n_policy), keep=keep,
        mu_c0=mu_c0, sigma_c0=sigma_c0, alpha_c0=alpha_c0,
        mu_tmp=pol[3][:, :nkk], sigma_tmp=pol[4][:, :nkk], alpha_tmp=pol[5][:, :nkk],
        all_traj=traj[keep], closest_dist_all=cdist[keep], kernel_val_all=m.kernel_val_all[keep][:, :, :nkk],
        dot_products=dots[keep], kernel_activations=acts[keep], qdot=m.qdot[keep], cost=cost,
        mu_c1=P.mu_c.clone(), sigma_c1=P.sigma_c.clone(), alpha_c1=P.alpha_c.clone(), n_updated=np.int64(int(n_upd)),
        qdot_best=qd_best, qdot_weighted=qd_weighted,
    )
    if store_basis:
        out["norm_basis"] = m.norm_basis[keep]
    out = {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
    np.savez_compressed(os.path.join(HERE, f"{tag}.npz"), **out)
    print(f"{tag}: N={N} H={H} M={p['obs'].shape[0]} K={p['K']} nk={p['nk']} collisions={int((cdist < 0).sum())} "
          f"act>0: {int((acts > 0).sum())} n_updated={int(n_upd)} cost[{cost.min():.2f},{cost.max():.2f}] "
          f"size={os.path.getsize(os.path.join(HERE, tag + '.npz')) / 1024:.0f} KiB")
    return it, p


def run_kernel_bases(tag, it, p):
    """update_kernel_normal_bases at the kernel centres of a finished iteration (MPPI.py:284-304)."""
    m = it.mppi
    nk = m.Policy.n_kernels
    with contextlib.redirect_stdout(io.StringIO()):
        m.update_kernel_normal_bases()
    out = dict(mu_c=m.Policy.mu_c[:nk].clone(), bases=m.Policy.kernel_obstacle_bases[:nk].clone(), obs=p["obs"],
               K=np.int64(p["K"]), ignored_links=np.array(list(p["ignored"]), dtype=np.int64), net=np.array(p["net"]))
    out = {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
    np.savez_compressed(os.path.join(HERE, f"{tag}.npz"), **out)
    print(f"{tag}: nk={nk} bases {tuple(out['bases'].shape)}")


def main():
    only = sys.argv[1:]
    if not only or "c1" in only:
        # next to the first obstacle: link 1 tip at (3 cos q1, 3 sin q1) ... chosen so the run is not degenerate
        it, p = run("case_c1_planar2", "planar2", 100, 10, 100, 100, q0=torch.tensor([-0.6, 0.9]), store_basis=True)
        run_kernel_bases("edges_planar2_bases", it, p)
    if not only or "c2" in only:
        it, p = run("bench_c2_planar7", "planar7", 1000, 30, 100, 1000, keep_every=8)
        run_kernel_bases("edges_planar7_bases", it, p)
    if not only or "c3" in only:
        it, p = run("bench_c3_franka2064", "franka_shelf_2064", 64, 50, 100, 4096)
        run_kernel_bases("edges_franka2064_bases", it, p)


if __name__ == "__main__":
    main()
