"""Golden fixtures for the true-distance (forward-kinematics) provider, produced by EXECUTING THE UNMODIFIED
REFERENCE: MPPI.distance_repulsion_fk (ds_mppi/functions/MPPI.py:306-313 -> fk_num.numeric_fk_model_vec,
fk_num.dist_tens, fk_sym_gen.lambda_rep_vec) on the planar 7-DoF arm, and one MPPI iteration with that provider
switched in the way the reference's own comment at MPPI.py:115 does (the object's distance_repulsion_nn attribute
is pointed at its distance_repulsion_fk method; no reference source or arithmetic is touched).

    python tests/golden/make_golden_fk.py       # build container only (needs /root/reference)

Writes tests/golden/fkdist_*.npz and tests/golden/fkcase_planar7.npz."""
import contextlib
import io
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

torch.set_num_threads(8)
ns = rh.load_reference()
with contextlib.redirect_stdout(io.StringIO()):
    NET = rh.make_ref_net(ns, 7, 7, "7dof_sdf_256x5_mesh.pt", torch)


def planar_dh(dof, L):
    dh_a = torch.zeros(dof + 1)
    dh_a[1:] = L
    return torch.vstack((dh_a * 0, dh_a * 0, dh_a, dh_a * 0)).T, dh_a


def make(obs, N, H, dt, q0, qf):
    dh, dh_a = planar_dh(7, 1)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ns.MPPI(q0, qf, dh, obs, dt, H, N, [ns.LinDS(qf), ns.LinDS(q0)], dh_a, NET, 1)
    return m, dh, dh_a


def run_dist(tag, obs, n, qrange, seed):
    q0 = torch.zeros(7)
    m, dh, dh_a = make(obs, n, 1, 0.1, q0, q0 + 1)
    g = torch.Generator().manual_seed(seed)
    q = (torch.rand(n, 7, generator=g) * 2 - 1) * qrange
    dist, rep = m.distance_repulsion_fk(q.clone())
    links, _ = ns.mod.numeric_fk_model_vec(q, dh, 10)
    _, io_, il, ip = ns.mod.get_mindist(links, obs)
    idx = torch.cat((io_, il, ip), 1)
    np.savez_compressed(os.path.join(HERE, f"fkdist_{tag}.npz"), q=q.numpy(), obs=obs.numpy(), dh_params=dh.numpy(),
                        distance=dist.numpy(), rep_vec=rep.numpy(), idx=idx.numpy().astype(np.int64))
    print(f"fkdist_{tag}: n={n} M={obs.shape[0]} dist[{dist.min():.3f},{dist.max():.3f}] links used "
          f"{sorted(set(il.flatten().tolist()))}")


def run_case(tag, obs, N, H, dt, nk, seed):
    pi = math.pi
    q0 = torch.zeros(7); q0[0] = pi / 2
    qf = torch.zeros(7); qf[0] = -pi / 2
    torch.manual_seed(seed)
    m, dh, dh_a = make(obs, N, H, dt, q0, qf)
    m.distance_repulsion_nn = lambda q, aot=False: m.distance_repulsion_fk(q)      # MPPI.py:113 <-> :115
    m.Policy.sigma_c_nominal, m.Policy.alpha_s = 0.5, 0.75
    m.dst_thr, m.ker_thr, m.ignored_links = 0.25, 1e-3, []
    g = torch.Generator().manual_seed(seed + 1)
    P = m.Policy
    P.n_kernels = nk
    P.mu_c[:nk] = q0 + 0.15 * torch.randn(nk, 7, generator=g)
    P.sigma_c[:nk] = 0.5
    P.alpha_c[:nk] = torch.randn(nk, 7, generator=g)
    mu_c0, sigma_c0, alpha_c0 = P.mu_c.clone(), P.sigma_c.clone(), P.alpha_c.clone()
    torch.manual_seed(seed + 2)
    P.sample_policy()
    mu_tmp, sigma_tmp, alpha_tmp = P.mu_tmp.clone(), P.sigma_tmp.clone(), P.alpha_tmp.clone()
    with contextlib.redirect_stdout(io.StringIO()):
        all_traj, cdist, kval, dots, acts = m.propagate()
        cost = m.get_cost()
    out = dict(net=np.array("planar7"), q0=q0, qf=qf, dh_params=dh, dh_a=dh_a, obs=obs, dt=np.float64(dt),
               H=np.int64(H), N=np.int64(N), K=np.int64(1), nk=np.int64(nk), dst_thr=np.float64(0.25),
               ker_thr=np.float64(1e-3), p=np.float64(2), ignored_links=np.zeros(0, dtype=np.int64),
               q_min=m.Cost.q_min, q_max=m.Cost.q_max, q_cur=m.q_cur, mu_c0=mu_c0, sigma_c0=sigma_c0,
               alpha_c0=alpha_c0, mu_tmp=mu_tmp[:, :nk], sigma_tmp=sigma_tmp[:, :nk], alpha_tmp=alpha_tmp[:, :nk],
               all_traj=all_traj, closest_dist_all=cdist, kernel_val_all=m.kernel_val_all[:, :, :nk],
               dot_products=dots, kernel_activations=acts, qdot=m.qdot, norm_basis=m.norm_basis, cost=cost)
    out = {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
    np.savez_compressed(os.path.join(HERE, f"fkcase_{tag}.npz"), **out)
    print(f"fkcase_{tag}: N={N} H={H} collisions={int((cdist < 0).sum())} act>0 {int((acts > 0).sum())} "
          f"cost[{cost.min():.2f},{cost.max():.2f}]")


def main():
    obs7 = torch.tensor([[6, 2, 0, .5], [4., -1, 0, .5], [5, 0, 0, .5], [6, 6, 6, .1]])
    run_dist("planar7", obs7, 96, 1.5, seed=31)
    g = torch.Generator().manual_seed(32)
    many = torch.cat(((torch.rand(48, 2, generator=g) * 2 - 1) * 7, torch.zeros(48, 1),
                      0.1 + 0.4 * torch.rand(48, 1, generator=g)), 1)
    run_dist("planar7_many", many, 64, 2.5, seed=33)         # M >= 32: the warp-per-sample kernel
    run_case("planar7", obs7, N=32, H=8, dt=0.3, nk=4, seed=34)


if __name__ == "__main__":
    main()
