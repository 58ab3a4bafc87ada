"""Generate the committed golden fixtures by EXECUTING THE UNMODIFIED REFERENCE in this container.

    python tests/golden/make_golden.py          # needs /root/reference (read-only) -- build container only

Writes
  tests/golden/weights/{planar2,planar7,franka}.npz   fp32 weights/biases of the shipped checkpoints
                                                      (mlp_learn/models/*.pt, model_state_dict only)
  tests/golden/kat.npz                                known-answer forward/VJP vectors
  tests/golden/case_*.npz                             seeded MPPI iterations: inputs + every output of
                                                      propagate / get_cost / shift_policy_means
  tests/golden/distgrad_*.npz                         distance_repulsion_nn on random joint batches
The reference is driven through its public API exactly as its scripts do (standalonePlanar2d.py:109-131,
standalonePlanar7d.py:95-107, frankaPlanner.py:83-90) with the two import shims documented in
ref_harness.py.  Nothing here is imported by the product or by the GPU-side tests.
"""
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

NETS = {
    "planar2": dict(dof=2, out=2, fname="2dof_sdf_256x5_mesh.pt"),
    "planar7": dict(dof=7, out=7, fname="7dof_sdf_256x5_mesh.pt"),
    "franka": dict(dof=7, out=9, fname="franka_collision_model.pt"),
}
_net_cache = {}


def get_net(name):
    if name not in _net_cache:
        c = NETS[name]
        with contextlib.redirect_stdout(io.StringIO()):
            _net_cache[name] = rh.make_ref_net(ns, c["dof"], c["out"], c["fname"], torch)
    return _net_cache[name]


def export_weights():
    os.makedirs(os.path.join(HERE, "weights"), exist_ok=True)
    for name in NETS:
        sd = get_net(name).model.state_dict()
        out = {}
        for i in range(5):
            out[f"W{i}"] = sd[f"layers.0.{i}.0.weight"].numpy().astype(np.float32)
            out[f"b{i}"] = sd[f"layers.0.{i}.0.bias"].numpy().astype(np.float32)
        np.savez(os.path.join(HERE, "weights", f"{name}.npz"), **out)


def planar_dh(dof, L):
    dh_a = torch.zeros(dof + 1)
    dh_a[1:] = L
    return torch.vstack((dh_a * 0, dh_a * 0, dh_a, dh_a * 0)).T, dh_a


def franka_dh():
    pi = math.pi
    dh_a = torch.tensor([0, 0, 0, 0.0825, -0.0825, 0, 0.088, 0])
    dh_d = torch.tensor([0.333, 0, 0.316, 0, 0.384, 0, 0, 0.107])
    dh_alpha = torch.tensor([0, -pi / 2, pi / 2, pi / 2, -pi / 2, pi / 2, pi / 2, 0])
    return torch.vstack((dh_d, dh_a * 0, dh_a, dh_alpha)).T, dh_a


def shelf(n_pts=12):
    """Same construction as obstacleStreamer.py:87-108 (restated, not imported: that file opens sockets)."""
    r = 0.03
    length = max(1, 2 * n_pts - 2) * r * 1.5
    z0, x0, y0 = 0.15, 0.45, 0
    posA = torch.tensor([x0, y0, z0 + length, r])
    posB = posA + torch.tensor([length / 3, 0.0, 0.0, 0.0])
    line = posA + torch.linspace(0, 1, n_pts // 2).reshape(-1, 1) * (posB - posA)
    out = line
    for sphere in line:
        sphere_down = sphere - torch.tensor([0, 0, length, 0])
        line_down = sphere + torch.linspace(0, 1, n_pts).reshape(-1, 1) * (sphere_down - sphere)
        sphere_left = sphere + torch.tensor([0, -length / 2, -length / 2, 0])
        sphere_right = sphere + torch.tensor([0, length / 2, -length / 2, 0])
        line_lr = sphere_left + torch.linspace(0, 1, n_pts).reshape(-1, 1) * (sphere_right - sphere_left)
        top = line_lr + torch.tensor([0, 0, length / 2, 0])
        bottom = line_lr + torch.tensor([0, 0, -length / 2, 0])
        out = torch.vstack((out, line_down, line_lr, top, bottom))
    return out


def run_case(tag, net_name, q0, qf, dh_params, dh_a, obs, dt, H, N, K, nk, alpha_s, sigma_nom,
             dst_thr, ker_thr, ignored_links, q_lim=None, p=2, q_cur_batch=None, seed=0,
             kernel_spread=0.15, alpha_scale=1.0):
    net = get_net(net_name)
    d = q0.shape[0]
    DS = [ns.LinDS(qf), ns.LinDS(q0)]
    torch.manual_seed(seed)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        mppi = ns.MPPI(q0, qf, dh_params, obs, dt, H, N, DS, dh_a, net, K)
    mppi.Policy.sigma_c_nominal = sigma_nom
    mppi.Policy.alpha_s = alpha_s
    mppi.Policy.policy_upd_rate = 0.5
    mppi.Policy.p = p
    mppi.dst_thr = dst_thr
    mppi.ker_thr = ker_thr
    mppi.ignored_links = list(ignored_links)
    if q_lim is not None:
        mppi.Cost.q_min, mppi.Cost.q_max = q_lim
    if d != 7:
        mppi.Cost.rest = torch.zeros(d)         # value unused, shape bug in cost.py:46
    # pre-place nk kernels around the start state (SURVEY 8(d) synthetic inputs)
    g = torch.Generator().manual_seed(seed + 1)
    P = mppi.Policy
    P.n_kernels = nk
    if nk > 0:
        P.mu_c[:nk] = q0 + kernel_spread * torch.randn(nk, d, generator=g)
        P.sigma_c[:nk] = sigma_nom
        P.alpha_c[:nk] = alpha_scale * torch.randn(nk, d, generator=g)
    mu_c0, sigma_c0, alpha_c0 = P.mu_c.clone(), P.sigma_c.clone(), P.alpha_c.clone()
    if q_cur_batch is not None:
        mppi.q_cur = q_cur_batch
    torch.manual_seed(seed + 2)
    P.sample_policy()
    mu_tmp, sigma_tmp, alpha_tmp = P.mu_tmp.clone(), P.sigma_tmp.clone(), P.alpha_tmp.clone()
    with contextlib.redirect_stdout(sink):
        all_traj, cdist, kval, dots, acts = mppi.propagate()
        cost = mppi.get_cost()
        _, n_upd = mppi.shift_policy_means()
    out = dict(
        net=np.array(net_name), q0=q0, qf=qf, dh_params=dh_params, dh_a=dh_a, obs=obs,
        dt=np.float64(dt), H=np.int64(H), N=np.int64(N), K=np.int64(K), nk=np.int64(nk),
        dst_thr=np.float64(dst_thr), ker_thr=np.float64(ker_thr), p=np.float64(p),
        ignored_links=np.array(list(ignored_links), dtype=np.int64),
        q_min=mppi.Cost.q_min, q_max=mppi.Cost.q_max,
        q_cur=mppi.q_cur,
        mu_c0=mu_c0, sigma_c0=sigma_c0, alpha_c0=alpha_c0,
        mu_tmp=mu_tmp[:, :max(nk, 1)], sigma_tmp=sigma_tmp[:, :max(nk, 1)], alpha_tmp=alpha_tmp[:, :max(nk, 1)],
        all_traj=all_traj, closest_dist_all=cdist, kernel_val_all=mppi.kernel_val_all[:, :, :max(nk, 1)],
        dot_products=dots, kernel_activations=acts, qdot=mppi.qdot, nn_grad_last=mppi.nn_grad,
        norm_basis=mppi.norm_basis, cost=cost,
        mu_c1=P.mu_c.clone(), sigma_c1=P.sigma_c.clone(), alpha_c1=P.alpha_c.clone(),
        n_updated=np.int64(int(n_upd)),
    )
    out = {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
    np.savez_compressed(os.path.join(HERE, f"case_{tag}.npz"), **out)
    ncoll = int((cdist < 0).sum())
    print(f"case_{tag}: N={N} H={H} M={obs.shape[0]} K={K} nk={nk}  collisions={ncoll} "
          f"act>0: {int((acts > 0).sum())}  n_updated={int(n_upd)}  cost[{cost.min():.2f},{cost.max():.2f}]")


def run_distgrad(tag, net_name, obs, K, ignored_links, n, qrange, seed=0):
    net = get_net(net_name)
    d = NETS[net_name]["dof"]
    q0 = torch.zeros(d)
    dh_params, dh_a = (franka_dh() if net_name == "franka" else planar_dh(d, 1))
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        mppi = ns.MPPI(q0, q0 + 1, dh_params, obs, 0.1, 1, n, [ns.LinDS(q0 + 1)], dh_a, net, K)
    mppi.ignored_links = list(ignored_links)
    g = torch.Generator().manual_seed(seed)
    q = (torch.rand(n, d, generator=g) * 2 - 1) * qrange
    dist, grad = mppi.distance_repulsion_nn(q.clone(), aot=True)
    np.savez_compressed(os.path.join(HERE, f"distgrad_{tag}.npz"), net=np.array(net_name), q=q.numpy(),
                        obs=obs.numpy(), K=np.int64(K), ignored_links=np.array(list(ignored_links), dtype=np.int64),
                        distance=dist.detach().numpy(), nn_grad=grad[:n].detach().numpy())
    print(f"distgrad_{tag}: n={n} M={obs.shape[0]} K={K} dist[{dist.min():.3f},{dist.max():.3f}]")


def export_kat():
    rows = {
        "franka": [-0.88, 0.38, 0.5, -1, 0.45, 1.9, 0.31, 0.45, 0, 0.8],
        "planar7": [1.5708, 0, 0, 0, 0, 0, 0, 5, 0, 0],
        "planar2": [-3.14, 0, 6, 0, 0],
    }
    out = {}
    for name, x in rows.items():
        net = get_net(name)
        xt = torch.tensor([x], dtype=torch.float32)
        z, g, idx = net.functorch_vjp(xt)
        out[f"{name}_x"] = xt.numpy()
        out[f"{name}_z"] = z.numpy()
        out[f"{name}_grad"] = g.numpy()
        out[f"{name}_argmin"] = idx.numpy()
        # a batch of random rows as well
        gen = torch.Generator().manual_seed(7)
        xb = (torch.rand(64, len(x), generator=gen) * 2 - 1) * 2.0
        zb, gb, ib = net.functorch_vjp(xb)
        out[f"{name}_xb"], out[f"{name}_zb"], out[f"{name}_gb"], out[f"{name}_ib"] = \
            xb.numpy(), zb.numpy(), gb.numpy(), ib.numpy()
    np.savez_compressed(os.path.join(HERE, "kat.npz"), **out)
    print("kat.npz written")


def main():
    export_weights()
    export_kat()
    pi = math.pi
    # C1-like: planar 2-DoF (standalonePlanar2d.py:57-131)
    dh2, a2 = planar_dh(2, 3)
    obs2 = torch.tensor([[6.0, 0.0, 0, .5], [0.0, 4.5, 0, .5]])
    lim2 = (-0.99 * 3.14 * torch.ones(2), 0.99 * 3.14 * torch.ones(2))
    run_case("planar2", "planar2", torch.tensor([-3.14, 0.0]), torch.tensor([3.14, 0.0]), dh2, a2, obs2,
             dt=0.3, H=10, N=100, K=2, nk=3, alpha_s=2, sigma_nom=0.5, dst_thr=0.25, ker_thr=1e-3,
             ignored_links=[], q_lim=lim2, kernel_spread=0.3)
    # start close to the obstacle so collisions / repulsion / l_n are exercised
    run_case("planar2_near", "planar2", torch.tensor([-0.6, 0.9]), torch.tensor([3.14, 0.0]), dh2, a2, obs2,
             dt=0.3, H=12, N=64, K=2, nk=4, alpha_s=2, sigma_nom=0.5, dst_thr=0.25, ker_thr=1e-3,
             ignored_links=[], q_lim=lim2, kernel_spread=0.3, seed=3)
    # no kernels, K=1, single sample, two steps (integrator shape: frankaIntegrator.py:101-121)
    run_case("planar2_nk0", "planar2", torch.tensor([-1.0, 0.5]), torch.tensor([3.14, 0.0]), dh2, a2, obs2,
             dt=0.1, H=2, N=1, K=1, nk=0, alpha_s=0, sigma_nom=0.5, dst_thr=0.5, ker_thr=1e-3,
             ignored_links=[], q_lim=lim2, seed=4)
    # C4-like: dense field, per-sample start states, H=1 (standalonePlanar2d_policyPlots.py:160,257-259)
    G = 24
    grid = torch.meshgrid([torch.linspace(-pi, pi, G) for _ in range(2)], indexing="ij")
    q_tens = torch.stack(grid, dim=-1).reshape(-1, 2)
    run_case("field2", "planar2", torch.tensor([-3.14, 0.0]), torch.tensor([3.14, 0.0]), dh2, a2, obs2,
             dt=0.05, H=1, N=G * G, K=1, nk=3, alpha_s=0, sigma_nom=0.5, dst_thr=0.25, ker_thr=1e-3,
             ignored_links=[], q_lim=lim2, q_cur_batch=q_tens, kernel_spread=1.0, seed=5)
    # C2-like: planar 7-DoF (standalonePlanar7d.py:58-107)
    dh7, a7 = planar_dh(7, 1)
    obs7 = torch.tensor([[6, 2, 0, .5], [4., -1, 0, .5], [5, 0, 0, .5], [6, 6, 6, .1]])
    q0 = torch.zeros(7); q0[0] = pi / 2
    qf = torch.zeros(7); qf[0] = -pi / 2
    run_case("planar7", "planar7", q0, qf, dh7, a7, obs7, dt=0.3, H=12, N=64, K=1, nk=4, alpha_s=0.75,
             sigma_nom=0.5, dst_thr=0.25, ker_thr=1e-3, ignored_links=[], seed=6)
    q0b = torch.tensor([0.35, 0.1, -0.1, 0.0, 0.05, 0.0, 0.0])
    run_case("planar7_near", "planar7", q0b, qf, dh7, a7, obs7, dt=0.3, H=10, N=48, K=2, nk=6, alpha_s=0.75,
             sigma_nom=0.5, dst_thr=0.25, ker_thr=1e-3, ignored_links=[], seed=7)
    # C3-like: Franka shelf (config.yaml, frankaPlanner.py:53-90)
    dhf, af = franka_dh()
    sh = shelf(12)
    assert sh.shape[0] == 294
    q0f = torch.tensor([-0.88, 0.38, 0.5, -1, 0.45, 1.9, 0.31])
    qff = torch.tensor([-1.24, 1.53, 1.22, -1.21, -0.21, 1.55, 0.08])
    run_case("franka_shelf", "franka", q0f, qff, dhf, af, sh, dt=0.5, H=8, N=32, K=5, nk=5, alpha_s=3,
             sigma_nom=1.0, dst_thr=0.01, ker_thr=0.1, ignored_links=[0, 1, 2], seed=8)
    run_case("franka_shelf_b", "franka", qff, q0f, dhf, af, sh, dt=0.5, H=6, N=24, K=5, nk=8, alpha_s=3,
             sigma_nom=1.0, dst_thr=0.03, ker_thr=0.1, ignored_links=[0, 1, 2], seed=9, alpha_scale=2.0)
    # raw distance/gradient queries
    run_distgrad("franka", "franka", sh, 5, [0, 1, 2], 96, 2.0, seed=11)
    run_distgrad("planar7", "planar7", obs7, 2, [], 128, 3.0, seed=12)
    run_distgrad("planar2", "planar2", obs2, 2, [], 128, 3.0, seed=13)


if __name__ == "__main__":
    main()
