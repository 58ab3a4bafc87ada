"""Golden fixtures for the 2-D point "toy" variant, produced by EXECUTING THE UNMODIFIED REFERENCE class
ds_mppi/functions/MPPI_toy.py (with cost_toy.py) in this container.

    python tests/golden/make_golden_toy.py      # build container only (needs /root/reference)

Writes tests/golden/weights/toy2.npz (the shipped 2dof_sdf_256x5_toy.pt, model_state_dict only) and
tests/golden/toycase_*.npz: inputs + every output of propagate / get_cost / shift_policy_means, driven as
scripts/standaloneToy2d.py:50-98 drives them (arc of 20 unit spheres, A = -I, N = 100, H = 10, dt = 0.5).
Same two import shims as make_golden.py (ref_harness.py); no arithmetic is touched.
"""
import contextlib
import importlib
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
ref_toy = importlib.import_module("MPPI_toy")
assert ref_toy.__file__.startswith(rh.REF), ref_toy.__file__


def toy_net():
    params = {"device": "cpu", "dtype": torch.float32}
    with contextlib.redirect_stdout(io.StringIO()):
        nn_model = ns.RobotSdfCollisionNet(in_channels=4, out_channels=1, layers=[256] * 4, skips=[])
        nn_model.load_weights(os.path.join(rh.REF_MODELS, "2dof_sdf_256x5_toy.pt"), params)
    nn_model.model.to(**params)
    nn_model.model_jit = torch.jit.optimize_for_inference(torch.jit.script(nn_model.model))
    nn_model.aot_lambda = nn_model.functorch_vjp      # shim 2 (robot_sdf.py:161-162)
    return nn_model


NET = toy_net()


def export_weights():
    sd = NET.model.state_dict()
    out = {}
    for i in range(5):
        out[f"W{i}"] = sd[f"layers.0.{i}.0.weight"].numpy().astype(np.float32)
        out[f"b{i}"] = sd[f"layers.0.{i}.0.bias"].numpy().astype(np.float32)
    np.savez(os.path.join(HERE, "weights", "toy2.npz"), **out)


def arc(n=20, r_arc=3.0, r_sph=1.0):
    """standaloneToy2d.py:57-64."""
    t = torch.linspace(-math.pi / 3, math.pi / 3, n)
    return torch.vstack([r_arc * torch.cos(t), r_arc * torch.sin(t), r_sph * torch.ones(n)]).transpose(0, 1).contiguous()


def run_case(tag, q0, qf, obs, A, dt, H, N, K, nk, alpha_s, sigma_nom, dst_thr, ker_thr, seed, kernel_spread=0.5,
             alpha_scale=1.0, q_cur_batch=None):
    d = q0.shape[0]
    dh_params = torch.zeros(4, 4)                       # standaloneToy2d.py:55 (dummy)
    torch.manual_seed(seed)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        mppi = ref_toy.MPPI(q0, qf, dh_params, obs, dt, H, N, A, 0, NET, K)
    mppi.Policy.sigma_c_nominal = sigma_nom
    mppi.Policy.alpha_s = alpha_s
    mppi.dst_thr = dst_thr
    mppi.ker_thr = ker_thr
    mppi.ignored_links = []
    mppi.Cost.q_min = -10 * torch.ones(d)
    mppi.Cost.q_max = 10 * torch.ones(d)
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
        all_traj, cdist, kval, dots = mppi.propagate()
        cost = mppi.get_cost()
        ret = mppi.shift_policy_means()
    assert ret == 0
    n_changed = int(((P.mu_c - mu_c0).abs().sum(1) + (P.alpha_c - alpha_c0).abs().sum(1) > 0).sum())
    out = dict(
        net=np.array("toy2"), q0=q0, qf=qf, obs=obs, A=A, dt=np.float64(dt), H=np.int64(H), N=np.int64(N),
        K=np.int64(K), nk=np.int64(nk), dst_thr=np.float64(dst_thr), ker_thr=np.float64(ker_thr),
        p=np.float64(P.p), upd_rate=np.float64(mppi.policy_upd_rate), q_cur=mppi.q_cur,
        mu_c0=mu_c0, sigma_c0=sigma_c0, alpha_c0=alpha_c0,
        mu_tmp=mu_tmp[:, :max(nk, 1)], sigma_tmp=sigma_tmp[:, :max(nk, 1)], alpha_tmp=alpha_tmp[:, :max(nk, 1)],
        all_traj=all_traj, closest_dist_all=cdist, kernel_val_all=mppi.kernel_val_all[:, :, :max(nk, 1)],
        dot_products=dots, qdot=mppi.qdot.reshape(N, d), nn_grad_last=mppi.nn_grad, norm_basis=mppi.norm_basis,
        cost=cost, mu_c1=P.mu_c.clone(), sigma_c1=P.sigma_c.clone(), alpha_c1=P.alpha_c.clone(),
        n_changed=np.int64(n_changed),
    )
    out = {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
    np.savez_compressed(os.path.join(HERE, f"toycase_{tag}.npz"), **out)
    print(f"toycase_{tag}: N={N} H={H} M={obs.shape[0]} K={K} nk={nk} collisions={int((cdist < 0).sum())} "
          f"kval>0: {int((mppi.kernel_val_all > 0).sum())} kernels changed={n_changed} "
          f"cost[{cost.min():.2f},{cost.max():.2f}]")


def main():
    export_weights()
    A = -1 * torch.diag(torch.ones(2))
    obs = arc()
    qf = torch.tensor([8.0, 0.0])
    # the script's own shape (standaloneToy2d.py:70-91)
    run_case("toy2", torch.tensor([-1.0, 0.3]), qf, obs, A, dt=0.5, H=10, N=100, K=1, nk=3, alpha_s=0.75, sigma_nom=0.1, dst_thr=0.25,
             ker_thr=0.05, seed=20, kernel_spread=1.0)
    # start inside the arc's mouth: collisions, repulsion, the steep sigmoids and K = 2 are exercised
    run_case("toy2_near", torch.tensor([1.2, 0.4]), qf, obs, A, dt=0.5, H=8, N=64, K=2, nk=4, alpha_s=0.75,
             sigma_nom=0.1, dst_thr=0.25, ker_thr=0.05, seed=21, kernel_spread=0.8, alpha_scale=2.0)
    # a rotating, non-symmetric nominal DS
    A2 = torch.tensor([[-1.0, 0.4], [-0.4, -0.7]])
    run_case("toy2_rot", torch.tensor([-0.5, 1.0]), qf, obs, A2, dt=0.5, H=10, N=48, K=1, nk=2, alpha_s=0.75,
             sigma_nom=0.1, dst_thr=0.25, ker_thr=0.01, seed=22, kernel_spread=1.0)
    # the 1 x 1 mppi_step object (standaloneToy2d.py:93-95)
    run_case("toy2_step", torch.tensor([-2.0, 0.3]), qf, obs, A, dt=0.1, H=1, N=1, K=1, nk=0, alpha_s=0.0,
             sigma_nom=0.1, dst_thr=0.1, ker_thr=0.5, seed=23)


if __name__ == "__main__":
    main()
