"""Golden fixtures for the SEDS nominal DS, produced by EXECUTING THE UNMODIFIED REFERENCE class
ds_mppi/functions/SEDS.py on its shipped models (ds_mppi/content/ds/*.mat), and one MPPI rollout of the reference
MPPI object with a SEDS in its DS_ARRAY (the configuration commented out at frankaIntegrator.py:70-71).

    python tests/golden/make_golden_seds.py     # build container only (needs /root/reference)

The reference's get_velocity only broadcasts for ONE state per call (SEDS.py:70-71), so velocities are collected
state by state and the MPPI case uses N_traj = 1 (the integrator shape, frankaIntegrator.py:101-121).
Writes tests/golden/seds_<model>.npz: the model arrays (Mu, Sigma, Priors, xT), query states, velocities; and
tests/golden/sedscase_*.npz."""
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
ref_seds = importlib.import_module("SEDS")
assert ref_seds.__file__.startswith(rh.REF)
DS_DIR = os.path.join(rh.REF, "python_scripts/ds_mppi/content/ds")


def queries(ds, seed):
    """States with well-conditioned responsibilities (around every Gaussian's input mean), states so far away that
    every responsibility underflows to exactly 0 (linear fallback), and states inside lin_thr of the goal."""
    g = torch.Generator().manual_seed(seed)
    d, G = ds.dof, ds.n_gaussians
    goal = ds.q_goal.reshape(-1)
    xs = [goal + ds.Mu[:d, j] + 0.03 * torch.randn(12, d, generator=g) for j in range(G)]
    xs.append(goal + 60.0 * torch.nn.functional.normalize(torch.randn(8, d, generator=g), dim=1))
    xs.append(goal + 2e-3 * torch.randn(6, d, generator=g))
    return torch.cat(xs)


def run_model(name, seed):
    ds = ref_seds.SEDS(os.path.join(DS_DIR, name + ".mat"))
    x = queries(ds, seed)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        v = torch.cat([ds.get_velocity(x[i:i + 1].clone()) for i in range(x.shape[0])])
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), Mu=ds.Mu.numpy(), Sigma=ds.Sigma.numpy(),
                        Priors=ds.Priors.numpy(), xT=ds.q_goal.numpy(), x=x.numpy(), velocity=v.numpy())
    print(f"{name}: d={ds.dof} G={ds.n_gaussians} n={x.shape[0]} lin fallbacks={sink.getvalue().count('lin!')} "
          f"|v| in [{v.norm(dim=1).min():.3g}, {v.norm(dim=1).max():.3g}]")
    return ds


def run_case(tag, model, H, dt, seed):
    pi = math.pi
    ds = ref_seds.SEDS(os.path.join(DS_DIR, model + ".mat"))
    with contextlib.redirect_stdout(io.StringIO()):
        net = rh.make_ref_net(ns, 7, 9, "franka_collision_model.pt", torch)
    dh_a = torch.tensor([0, 0, 0, 0.0825, -0.0825, 0, 0.088, 0])
    dh_d = torch.tensor([0.333, 0, 0.316, 0, 0.384, 0, 0, 0.107])
    dh_alpha = torch.tensor([0, -pi / 2, pi / 2, pi / 2, -pi / 2, pi / 2, pi / 2, 0])
    dh = torch.vstack((dh_d, dh_a * 0, dh_a, dh_alpha)).T
    goal = ds.q_goal.reshape(-1)
    q0 = goal + ds.Mu[:7, 0] + 0.02                     # on the demonstrated region
    g = torch.Generator().manual_seed(seed)
    obs = torch.cat((torch.tensor([[0.5, 0.0, 0.5]]) + 0.15 * torch.randn(12, 3, generator=g), 0.03 * torch.ones(12, 1)), 1)
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ns.MPPI(q0, goal, dh, obs, dt, H, 1, [ds], dh_a, net, 5)
    m.dst_thr, m.ker_thr, m.ignored_links = 0.01, 0.1, [0, 1, 2]
    m.Policy.sample_policy()
    with contextlib.redirect_stdout(io.StringIO()):
        all_traj, cdist, kval, dots, acts = m.propagate()
    out = dict(model=np.array(model), q0=q0, qf=goal, dh_params=dh, dh_a=dh_a, obs=obs, dt=np.float64(dt),
               H=np.int64(H), all_traj=all_traj, closest_dist_all=cdist, dot_products=dots, kernel_activations=acts,
               qdot=m.qdot.reshape(1, 7))
    out = {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
    np.savez_compressed(os.path.join(HERE, f"sedscase_{tag}.npz"), **out)
    print(f"sedscase_{tag}: H={H} dist[{cdist.min():.3f},{cdist.max():.3f}] step lengths "
          f"{(all_traj[0, 1:] - all_traj[0, :-1]).norm(dim=1)[:4].tolist()}")


def main():
    run_model("seds_left10", 41)
    run_model("seds_2d", 42)
    run_model("seds_sine", 43)
    run_case("franka_left10", "seds_left10", H=6, dt=0.05, seed=44)


if __name__ == "__main__":
    main()
