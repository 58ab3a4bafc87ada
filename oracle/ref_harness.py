"""Import and drive the UNMODIFIED reference -- test / baseline infrastructure, never product code.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs (`--impl reference`, `cpu_baseline`) may import this.
The reference tree is looked up in this order:
  1. $DSMPPI_REFERENCE                     (explicit)
  2. oracle/_ref/                          (staged by oracle/make_ref.sh; git-ignored, travels to the GPU box)
  3. /root/reference                       (the read-only checkout of the build container)
Two shims, neither touching arithmetic:
  1. matplotlib is not installed and fk_num.py:4 star-imports plots.py:1 -> oracle/shims/ holds an inert
     `matplotlib` / `mpl_toolkits`;
  2. nn_model.aot_lambda = nn_model.functorch_vjp -- the reference's own commented fallback (robot_sdf.py:161-162);
     functorch.compile.aot_function (robot_sdf.py:164-166) asserts under torch 2.11.
For 2-joint robots `Cost.rest` is set to zeros(d): cost.py:10-12 hard-codes a 7-vector whose value is never added to
the total (cost.py:20-21) but whose shape breaks `rest_cost`.
"""
import contextlib
import importlib
import io
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
SHIMS = os.path.join(HERE, "shims")
STAGED = os.path.join(HERE, "_ref")


def reference_root():
    env = os.environ.get("DSMPPI_REFERENCE")
    for cand in ([env] if env else []) + [STAGED, "/root/reference"]:
        if cand and os.path.isdir(os.path.join(cand, "python_scripts", "ds_mppi", "functions")):
            return cand
    return None


REF = reference_root() or "/root/reference"
REF_FUNCS = os.path.join(REF, "python_scripts/ds_mppi/functions")
REF_MLP = os.path.join(REF, "python_scripts/mlp_learn")
REF_MODELS = os.path.join(REF_MLP, "models")
CHECKPOINTS = {"planar2": "2dof_sdf_256x5_mesh.pt", "planar7": "7dof_sdf_256x5_mesh.pt",
               "franka": "franka_collision_model.pt", "toy2": "2dof_sdf_256x5_toy.pt"}


def available():
    return reference_root() is not None


def _stub_matplotlib():
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:  # noqa: BLE001
        if SHIMS not in sys.path:
            sys.path.insert(0, SHIMS)
        import matplotlib.pyplot  # noqa: F401


def load_reference():
    """Namespace with the reference's MPPI, LinDS, RobotSdfCollisionNet classes (and the MPPI module as `mod`)."""
    assert available(), "reference tree not present (run oracle/make_ref.sh in the build container)"
    _stub_matplotlib()
    for p in (REF_FUNCS, REF_MLP):
        if p not in sys.path:
            sys.path.append(p)
    ref_mppi = importlib.import_module("MPPI")
    assert os.path.realpath(ref_mppi.__file__).startswith(os.path.realpath(REF)), \
        f"`MPPI` resolved to {ref_mppi.__file__}, not the reference (is the drop-in directory on sys.path?)"
    ref_linds = importlib.import_module("LinDS")
    ref_sdf = importlib.import_module("sdf.robot_sdf")
    return types.SimpleNamespace(MPPI=ref_mppi.MPPI, LinDS=ref_linds.LinDS,
                                 RobotSdfCollisionNet=ref_sdf.RobotSdfCollisionNet, mod=ref_mppi, root=REF)


def make_ref_net(ns, dof, out_channels, fname, torch):
    """The reference's network object, prepared the way its scripts do (standalonePlanar7d.py:38-55)."""
    params = {"device": "cpu", "dtype": torch.float32}
    nn_model = ns.RobotSdfCollisionNet(in_channels=dof + 3, out_channels=out_channels,
                                       layers=[256] * 4, skips=[])
    nn_model.load_weights(os.path.join(REF_MODELS, fname), params)
    nn_model.model.to(**params)
    nn_model.model_jit = nn_model.model
    nn_model.model_jit = torch.jit.script(nn_model.model_jit)
    nn_model.model_jit = torch.jit.optimize_for_inference(nn_model.model_jit)
    nn_model.aot_lambda = nn_model.functorch_vjp      # shim 2
    return nn_model


class ReferenceIteration:
    """One MPPI object of the reference on a bench.py problem dictionary; `step()` is what bench.py times:
    propagate + get_cost + shift_policy_means (policy sampling excluded)."""

    def __init__(self, p, N, H, seed=0, policy=None):
        import torch
        self.torch = torch
        ns = load_reference()
        with contextlib.redirect_stdout(io.StringIO()):
            net = make_ref_net(ns, p["dof"], p["out"], CHECKPOINTS[p["net"]], torch)
            torch.manual_seed(seed)
            DS = [ns.LinDS(p["qf"].clone()), ns.LinDS(p["q0"].clone())]
            m = ns.MPPI(p["q0"].clone(), p["qf"].clone(), p["dh"].clone(), p["obs"].clone(), p["dt"], H, N, DS,
                        p["dh_a"].clone(), net, p["K"])
        m.dst_thr, m.ker_thr, m.ignored_links = p["dst_thr"], p["ker_thr"], list(p["ignored"])
        m.Cost.q_min, m.Cost.q_max = p["qlim"][0].clone(), p["qlim"][1].clone()
        if p["dof"] != 7:
            m.Cost.rest = torch.zeros(p["dof"])
        m.Policy.alpha_s = p["alpha_s"]
        m.Policy.sigma_c_nominal = p["sigma"]
        self.mppi, self.N, self.H, self.nk = m, N, H, p["nk"]
        if policy is not None:
            self.load_policy(*policy)

    def load_policy(self, mu_c, sigma_c, alpha_c, mu_tmp, sigma_tmp, alpha_tmp):
        P = self.mppi.Policy
        P.n_kernels = self.nk
        P.mu_c.copy_(mu_c); P.sigma_c.copy_(sigma_c); P.alpha_c.copy_(alpha_c)
        P.mu_tmp.copy_(mu_tmp); P.sigma_tmp.copy_(sigma_tmp); P.alpha_tmp.copy_(alpha_tmp)

    def set_q_cur(self, q):
        self.mppi.q_cur = q

    def step(self):
        m = self.mppi
        out = m.propagate()
        cost = m.get_cost()
        upd = m.shift_policy_means()
        return out, cost, upd
