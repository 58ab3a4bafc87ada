"""Import the UNMODIFIED reference (read-only at /root/reference) in this container.

Only usable where /root/reference exists (the build container) -- never imported by the
`-m gpu` tests, smoke() or bench.py.  Two shims, neither touching arithmetic:
  1. a stub `matplotlib` / `mpl_toolkits` (fk_num.py:4 star-imports plots.py:1);
  2. nn_model.aot_lambda = nn_model.functorch_vjp (the reference's own commented fallback,
     robot_sdf.py:161-162; functorch aot_function asserts under torch 2.11).
"""
import os
import sys
import types
from unittest import mock

REF = os.environ.get("DSMPPI_REFERENCE", "/root/reference")
REF_FUNCS = os.path.join(REF, "python_scripts/ds_mppi/functions")
REF_MLP = os.path.join(REF, "python_scripts/mlp_learn")
REF_MODELS = os.path.join(REF_MLP, "models")


def available():
    return os.path.isdir(REF_FUNCS)


def _stub_matplotlib():
    if "matplotlib" in sys.modules:
        return
    mpl = types.ModuleType("matplotlib")
    plt = mock.MagicMock(name="matplotlib.pyplot")
    plt.plot.return_value = [mock.MagicMock()]
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    tk = types.ModuleType("mpl_toolkits")
    tk.mplot3d = mock.MagicMock()
    sys.modules["mpl_toolkits"] = tk
    sys.modules["mpl_toolkits.mplot3d"] = tk.mplot3d


def load_reference():
    """Returns a namespace with the reference's MPPI, LinDS, RobotSdfCollisionNet."""
    assert available(), "reference tree not present"
    _stub_matplotlib()
    for p in (REF_FUNCS, REF_MLP):
        if p not in sys.path:
            sys.path.append(p)
    import importlib
    ref_mppi = importlib.import_module("MPPI")
    assert ref_mppi.__file__.startswith(REF), ref_mppi.__file__
    ref_linds = importlib.import_module("LinDS")
    ref_sdf = importlib.import_module("sdf.robot_sdf")
    ns = types.SimpleNamespace(MPPI=ref_mppi.MPPI, LinDS=ref_linds.LinDS,
                               RobotSdfCollisionNet=ref_sdf.RobotSdfCollisionNet,
                               mod=ref_mppi)
    return ns


def make_ref_net(ns, dof, out_channels, fname, torch):
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
