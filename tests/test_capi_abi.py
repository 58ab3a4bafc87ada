"""CPU-side checks of the C ABI (no GPU, no compute): the shared library loads, exports every function
include/dsmppi_b200.h declares, the ctypes mirror of the structs has the C compiler's layout, and the product
fails loudly -- never falls back -- when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dsmppi_b200.h")


@pytest.fixture(scope="module")
def capi():
    from optimalmodulationds_b200 import build, _capi
    build.build()
    _capi.load()
    return _capi


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(dsmppi_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(capi):
    names = declared_functions()
    assert len(names) >= 20
    lib = capi.load()
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in the header but not exported by libdsmppi_b200.so"
    assert sorted(capi.EXPORTS) == names, "ctypes prototypes and header disagree"


def test_struct_layout_matches_the_c_compiler(capi, tmp_path):
    structs = {"dsmppi_net": capi.Net, "dsmppi_rollout_args": capi.RolloutArgs, "dsmppi_cost_args": capi.CostArgs,
               "dsmppi_update_args": capi.UpdateArgs, "dsmppi_iteration_host_args": capi.IterationHostArgs, "dsmppi_modulation": capi.Modulation, "dsmppi_seds": capi.Seds,
               "dsmppi_tick_args": capi.TickArgs}
    probes = {"dsmppi_rollout_args": ["q_goal", "mod", "distance_provider", "fk_span", "q_cur_dev", "norm_basis_dev"],
              "dsmppi_modulation": ["lvel_mid", "repulsion", "ds_A"],
              "dsmppi_seds": ["seds_thr", "priors_host", "A_host"],
              "dsmppi_cost_args": ["terms", "q_max", "cost_dev"],
              "dsmppi_update_args": ["variant", "N_global", "ker_thr", "alpha_c_dev"],
              "dsmppi_iteration_host_args": ["q_min", "cost_terms", "q_cur_host", "n_updated_host", "d2h_bytes", "exchange",
                                             "stats_dev", "owns_sample0", "N_global"],
              "dsmppi_tick_args": ["n_obs", "q_cur_host", "alpha_tmp_host", "nn_grad_all_host", "recaptured"],
              "dsmppi_net": ["W_host", "b_host"]}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "dsmppi_b200.h"', 'int main(void){']
    for s, fields in probes.items():
        lines.append(f'printf("{s} %zu\\n", sizeof({s}));')
        for f in fields:
            lines.append(f'printf("{s}.{f} %zu\\n", offsetof({s}, {f}));')
    lines.append('return 0;}')
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for s, cls in structs.items():
        assert int(out[s]) == C.sizeof(cls), s
        for f in probes[s]:
            assert int(out[f"{s}.{f}"]) == getattr(cls, f).offset, f"{s}.{f}"


def test_constants_and_pure_helpers(capi):
    lib = capi.load()
    assert lib.dsmppi_version() >= 100
    assert lib.dsmppi_update_packed_len(10, 7) == 1 + 10 * (2 * 7 + 3)
    text = open(HEADER).read()
    for name, val in (("DSMPPI_N_KERNEL_MAX", capi.N_KERNEL_MAX), ("DSMPPI_MAX_DOF", capi.MAX_DOF),
                      ("DSMPPI_MAX_LINKS", capi.MAX_LINKS), ("DSMPPI_MAX_CLOSEST", capi.MAX_CLOSEST)):
        assert int(re.search(rf"#define {name}\s+(\d+)", text).group(1)) == val


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_fallback(capi):
    lib = capi.load()
    net = capi.Net()
    net.n_dof, net.n_out, net.n_point_dim = 2, 2, 3
    W = [torch.zeros(256, 15)] + [torch.zeros(256, 256)] * 3 + [torch.zeros(2, 256)]
    b = [torch.zeros(256)] * 4 + [torch.zeros(2)]
    for i in range(5):
        net.W_host[i], net.b_host[i] = W[i].data_ptr(), b[i].data_ptr()
    dh = torch.zeros(3, 4)
    handle = C.c_void_p()
    rc = lib.dsmppi_ctx_create(C.byref(handle), C.byref(net), dh.data_ptr(), 16, 0)
    assert rc != 0 and not handle.value
    assert lib.dsmppi_last_error()                     # a message, not silence
    with pytest.raises(RuntimeError):
        capi.check(rc)
    from tests.golden_util import load_npz
    from tests.mppi_factory import make_mppi
    with pytest.raises(RuntimeError, match="CUDA"):
        make_mppi(load_npz("case_planar2"))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "optimalmodulationds_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "mppi_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
