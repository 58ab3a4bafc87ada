"""Times compile-time variants of tc_exact.cu against each other on the GPU box.
`build` (here): one library per -D combination under optimalmodulationds_b200/variants/;
`run` (GPU box): a fresh process per library, same seeded workloads, best-of-3 of a 10-launch average."""
import itertools
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "optimalmodulationds_b200", "variants")
SWITCHES = {"TCX_ENC_PIPE": (0, 1), "TCX_DEFER_STG": (0, 1)}


def names():
    keys = list(SWITCHES)
    for combo in itertools.product(*(SWITCHES[k] for k in keys)):
        yield "_".join(f"{k[4:].lower()}{v}" for k, v in zip(keys, combo)), " ".join(f"-D{k}={v}" for k, v in zip(keys, combo))


def build():
    os.makedirs(VDIR, exist_ok=True)
    for name, flags in names():
        env = dict(os.environ, DSMPPI_EXTRA_NVCC_FLAGS=flags)
        subprocess.check_call([sys.executable, "-c", "from optimalmodulationds_b200 import build as b; b.build(force=True)"],
                              cwd=ROOT, env=env)
        os.replace(os.path.join(ROOT, "optimalmodulationds_b200", "libdsmppi_b200.so"), os.path.join(VDIR, f"lib_{name}.so"))
        print("built", name, flags)
    subprocess.check_call([sys.executable, "-c", "from optimalmodulationds_b200 import build as b; b.build(force=True)"], cwd=ROOT)


def one(lib):
    sys.path.insert(0, ROOT)
    import time
    import torch
    from optimalmodulationds_b200 import _capi
    _capi.LIB_PATH = lib
    from tests.golden_util import load_npz
    from tests.mppi_factory import make_mppi

    def timed(f, n=10):
        f()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            for _ in range(n):
                f()
            torch.cuda.synchronize()
            best = min(best, (time.perf_counter() - t0) / n * 1e3)
        return best

    out = []
    for case, n, M, pass1 in (("planar2", 300000, None, "exact"), ("franka_shelf", 4096, 300, "tc_f16")):
        torch.manual_seed(1)
        c = load_npz(f"case_{case}")
        if M is not None:
            obs = torch.rand(M, 4) * 1.2 - 0.6
            obs[:, 3] = 0.03
            c["obs"] = obs
        m = make_mppi(c, device="cuda", pass1=pass1)
        m.set_score_mode("tc_split")
        q = ((torch.rand(n, c["q0"].shape[0]) * 2 - 1) * 2.5).cuda()
        out.append(timed(lambda: m.distance_repulsion_nn(q)))
    for case in ("planar7", "franka_shelf"):
        c = load_npz(f"case_{case}")
        m = make_mppi(c, device="cuda", pass1="auto")
        m.set_score_mode("tc_split")
        out.append(timed(lambda: m.propagate()))
    print(f"{os.path.basename(lib):28s} dense600k {out[0]:.3f} ms  franka_cand {out[1]:.3f} ms  roll_planar7 {out[2]:.3f} ms  roll_franka {out[3]:.3f} ms",
          flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    elif sys.argv[1] == "one":
        one(sys.argv[2])
    else:
        for rep in range(2):
            for f in sorted(os.listdir(VDIR)):
                subprocess.call([sys.executable, os.path.abspath(__file__), "one", os.path.join(VDIR, f)])
