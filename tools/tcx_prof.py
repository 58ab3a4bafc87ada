"""Phase accounting of tc_exact_kernel (build the library with DSMPPI_EXTRA_NVCC_FLAGS=-DDSMPPI_TCX_PROF first).
`run` (GPU box): scores 600 k planar-2 rows once with the stamps on and writes gpurun_out/tcx_prof.bin;
`dense`: the same two launches with the product library (for `ncu -k regex:tc_exact_kernel -s 1 -c 1`);
`show` (anywhere): prints CTA 0's merged timeline (epilogue warps 0 and 4, MMA issuer) for one steady-state tile."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out", "tcx_prof.bin")

NAMES = {1: "enc done, A0 signalled", 60: "out layer: D ready", 61: "seed written, A0 signalled", 62: "W1^T: D ready",
         63: "tile done", 92: "issuer: got K quarter 0", 93: "issuer: got K quarter 1", 94: "issuer: got K quarter 2",
         95: "issuer: got K quarter 3", 100: "weights landed: N0 k<128", 101: "weights landed: N0 k>=128",
         228: "weights landed: N1 k<128", 229: "weights landed: N1 k>=128"}
NAMES.update({201: "    step_sample: nominal DS done", 202: "    step_sample: blend + e0 + dot done", 203: "    step_sample: sigmoids + activation done",
              204: "    step_sample: RBF policy done", 205: "    step_sample: modulation + Euler step done", 83: "  step: ranking done", 80: "  step: rows visible (bar.sync)", 81: "  step: ranking + modulation step done", 82: "  step: next state visible (bar.sync)",
              64: "  out: TMEM load back", 65: "  out: argmin done", 66: "  out: bar.sync passed", 67: "  seed: W5 rows read",
              68: "  seed: quarter 0 signalled", 70: "  L3 parked chunk 0 stored", 71: "  L3 quarter 0 signalled", 72: "  L3 E1: first TMEM load back",
              73: "  L3 E1: chunk 0 converted", 74: "  L3 E1: chunk 0 stored", 75: "  L3 E1: quarter 2 signalled",
              76: "  L3 E1: second TMEM load back", 77: "  L3 E1: chunk 1 converted"})
for l in range(8):
    tag = f"fwd L{l + 1}" if l < 4 else f"bwd l={l - 4}"
    NAMES[10 + l] = f"{tag}: D half0 ready"
    NAMES[20 + l] = f"{tag}: half0 converted (parked)"
    NAMES[30 + l] = f"{tag}: D half1 ready"
    NAMES[40 + l] = f"{tag}: half0 stored, quarters 0, 1 signalled"
    NAMES[50 + l] = f"{tag}: half1 stored, quarters 2, 3 signalled"


def run_rollout(case):
    """whole-horizon kernel (MODE 2): one propagate() of a golden case with the stamps on"""
    import torch
    from optimalmodulationds_b200 import _capi
    prof_lib = os.path.join(ROOT, "optimalmodulationds_b200", "libdsmppi_b200_prof.so")
    if os.path.exists(prof_lib):
        _capi.LIB_PATH = prof_lib
    from tests.golden_util import load_npz
    from tests.mppi_factory import make_mppi
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    m = make_mppi(load_npz(f"case_{case}"), device="cuda", pass1="auto")
    m.set_score_mode("tc_split")
    m.propagate()
    torch.cuda.synchronize()
    os.environ["DSMPPI_TCX_PROF_OUT"] = OUT
    m.propagate()
    torch.cuda.synchronize()
    print("wrote", OUT, os.path.getsize(OUT))


def run(prof=True):
    import torch
    from optimalmodulationds_b200 import _capi
    prof_lib = os.path.join(ROOT, "optimalmodulationds_b200", "libdsmppi_b200_prof.so")
    if prof and os.path.exists(prof_lib):          # tools/build_prof.sh: the instrumented build beside the product library
        _capi.LIB_PATH = prof_lib
    from tests.golden_util import load_npz
    from tests.mppi_factory import make_mppi
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    torch.manual_seed(1)
    c = load_npz("case_planar2")
    m = make_mppi(c, device="cuda", pass1="exact")
    m.set_score_mode("tc_split")
    q = ((torch.rand(300000, 2) * 2 - 1) * 2.5).cuda()
    m.distance_repulsion_nn(q)
    torch.cuda.synchronize()
    if prof:
        os.environ["DSMPPI_TCX_PROF_OUT"] = OUT
    m.distance_repulsion_nn(q)
    torch.cuda.synchronize()
    if prof:
        print("wrote", OUT, os.path.getsize(OUT))


def show(tile=3):
    a = np.fromfile(OUT, dtype=np.int64).reshape(4, 1024, 2)
    ev = []
    for r, who in enumerate(("epi w0", "epi w4", "issuer", "epi w0")):
        for i, t in a[r]:
            if t:
                ev.append((int(t), who, int(i)))
    ev.sort()
    # tile boundaries: event 1 (enc done) on epi w0
    starts = [t for t, who, i in ev if who == "epi w0" and i == 1]
    print(f"{len(starts)} tiles stamped; cycles per tile: {np.diff(starts).tolist()[:8]}")
    t0, t1 = starts[tile], starts[tile + 1]
    last = {}
    for t, who, i in ev:
        if t0 <= t < t1 + 200:
            d = t - last.get(who, t)
            last[who] = t
            try:
                print(f"{t - t0:8d}  (+{d:6d})  {who:7s} {NAMES.get(i, i)}")
            except BrokenPipeError:
                return


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run()
    elif sys.argv[1] == "rollout":
        run_rollout(sys.argv[2] if len(sys.argv) > 2 else "planar7")
    elif sys.argv[1] == "dense":          # the same launch with the product library: the target of an ncu capture
        run(prof=False)
    else:
        show(int(sys.argv[2]) if len(sys.argv) > 2 else 3)
