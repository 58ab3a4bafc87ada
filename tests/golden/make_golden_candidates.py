"""Golden vectors for TensorPolicyMPPI.check_traj_for_kernels, produced by the UNMODIFIED reference class
(ds_mppi/functions/policy.py:153-175) on the trajectories already stored in tests/golden/case_*.npz.

    python tests/golden/make_golden_candidates.py       # build container only (needs /root/reference)

Writes tests/golden/cand_<case>.npz: thresholds, the policy kernels used, and the candidate states returned."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_harness as rh  # noqa: E402
from tests.golden_util import case_names, load_npz  # noqa: E402

rh.load_reference()
import policy as ref_policy  # noqa: E402  (the reference's module, via ref_harness's sys.path)

assert ref_policy.__file__.startswith(rh.REF), ref_policy.__file__
params = {"device": "cpu", "dtype": torch.float32}
for tag in case_names():
    c = load_npz(f"case_{tag}")
    N, H, d = c["all_traj"].shape
    nk = int(c["nk"])
    P = ref_policy.TensorPolicyMPPI(N, d, params)
    P.n_kernels = nk
    P.mu_c[:] = c["mu_c0"]; P.sigma_c[:] = c["sigma_c0"]
    P.p = float(c["p"])
    dist, dots = c["closest_dist_all"], c["dot_products"]
    sets = []
    # thresholds spanning "nothing", "some", "everything" on this case's own value ranges
    for qd, qdot, thr_k in ((0.5, 0.5, 0.2), (0.9, 0.9, 0.9), (0.2, 0.3, 1e-3), (1.0, 1.0, 2.0)):
        thr_dist = float(torch.quantile(dist.flatten(), qd))
        thr_dot = float(torch.quantile(dots.flatten()[torch.isfinite(dots.flatten())], qdot))
        cand = P.check_traj_for_kernels(c["all_traj"], dist, dots, thr_dist, thr_k, thr_dot)
        sets.append((thr_dist, thr_k, thr_dot, cand.numpy()))
    out = {}
    for i, (a, b, e, cand) in enumerate(sets):
        out[f"thr{i}"] = np.array([a, b, e], dtype=np.float64)
        out[f"cand{i}"] = cand
    np.savez_compressed(os.path.join(HERE, f"cand_{tag}.npz"), n_sets=np.int64(len(sets)), **out)
    print(tag, [s[3].shape[0] for s in sets], "of", N * H)
