"""CPU tests (gloo, world_size 2) of the host-side sample-sharding logic (optimalmodulationds_b200/parallel.py):
shard ranges, the two all-reduces, and that summing per-shard packed partials reproduces the unsharded
policy update.  The partial vectors are produced by the oracle here (no GPU in this container); on the GPU
box the same packed layout comes from dsmppi_update_partial (tests/test_gpu_parity.py::test_sharded_update_*)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mppi_oracle as orc
from optimalmodulationds_b200 import parallel
from tests.golden_util import full_policy, load_npz


def test_shard_ranges_cover_and_balance():
    for n, w in ((4096, 8), (1000, 3), (7, 8), (1_000_000, 8)):
        spans = [parallel.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c = load_npz("case_planar7_near")
        nk, N, H, d = int(c["nk"]), int(c["N"]), int(c["H"]), 7
        mu, sg, al = (full_policy(c, k, N) for k in ("mu_tmp", "sigma_tmp", "alpha_tmp"))
        kv = torch.zeros(N, H, 50)
        kv[:, :, :nk] = c["kernel_val_all"]
        lo, hi = parallel.shard_range(N, rank, world)
        cost = c["cost"][lo:hi]
        # phase 1: local cost statistics -> all-reduce -> global beta
        # the layout of cost_stats_kernel: [sum cost, N, min cost, argmin (local)]
        stats = torch.stack((cost.sum(), torch.tensor(float(hi - lo)), cost.min(), cost.argmin().float()))
        parallel.allreduce_cost_stats(stats)                # ONE collective: SUM of the first two entries
        n_global = int(stats[1])
        assert n_global == N
        assert float(stats[2]) == float(cost.min())         # min / argmin stay per shard ...
        best = parallel.allreduce_best(stats, lo)           # ... until a caller asks for the global best sample
        assert int(best[1]) == int(c["cost"].argmin()) and float(best[0]) == float(c["cost"].min())
        beta = (stats[0] / stats[1]) / 50
        # phase 2: packed partial sums -> all-reduce
        packed = orc.policy_update_partials(cost, kv[lo:hi], c["kernel_activations"][lo:hi], mu[lo:hi], sg[lo:hi],
                                            al[lo:hi], nk, beta, owns_sample0=(rank == 0))
        parallel.allreduce_packed(packed)
        # phase 3: identical finalize on every rank
        mu1, sg1, al1, n_upd = parallel.finalize_from_packed(packed, nk, d, n_global, float(c["ker_thr"]), 0.1,
                                                            c["mu_c0"], c["sigma_c0"], c["alpha_c0"])
        out[rank] = (mu1, sg1, al1, n_upd, float(best[0]))
    finally:
        dist.destroy_process_group()


def test_sharded_policy_update_matches_reference_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    c = load_npz("case_planar7_near")
    for rank in range(world):
        mu1, sg1, al1, n_upd, cmin = out[rank]
        assert n_upd == int(c["n_updated"])
        assert cmin == pytest.approx(float(c["cost"].min()))
        torch.testing.assert_close(mu1, c["mu_c1"], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(sg1, c["sigma_c1"], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(al1, c["alpha_c1"], rtol=1e-4, atol=1e-6)
    # both ranks hold bit-identical policies afterwards
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][2], out[1][2])
