"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the drop-in MPPI object and
hence through the C ABI, against (a) the committed golden outputs of the unmodified reference and (b) the
CPU oracle on fresh seeded inputs.  Tolerances follow BASELINE.json's north_star / SURVEY 8(c):
distances, gradients, velocities 1e-5 relative; trajectories, cost, policy 1e-4."""
import pytest
import torch

from oracle import mppi_oracle as orc
from tests.golden_util import case_names, frac_within, full_policy, load_npz, load_weights

pytestmark = pytest.mark.gpu

# absolute term of the distance tolerance (metres; the relative term is 1e-5 in both modes).  Against the reference's
# golden distances the IEEE-FFMA path has an rms error of 0.5e-7 .. 3.7e-7 m, the tensor-core split path 1.3e-7 ..
# 1.1e-6 m (tools/score_error_vs_golden.py: the truncating accumulator of the tensor core).
DIST_ATOL = 2e-6
# the Householder basis amplifies the error of the unit gradient e0 (measured: 1e-9 FFMA / 4e-6 tensor-core split where the
# gradient itself is small, both inside the 1e-5 the north star asks of gradients) by ~7x
BASIS_ATOL = 1e-5
DOT_ATOL = 1e-5         # e0 . v_hat: the unit-gradient error again (x3 in the tensor-core mode)
STABLE_FULL_HORIZON = ["planar2", "planar2_nk0", "field2", "franka_shelf", "franka_shelf_b", "planar2_near"]


def check(a, b, rtol, atol, name, min_frac=0.99, loose=20, kink_samples=0):
    """>= min_frac of the elements within rtol/atol and none beyond `loose` x that.  Quantities that follow the
    DIRECTION of the distance gradient (dot product, basis, modulated velocity) may name `kink_samples`: that many
    samples (leading index) may be off altogether -- a hidden unit whose pre-activation is within rounding distance
    of zero flips its ReLU under ANY change of summation order (the reference's own oneDNN vs MKL builds included),
    which moves the gradient by a finite amount while the distance stays within 1e-6 ("ReLU kink")."""
    a, b = a.detach().cpu(), b.detach().cpu()
    assert a.shape == b.shape, f"{name}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    if kink_samples and a.shape[0] > 1:
        viol = ((a.double() - b.double()).abs() / (atol + rtol * b.double().abs())).reshape(a.shape[0], -1)
        worst = viol.max(dim=1)[0]
        drop = torch.argsort(worst, descending=True)[:kink_samples]
        keep = torch.ones(a.shape[0], dtype=torch.bool)
        keep[drop[worst[drop] > 1.0]] = False
        a, b = a[keep], b[keep]
    f = frac_within(a, b, rtol, atol)
    min_frac = min(min_frac, 1.0 - 1.0 / max(a.numel(), 1)) if min_frac < 1.0 else 1.0   # always allow one outlier
    assert f >= min_frac, f"{name}: only {f:.4f} within rtol={rtol} (max abs diff {(a - b).abs().max():.3e})"
    assert frac_within(a, b, loose * rtol, loose * atol) == 1.0, \
        f"{name}: outliers beyond {loose}x tolerance (max abs diff {(a - b).abs().max():.3e})"


@pytest.fixture(scope="module")
def factory():
    from tests import mppi_factory
    return mppi_factory


@pytest.fixture(autouse=True, params=["tc_split", "ffma"])
def score_mode(request, factory):
    """Every test of this module runs in both scoring arithmetics: the default tensor-core path (split-fp16 tcgen05,
    csrc/tc_exact.cu) and the strict IEEE-FFMA path (csrc/exact_mlp.cu).  The RELATIVE tolerances are the same in
    both (north_star: distances / gradients / velocities 1e-5, trajectories / cost / policy 1e-4); the ABSOLUTE floors
    are wider for the tensor-core mode and set right here: distance 5e-6 m instead of 2e-6 m, unit-gradient dot
    product 3e-5 instead of 1e-5, Householder basis 5e-5 instead of 1e-5.  Why a floor at all: a distance near zero
    (a link touching an obstacle -- exactly where the modulation acts) has no meaningful relative error; the floors
    are the measured fp32-vs-fp64 noise of the quantity (IEEE FFMA 2.5e-7 of the rms distance, split-fp16 3.1e-7:
    DESIGN.md section 3) times the rms distance of the fixtures (0.26 .. 1.7 m) with a margin of 4 (FFMA: same
    summation order as the reference's oneDNN build) / 10 (tensor core: an independent rounding of the same
    quantity).  Callers that need the 2e-6 floor select `set_score_mode('ffma')` (15 % slower on the headline)."""
    global DIST_ATOL, BASIS_ATOL, DOT_ATOL
    factory.DEFAULT_SCORE = request.param
    DIST_ATOL = 2e-6 if request.param == "ffma" else 5e-6
    BASIS_ATOL = 1e-5 if request.param == "ffma" else 5e-5
    DOT_ATOL = 1e-5 if request.param == "ffma" else 3e-5
    yield request.param
    factory.DEFAULT_SCORE = None
    DIST_ATOL, BASIS_ATOL, DOT_ATOL = 2e-6, 1e-5, 1e-5


@pytest.mark.parametrize("tag", case_names())
def test_one_step_map_teacher_forced_vs_reference(tag, factory):
    """Every step of every golden case: feed the reference's own state, compare the step outputs."""
    c = load_npz(f"case_{tag}")
    N, H, nk, dt = int(c["N"]), int(c["H"]), int(c["nk"]), float(c["dt"])
    m = factory.make_mppi(c, device="cpu", H=1)
    for t in range(H):
        q = c["all_traj"][:, t, :]
        m.q_cur = q
        traj, dist, kv, dots, acts = m.propagate()
        assert traj.device.type == "cpu" and traj.shape == (N, 1, c["q0"].shape[0])
        check(dist[:, 0], c["closest_dist_all"][:, t], 1e-5, DIST_ATOL, f"dist[{t}]")
        # the normalised-gradient dot product sits on the fp32 noise floor of the 256-term backward sums:
        # >= 90% within 1e-5, everything within 2e-4 (test_gradient_error_vs_fp64 bounds it against fp64)
        kink = max(1, N // 200)
        check(dots[:, 0], c["dot_products"][:, t], 1e-5, DOT_ATOL, f"dot[{t}]", min_frac=0.9, kink_samples=kink)
        check(acts[:, 0], c["kernel_activations"][:, t], 1e-4, 1e-5, f"act[{t}]", kink_samples=kink)
        check(m.norm_basis[:, 0], c["norm_basis"][:, t], 1e-4, BASIS_ATOL, f"basis[{t}]", kink_samples=kink)
        if nk > 0:
            check(kv[:, 0, :], c["kernel_val_all"][:, t, :nk], 1e-4, 1e-6, f"kval[{t}]")
        if t == 0:
            check(m.qdot, c["qdot"], 1e-5, 1e-5, "qdot", kink_samples=kink)
        if t + 1 < H:
            check(q + dt * m.qdot, c["all_traj"][:, t + 1, :], 1e-5, 1e-5, f"traj[{t + 1}]", kink_samples=kink)


@pytest.mark.parametrize("tag", case_names())
@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_full_horizon_rollout_vs_reference(tag, device, factory):
    c = load_npz(f"case_{tag}")
    nk = int(c["nk"])
    m = factory.make_mppi(c, device=device)
    traj, dist, kv, dots, acts = m.propagate()
    assert traj.device.type == device
    # chaotic cases (see test_oracle_golden.py): only the first two steps are comparable, at 10x tolerance
    stable = tag in STABLE_FULL_HORIZON
    steps = int(c["H"]) if stable else 2
    tight = 1.0 if stable else 10.0
    # free-running: a sample that crosses a ReLU kink of the network within rounding distance takes another path from
    # there on (see check()); at most max(1, N/200) such samples are set aside, everything else is held to the tolerance
    kink = max(1, int(c["N"]) // 200)
    check(dist[:, :steps], c["closest_dist_all"][:, :steps], 1e-5 * tight, DIST_ATOL * tight, "closest_dist_all",
          kink_samples=kink)
    check(dots[:, :steps], c["dot_products"][:, :steps], 1e-5 * tight, DOT_ATOL * tight, "dot_products", min_frac=0.9,
          kink_samples=kink)
    check(m.qdot, c["qdot"], 1e-5, 1e-5, "qdot", kink_samples=kink)
    check(traj[:, :steps], c["all_traj"][:, :steps], 1e-4 * tight, 1e-5 * tight, "all_traj", kink_samples=kink)
    check(acts[:, :steps], c["kernel_activations"][:, :steps], 1e-4 * tight, 1e-5 * tight, "kernel_activations",
          kink_samples=kink)
    if nk > 0:
        check(kv[:, :steps], c["kernel_val_all"][:, :steps, :nk], 1e-4 * tight, 1e-6 * tight, "kernel_val_all",
              kink_samples=kink)
    check(m.norm_basis[:, :steps], c["norm_basis"][:, :steps], 1e-4 * tight, BASIS_ATOL * tight, "norm_basis",
          kink_samples=kink)
    assert m.kernel_val_all.shape == (int(c["N"]), int(c["H"]), 50)
    if tag in STABLE_FULL_HORIZON and torch.isfinite(c["cost"]).all():
        cost = m.get_cost()
        check(cost, c["cost"], 1e-4, 1e-3, "cost")
        _, n_upd = m.shift_policy_means()
        assert int(n_upd) == int(c["n_updated"])
        check(m.Policy.mu_c, c["mu_c1"], 1e-4, 1e-5, "mu_c")
        check(m.Policy.sigma_c, c["sigma_c1"], 1e-4, 1e-5, "sigma_c")
        check(m.Policy.alpha_c, c["alpha_c1"], 1e-4, 1e-5, "alpha_c")


@pytest.mark.parametrize("tag", case_names())
def test_cost_and_policy_update_on_reference_trajectories(tag, factory):
    """Cost / update kernels in isolation: inputs are the reference's own rollout outputs."""
    c = load_npz(f"case_{tag}")
    nk, N, H = int(c["nk"]), int(c["N"]), int(c["H"])
    m = factory.make_mppi(c, device="cpu")
    cost = m.Cost.evaluate_costs(c["all_traj"], c["closest_dist_all"])
    if not torch.isfinite(c["cost"]).all():
        assert torch.equal(torch.isfinite(cost), torch.isfinite(c["cost"]))
        return
    check(cost, c["cost"], 1e-5, 1e-4, "cost")
    m.cur_cost = c["cost"]
    kv = torch.zeros(N, H, 50)
    kv[:, :, :max(nk, 1)] = c["kernel_val_all"]
    m.kernel_val_all, m.kernel_activations = kv, c["kernel_activations"]
    _, n_upd = m.shift_policy_means()
    assert int(n_upd) == int(c["n_updated"])
    check(m.Policy.mu_c, c["mu_c1"], 1e-5, 1e-6, "mu_c")
    check(m.Policy.sigma_c, c["sigma_c1"], 1e-5, 1e-6, "sigma_c")
    check(m.Policy.alpha_c, c["alpha_c1"], 1e-5, 1e-6, "alpha_c")


@pytest.mark.parametrize("name", ["franka", "planar7", "planar2"])
def test_distance_repulsion_vs_reference(name, factory):
    g = load_npz(f"distgrad_{name}")
    c = load_npz({"franka": "case_franka_shelf", "planar7": "case_planar7", "planar2": "case_planar2"}[name])
    c = dict(c, obs=g["obs"], K=g["K"], ignored_links=g["ignored_links"])
    m = factory.make_mppi(c, device="cpu", H=1)
    dist, grad = m.distance_repulsion_nn(g["q"])
    check(dist, g["distance"], 1e-5, DIST_ATOL, "distance", min_frac=1.0)
    check(grad, g["nn_grad"], 1e-4, 1e-4 * g["nn_grad"].abs().max().item(), "nn_grad", min_frac=1.0)


@pytest.mark.parametrize("name,N,H,M,K,p", [("franka", 300, 3, 37, 5, 2.0), ("planar7", 257, 4, 5, 2, 2.0),
                                             ("planar2", 129, 5, 3, 3, 2.0), ("planar7", 131, 3, 4, 1, 3.0),
                                             ("planar2", 77, 4, 2, 2, 1.5)])
def test_rollout_vs_oracle_random_inputs(name, N, H, M, K, p, factory):
    """Fresh seeded inputs (ragged sizes: N not a multiple of any tile) against the CPU oracle; p != 2 exercises the
    general p-norm of the RBF kernels (policy.py:186-199), which lives out of line in the step."""
    torch.manual_seed(1234)
    d = {"franka": 7, "planar7": 7, "planar2": 2}[name]
    W, b = load_weights(name)
    net = orc.Net(W, b)
    base = load_npz({"franka": "case_franka_shelf", "planar7": "case_planar7", "planar2": "case_planar2"}[name])
    scale = 0.6 if name == "franka" else 4.0
    obs = torch.cat(((torch.rand(M, 3) - 0.5) * 2 * scale, 0.03 + 0.2 * torch.rand(M, 1)), 1)
    q_cur = base["q0"] + 0.3 * torch.randn(N, d)
    nk = 6
    c = dict(base, obs=obs, K=K, N=N, H=H, nk=nk)
    m = factory.make_mppi(c, device="cpu", N=N, H=H, q_cur=q_cur, copy_policy=False)
    P = m.Policy
    P.n_kernels = nk
    P.mu_c[:nk] = base["q0"] + 0.3 * torch.randn(nk, d)
    P.sigma_c[:nk] = 0.7
    P.alpha_c[:nk] = torch.randn(nk, d)
    P.alpha_s = 1.5
    P.p = p
    P.sample_policy()
    ign = base["ignored_links"].tolist()
    prm = orc.RolloutParams(dt=float(base["dt"]), dt_H=1, n_closest_obs=K, dst_thr=float(base["dst_thr"]),
                            ignored_links=ign, p=p, with_basis=False)
    traj, dist, kv, dots, acts = m.propagate()
    # teacher-forced against the oracle, step by step, using the GPU's own states
    for t in range(H):
        o = orc.rollout(net, traj[:, t, :], base["qf"], obs, P.mu_tmp, P.sigma_tmp, P.alpha_tmp, nk, prm, N)
        check(dist[:, t], o.closest_dist_all[:, 0], 1e-5, DIST_ATOL, f"dist[{t}]")
        kink = max(1, N // 200)
        check(dots[:, t], o.dot_products[:, 0], 1e-5, DOT_ATOL, f"dot[{t}]", min_frac=0.9, kink_samples=kink)
        check(acts[:, t], o.kernel_activations[:, 0], 1e-4, 1e-5, f"act[{t}]", kink_samples=kink)
        check(kv[:, t, :], o.kernel_val_all[:, 0, :nk], 1e-4, 1e-6, f"kval[{t}]")
        if t + 1 < H:
            check(traj[:, t + 1, :], traj[:, t, :] + float(base["dt"]) * o.qdot, 1e-5, 1e-5, f"traj[{t + 1}]",
                  kink_samples=kink)
    cost = m.get_cost()
    ocost = orc.evaluate_costs(traj, dist, base["qf"], base["dh_params"], base["q_min"], base["q_max"])
    check(cost, ocost, 1e-5, 1e-4, "cost")
    mu0, sg0, al0 = P.mu_c.clone(), P.sigma_c.clone(), P.alpha_c.clone()
    _, n_upd = m.shift_policy_means()
    mu1, sg1, al1, on, _ = orc.policy_update(cost, m.kernel_val_all, acts, P.mu_tmp, P.sigma_tmp, P.alpha_tmp, mu0,
                                             sg0, al0, nk, float(base["ker_thr"]))
    assert int(n_upd) == on
    check(P.mu_c, mu1, 1e-4, 1e-6, "mu_c")
    check(P.alpha_c, al1, 1e-4, 1e-6, "alpha_c")


@pytest.mark.parametrize("name", ["franka", "planar7", "planar2"])
def test_gradient_error_vs_fp64(name, factory):
    """The GPU's distance / gradient are as close to an fp64 evaluation of the same network as the
    reference's own fp32 numbers are (so residual GPU-vs-reference differences are fp32 noise, not bias)."""
    g = load_npz(f"distgrad_{name}")
    c = load_npz({"franka": "case_franka_shelf", "planar7": "case_planar7", "planar2": "case_planar2"}[name])
    c = dict(c, obs=g["obs"], K=g["K"], ignored_links=g["ignored_links"])
    m = factory.make_mppi(c, device="cpu", H=1)
    dist, grad = m.distance_repulsion_nn(g["q"])
    W, b = load_weights(name)
    net64 = orc.Net([w.double() for w in W], [x.double() for x in b])
    d64, g64 = orc.distance_repulsion(net64, g["q"].double(), g["obs"].double(), int(g["K"]),
                                      g["ignored_links"].tolist())
    ref_err_d = (g["distance"].double() - d64).abs().max().item()
    gpu_err_d = (dist.double() - d64).abs().max().item()
    ref_err_g = (g["nn_grad"].double() - g64).abs().max().item()
    gpu_err_g = (grad.double() - g64).abs().max().item()
    print(f"{name}: distance err vs fp64  ref {ref_err_d:.2e} gpu {gpu_err_d:.2e};  grad err ref {ref_err_g:.2e} "
          f"gpu {gpu_err_g:.2e}")
    assert gpu_err_d <= 4 * ref_err_d + 1e-6
    assert gpu_err_g <= 4 * ref_err_g + 1e-6 * g64.abs().max().item()


def test_tensor_core_prefilter_error_bound(factory):
    """tcgen05 fp16 / bf16 pass 1 vs fp32 pass 1 on the Franka shelf as shipped: a sanity bound on the prefilter's
    error (the guard band itself is calibrated per network and obstacle set: DESIGN.md section 3; measured here
    ~1.5 mm with fp16 accumulators in the hidden layers, ~1 mm with fp32 ones)."""
    c = load_npz("case_franka_shelf")
    m = factory.make_mppi(c, device="cuda", N=32, H=1)
    torch.manual_seed(0)
    q = c["q0"].cuda() + 0.5 * torch.randn(700, 7, device="cuda")        # 700*294 rows: ragged last tile
    ex = m.debug_pass1(q, "exact")
    for mode, bound in (("tc_f16", 3e-3), ("tc_bf16", 1.8e-2)):
        tc = m.debug_pass1(q, mode)
        err = (tc - ex).abs().max().item()
        print(f"{mode}: max |d_tc - d_fp32| = {err:.3e} m over {ex.numel()} pairs")
        assert err < bound


@pytest.mark.parametrize("tag,N,H", [("franka_shelf", 777, 6), ("franka_shelf_b", 300, 4)])
def test_tensor_core_mode_is_bitwise_identical_to_fp32_mode(tag, N, H, factory):
    """The prefilter only prunes; ranking, distances and gradients come from the fp32 path, so a rollout in
    tc mode must reproduce the fp32-mode rollout BIT FOR BIT (and never overflow the candidate band)."""
    c = load_npz(f"case_{tag}")
    torch.manual_seed(5)
    outs = {}
    q_cur = c["q0"] + 0.2 * torch.randn(N, 7)
    for mode in ("exact", "tc_f16"):
        m = factory.make_mppi(c, device="cuda", N=N, H=H, q_cur=q_cur, pass1=mode, copy_policy=False)
        P = m.Policy
        P.alpha_s = 3.0
        torch.manual_seed(11)
        P.sample_policy()
        traj, dist, kv, dots, acts = m.propagate()
        cost = m.get_cost()
        outs[mode] = (traj.clone(), dist.clone(), dots.clone(), acts.clone(), cost.clone())
        if mode == "tc_f16":
            # (`band_overflows` counts samples whose guard band held more than 16 obstacles: they take more of the shared
            # candidate list, nothing is truncated -- what must hold is that no call fell back to all-pairs fp32)
            st, xs = m.pass1_stats(), m.exactness_stats()
            assert st["mode"] == 1 and xs["exact_fallbacks"] == 0, (st, xs)
            print(f"{tag}: re-scored pairs per state-step = {st['rescored_pairs'] / (N * H):.2f} (K = {int(c['K'])})")
    for a, b2 in zip(outs["exact"], outs["tc_f16"]):
        assert torch.equal(a, b2)


def test_sharded_update_kernels_match_unsharded(factory):
    """SURVEY 8(e) on one GPU: run the cost-stats / partial kernels on two sample shards, add the packed
    vectors (what the NCCL all-reduce does), finalize -> same policy as the unsharded update."""
    from optimalmodulationds_b200 import _capi
    c = load_npz("case_franka_shelf")
    N, H, nk, d = int(c["N"]), int(c["H"]), int(c["nk"]), 7
    m = factory.make_mppi(c, device="cuda")
    m.propagate(); m.get_cost()
    P = m.Policy
    ref_mu, ref_sg, ref_al = P.mu_c.clone(), P.sigma_c.clone(), P.alpha_c.clone()
    mu0, sg0, al0 = P.mu_c.clone(), P.sigma_c.clone(), P.alpha_c.clone()
    _, n_ref = m.shift_policy_means()
    ref_mu, ref_sg, ref_al = P.mu_c.clone(), P.sigma_c.clone(), P.alpha_c.clone()
    lib, ctx, st = m._lib, m._ctx, m._stream()
    L = int(lib.dsmppi_update_packed_len(nk, d))
    halves = [(0, N // 2), (N // 2, N)]
    stats = []
    for lo, hi in halves:
        s_ = torch.empty(4, device="cuda")
        _capi.check(lib.dsmppi_update_cost_stats(ctx, m.cur_cost[lo:hi].contiguous().data_ptr(), hi - lo, s_.data_ptr(), st))
        stats.append(s_)
    # layout [sum cost, N, min cost, argmin]: the SUM all-reduce acts on the first two entries
    g = torch.stack((stats[0][0] + stats[1][0], stats[0][1] + stats[1][1], torch.minimum(stats[0][2], stats[1][2]),
                     stats[0][3]))
    assert int(g[1]) == N
    total = torch.zeros(L, device="cuda")
    args = None
    for r, (lo, hi) in enumerate(halves):
        a = _capi.UpdateArgs()
        a.N, a.H, a.n_kernels, a.owns_sample0, a.N_global = hi - lo, H, nk, 1 if r == 0 else 0, N
        a.ker_thr, a.upd_rate = float(m.ker_thr), float(m.policy_upd_rate)
        keep = [m.cur_cost[lo:hi].contiguous(), m.kernel_val_all[lo:hi].contiguous(),
                m.kernel_activations[lo:hi].contiguous(), P.mu_tmp[lo:hi].contiguous(),
                P.sigma_tmp[lo:hi].contiguous(), P.alpha_tmp[lo:hi].contiguous()]
        (a.cost_dev, a.kernel_val_all_dev, a.kernel_activations_dev, a.mu_tmp_dev, a.sigma_tmp_dev,
         a.alpha_tmp_dev) = [t.data_ptr() for t in keep]
        packed = torch.empty(L, device="cuda")
        _capi.check(lib.dsmppi_update_partial(ctx, _capi.C.byref(a), g.data_ptr(), packed.data_ptr(), st))
        torch.cuda.synchronize()
        total += packed
        args = a
    args.mu_c_dev, args.sigma_c_dev, args.alpha_c_dev = mu0.data_ptr(), sg0.data_ptr(), al0.data_ptr()
    n_upd = torch.zeros(1, dtype=torch.int32, device="cuda")
    _capi.check(lib.dsmppi_update_finalize(ctx, _capi.C.byref(args), total.data_ptr(), n_upd.data_ptr(), st))
    torch.cuda.synchronize()
    assert int(n_upd) == int(n_ref)
    check(mu0, ref_mu, 1e-5, 1e-6, "mu_c", min_frac=1.0)
    check(sg0, ref_sg, 1e-5, 1e-6, "sigma_c", min_frac=1.0)
    check(al0, ref_al, 1e-5, 1e-6, "alpha_c", min_frac=1.0)


def test_constructor_rejects_missing_cuda_path(monkeypatch, factory):
    """The product must fail loudly rather than fall back when the shared library is unavailable."""
    from optimalmodulationds_b200 import _capi
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "LIB_PATH", "/nonexistent/libdsmppi_b200.so")
    c = load_npz("case_planar2")
    with pytest.raises(RuntimeError):
        factory.make_mppi(c)


def _subset_matches_big_run(c, factory, N, H, q_big, idx, pass1):
    """Roll out a big batch, then only the samples `idx` of it in a fresh object: per-sample results must be
    BITWISE equal (samples never interact; blocks, tiles and candidate-row order must not leak into results)."""
    big = factory.make_mppi(c, device="cuda", N=N, H=H, q_cur=q_big, pass1=pass1, copy_policy=False)
    torch.manual_seed(21)
    big.Policy.sample_policy()
    traj, dist, kv, dots, acts = big.propagate()
    cost = big.get_cost()
    assert torch.isfinite(traj).all() and torch.isfinite(dist).all()
    small = factory.make_mppi(c, device="cuda", N=len(idx), H=H, q_cur=q_big[idx], pass1=pass1, copy_policy=False)
    for name in ("mu_tmp", "sigma_tmp", "alpha_tmp"):
        getattr(small.Policy, name).copy_(getattr(big.Policy, name)[idx.to(traj.device)])
    traj_s, dist_s, kv_s, dots_s, acts_s = small.propagate()
    cost_s = small.get_cost()
    j = idx.to(traj.device)
    for a, b2, name in ((traj[j], traj_s, "all_traj"), (dist[j], dist_s, "closest_dist_all"), (kv[j], kv_s, "kval"),
                        (dots[j], dots_s, "dots"), (acts[j], acts_s, "acts"), (cost[j], cost_s, "cost"),
                        (big.qdot[j], small.qdot, "qdot")):
        assert torch.equal(a, b2), f"{name}: big-batch rows differ from the same samples rolled out alone"
    return big, small


def test_dense_field_1m_points_config4(factory):
    """BASELINE config 4 at full size: 10^6 grid points x 1 step (policy-plot path).  Size-independent checks:
    a sample subset spanning the internal block boundaries equals the same samples run alone bit for bit, and
    that subset matches the CPU oracle."""
    import math
    c = load_npz("case_field2")
    G = 1000
    g = torch.linspace(-math.pi, math.pi, G)
    q = torch.stack(torch.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2).contiguous()
    N = q.shape[0]
    gen = torch.Generator().manual_seed(3)
    idx = torch.cat((torch.tensor([0, 1, 262143, 262144, 262145, 524287, 524288, N - 2, N - 1]),
                     torch.randint(0, N, (1500,), generator=gen))).unique()
    big, small = _subset_matches_big_run(c, factory, N, 1, q, idx, "exact")
    W, b = load_weights(c["net"])
    prm = orc.RolloutParams(dt=float(c["dt"]), dt_H=1, n_closest_obs=int(c["K"]), dst_thr=float(c["dst_thr"]),
                            ignored_links=c["ignored_links"].tolist(), p=float(c["p"]), with_basis=False)
    P = small.Policy
    o = orc.rollout(orc.Net(W, b), q[idx], c["qf"], c["obs"], P.mu_tmp.cpu(), P.sigma_tmp.cpu(), P.alpha_tmp.cpu(),
                    int(c["nk"]), prm, len(idx))
    check(small.closest_dist_all, o.closest_dist_all, 1e-5, DIST_ATOL, "closest_dist_all")
    check(small.qdot, o.qdot, 1e-5, 1e-5, "qdot")
    check(small.kernel_activations, o.kernel_activations, 1e-4, 1e-5, "kernel_activations")


def test_franka_large_batch_is_blockwise_consistent(factory):
    """A Franka-shelf batch larger than one internal sample block (2^18) through the tensor-core path."""
    c = load_npz("case_franka_shelf")
    N, H = 270_000, 2
    gen = torch.Generator().manual_seed(4)
    q = c["q0"] + 0.25 * torch.randn(N, 7, generator=gen)
    idx = torch.cat((torch.tensor([0, 262143, 262144, N - 1]), torch.randint(0, N, (300,), generator=gen))).unique()
    big, _ = _subset_matches_big_run(c, factory, N, H, q, idx, "tc_f16")
    st, xs = big.pass1_stats(), big.exactness_stats()
    assert st["mode"] == 1 and xs["exact_fallbacks"] == 0, (st, xs)     # crowded bands are re-scored, not truncated


@pytest.mark.parametrize("tag", case_names())
def test_kernel_candidates_vs_reference(tag, factory):
    """Policy.check_traj_for_kernels through the CUDA pass (dsmppi_kernel_candidates) on the reference's own
    trajectories: same candidate states, same (sample, step) order as the reference class returned."""
    c = load_npz(f"case_{tag}")
    g = load_npz(f"cand_{tag}")
    m = factory.make_mppi(c, device="cpu")
    for i in range(int(g["n_sets"])):
        thr_dist, thr_kernel, thr_dot = (float(x) for x in g[f"thr{i}"])
        cand = m.Policy.check_traj_for_kernels(c["all_traj"], c["closest_dist_all"], c["dot_products"], thr_dist,
                                               thr_kernel, thr_dot)
        assert cand.device.type == "cpu" and cand.shape == g[f"cand{i}"].shape, (cand.shape, g[f"cand{i}"].shape)
        assert torch.equal(cand, g[f"cand{i}"])


@pytest.mark.parametrize("tag", ["planar2", "planar2_near", "planar2_nk0", "field2", "planar7", "planar7_near"])
def test_whole_horizon_kernel_is_bitwise_identical_to_per_step_launches(tag, factory, score_mode):
    """Small obstacle sets are rolled out by ONE launch (rollout_fused_kernel, or tc_exact_kernel in whole-horizon
    mode: network tile + ranking + step inside the kernel's own horizon loop); it must reproduce the per-step launch
    sequence of the same scoring arithmetic bit for bit."""
    c = load_npz(f"case_{tag}")
    outs = []
    for whole in (True, False):
        m = factory.make_mppi(c, device="cuda")
        m.set_whole_horizon(whole)
        n0 = m.launch_count()
        res = [x.clone() for x in m.propagate()]
        launches = m.launch_count() - n0
        outs.append((res + [m.qdot.clone(), m.norm_basis.clone()], launches))
    (a, la), (b, lb) = outs
    H = int(c["H"])
    assert la <= 4 and lb >= 3 * H, (la, lb)      # obstacle encodings (2) + init + fused vs 3 launches per step
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def _host_buffers(m, q_cur):
    N, H, d = m.N_traj, m.dt_H, m.n_dof
    P = m.Policy
    pin = lambda x: x.detach().cpu().contiguous().clone().pin_memory()  # noqa: E731
    return dict(q_cur=pin(q_cur), mu_tmp=pin(P.mu_tmp), sigma_tmp=pin(P.sigma_tmp), alpha_tmp=pin(P.alpha_tmp),
                mu_c=pin(P.mu_c), sigma_c=pin(P.sigma_c), alpha_c=pin(P.alpha_c),
                all_traj=pin(torch.empty(N, H, d)), closest_dist_all=pin(torch.empty(N, H)),
                kernel_val_all=pin(torch.zeros(N, H, 50)), dot_products=pin(torch.empty(N, H)),
                kernel_activations=pin(torch.empty(N, H)), qdot=pin(torch.empty(N, d)), cost=pin(torch.empty(N)),
                n_updated=pin(torch.zeros(1, dtype=torch.int32)))


@pytest.mark.parametrize("tag,N,H,batch_start", [("planar7", 700, 5, False), ("field2", 1000, 1, True),
                                                  ("franka_shelf", 333, 3, False)])
def test_host_buffer_iteration_matches_device_path_and_is_chunk_invariant(tag, N, H, batch_start, factory, monkeypatch):
    """dsmppi_iteration_host (what a CPU-tensor caller and bench.py's e2e pay) returns exactly what
    propagate + get_cost + shift_policy_means return, and its chunk pipeline for large batches (H2D | rollout + cost |
    D2H of neighbouring sample chunks overlapped on three streams) changes no bit: forced here with a 64-sample chunk."""
    c = load_npz(f"case_{tag}")
    gen = torch.Generator().manual_seed(11)
    d = c["q0"].shape[0]
    q_cur = (c["q_cur"].reshape(-1, d)[:1] + 0.2 * torch.randn(N, d, generator=gen)) if batch_start else c["q_cur"].reshape(-1)[:d]

    def fresh():
        m = factory.make_mppi(c, device="cuda", H=H, N=N, pass1="auto", copy_policy=False, q_cur=q_cur)
        g2 = torch.Generator().manual_seed(12)
        nk = int(c["nk"])
        P = m.Policy
        P.mu_tmp.zero_(); P.sigma_tmp.zero_(); P.alpha_tmp.zero_()
        if nk:
            P.mu_tmp[:, :nk] = (P.mu_c[:nk].cpu() + 0.05 * torch.randn(N, nk, d, generator=g2)).to(P.mu_tmp.device)
            P.sigma_tmp[:, :nk] = (P.sigma_c[:nk].cpu() * (1 + 0.1 * torch.rand(N, nk, generator=g2))).to(P.sigma_tmp.device)
            P.alpha_tmp[:, :nk] = (P.alpha_c[:nk].cpu() + 0.3 * torch.randn(N, nk, d, generator=g2)).to(P.alpha_tmp.device)
        return m

    m = fresh()
    host0 = _host_buffers(m, q_cur)
    traj, dist, kv, dots, acts = m.propagate()
    cost = m.get_cost()
    _, n_upd = m.shift_policy_means()
    want = dict(all_traj=traj, closest_dist_all=dist, dot_products=dots, kernel_activations=acts, qdot=m.qdot, cost=cost,
                mu_c=m.Policy.mu_c, sigma_c=m.Policy.sigma_c, alpha_c=m.Policy.alpha_c)
    for chunk in ("0", "64"):
        monkeypatch.setenv("DSMPPI_HOST_CHUNK", chunk)
        m2 = fresh()
        host = {k: v.clone().pin_memory() for k, v in host0.items()}
        m2.iteration_host(host)
        torch.cuda.synchronize()
        # bit patterns, so that NaN == NaN: the H = 1 field case has a 0/0 stagnation cost in both paths
        bits = lambda x: x.detach().cpu().contiguous().view(torch.int32)  # noqa: E731
        for k, v in want.items():
            assert torch.equal(bits(host[k]), bits(v)), f"chunk={chunk}: {k} differs from the device path"
        nk = int(c["nk"])
        assert torch.equal(bits(host["kernel_val_all"][:, :, :nk]), bits(kv[:, :, :nk])), f"chunk={chunk}: kernel_val_all"
        assert int(host["n_updated"][0]) == int(n_upd)


def test_parameter_vectors_on_the_device_are_read_at_call_time(factory):
    """Goal and joint limits handed over as CUDA tensors are cached on the host per (tensor, in-place version) so that
    a call does not synchronise with the device; an in-place edit or a rebinding must still be seen by the next call
    (the reference reads the attributes at call time)."""
    c = load_npz("case_planar7")
    m = factory.make_mppi(c, device="cuda", pass1="auto")
    m.propagate()
    m.Cost.q_max = torch.full_like(m.Cost.q_max, 100.0)          # rebound: nothing violates a joint limit
    m.Cost.q_min = torch.full_like(m.Cost.q_min, -100.0)
    base = m.get_cost().clone()
    assert torch.equal(m.get_cost(), base)                       # served from the cache: same numbers
    m.Cost.q_max[0] = -100.0                                     # in place: every trajectory now violates a limit
    assert torch.allclose(m.get_cost(), base + 100.0)
    m.Cost.q_max[0] = 100.0
    assert torch.equal(m.get_cost(), base)
    goal0 = m.DS.q_goal.clone()
    traj0 = m.propagate()[0].clone()
    m.DS.q_goal += 0.5                                           # in place: the nominal DS must follow
    assert not torch.equal(m.propagate()[0], traj0)
    m.DS.q_goal.copy_(goal0)
    assert torch.equal(m.propagate()[0], traj0)
