"""The tensor-core scoring kernel (csrc/tc_exact.cu: split-fp16 operands, fp32-accurate) against the strict IEEE-FFMA
kernel (csrc/exact_mlp.cu) on the same rows, through the public API / C ABI: every row-source mode, ragged sizes, and
the fp16 range guard (rows whose activations saturate are re-scored by the FFMA kernel)."""
import pytest
import torch

from tests.golden_util import load_npz

pytestmark = pytest.mark.gpu


def both_modes(m, q):
    out = {}
    for mode in ("ffma", "tc_split"):
        m.set_score_mode(mode)
        dist, grad = m.distance_repulsion_nn(q)
        out[mode] = (dist.clone().double().cpu(), grad.clone().double().cpu())
    return out


def rms(x):
    return x.pow(2).mean().sqrt().item()


@pytest.mark.parametrize("case,n,M,pass1", [
    ("planar2", 1, None, "exact"),            # one sample: a single, mostly padded tile
    ("planar2", 777, None, "exact"),          # M = 2: every pair differentiated in one launch (dense rows)
    ("planar7", 1000, None, "exact"),
    ("franka_shelf", 513, 40, "exact"),       # M > 16: forward-only launch on all pairs + forward/VJP on the K selected
    ("franka_shelf", 1500, 300, "tc_f16"),    # prefilter + device-counted candidate list
])
def test_tc_split_matches_ffma(case, n, M, pass1):
    from tests.mppi_factory import make_mppi
    torch.manual_seed(3)
    c = load_npz(f"case_{case}")
    if M is not None:
        obs = torch.rand(M, 4) * 1.2 - 0.6
        obs[:, 3] = 0.03
        c["obs"] = obs
    m = make_mppi(c, device="cuda", pass1=pass1)
    d = c["q0"].shape[0]
    q = ((torch.rand(n, d) * 2 - 1) * 2.5).cuda()
    o = both_modes(m, q)
    (d0, g0), (d1, g1) = o["ffma"], o["tc_split"]
    assert torch.isfinite(d1).all() and torch.isfinite(g1).all()
    # the two arithmetics agree to a few 1e-7 of the output scale; a sample whose K-th / (K+1)-th
    # obstacles or argmin links tie within that noise may pick the other one: at most 1 % of the samples
    scale_d = max(rms(d0), 1e-3)
    scale_g = max(rms(g0), 1e-3)
    ok = ((d1 - d0).abs() <= 1e-5 * d0.abs() + 5e-6 * max(scale_d, 1.0)) & \
         ((g1 - g0).abs().max(dim=1)[0] <= 2e-4 * scale_g)
    assert ok.double().mean().item() >= 0.99 or n == 1 and bool(ok.all()), \
        f"only {ok.double().mean().item():.4f} of the samples agree"
    assert rms((d1 - d0)[ok]) <= 2e-6 * max(scale_d, 1.0)
    assert rms((g1 - g0)[ok]) <= 2e-5 * scale_g
    st = m.score_stats()
    assert st["mode"] == "tc_split" and st["dropped_rows"] == 0


def test_tc_split_pass1_forward_only_matches_ffma():
    from tests.mppi_factory import make_mppi
    torch.manual_seed(4)
    c = load_npz("case_franka_shelf")
    m = make_mppi(c, device="cuda", pass1="exact")
    q = (c["q0"] + 0.5 * torch.randn(300, 7)).cuda()        # 300 x 294 rows: ragged last tile
    m.set_score_mode("ffma")
    a = m.debug_pass1(q, "exact").double().cpu()
    m.set_score_mode("tc_split")
    b = m.debug_pass1(q, "exact").double().cpu()
    err = (a - b).abs()
    finite = a.abs() < 1e5                                  # ignored links are the constant 1e6
    assert rms(err[finite]) <= 1e-6 and err[finite].max().item() <= 2e-5


def test_fp16_range_guard_rescored_by_ffma():
    """Inputs far outside the training range blow the hidden activations past 65504 (the learned net extrapolates
    to 1e4 .. 1e5): the tensor-core kernel must flag those rows and the FFMA kernel re-score them, so the result is the
    strict mode's -- not a saturated one."""
    from tests.mppi_factory import make_mppi
    torch.manual_seed(5)
    c = load_npz("case_planar2")
    c["obs"] = torch.tensor([[60.0, 0.0, 0.0, 0.5], [0.0, 45.0, 0.0, 0.5]])
    m = make_mppi(c, device="cuda", pass1="exact")
    n = 300
    q = (torch.rand(n, 2) * 2 - 1) * 3.0
    q[::7] *= 4.0e4                                         # every 7th state itself beyond the fp16 range
    q = q.cuda()
    o = both_modes(m, q)
    (d0, g0), (d1, g1) = o["ffma"], o["tc_split"]
    st = m.score_stats()
    assert st["range_fixup_rows"] > 0 and st["dropped_rows"] == 0, st
    assert torch.isfinite(d1).all() and torch.isfinite(g1).all()
    # the states beyond the fp16 range come back from the FFMA kernel bit for bit; the rest within the usual tolerance
    far = (q.abs().max(dim=1)[0] > 6.6e4).cpu()
    assert far.any()
    assert torch.equal(d1[far], d0[far]) and torch.equal(g1[far], g0[far])
    assert ((d1 - d0).abs() <= 2e-5 * d0.abs() + 1e-5).double().mean().item() >= 0.99
    # and a whole rollout from such states stays finite and equal to the strict mode's where it saturates
    outs = {}
    for mode in ("ffma", "tc_split"):
        m.set_score_mode(mode)
        m.q_cur = q[:int(c["N"])].clone()
        outs[mode] = [x.clone() for x in m.propagate()]
    assert all(torch.isfinite(x).all() for x in outs["tc_split"])
