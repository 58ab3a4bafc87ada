"""Builds product MPPI objects (optimalmodulationds_b200) from golden-case dictionaries."""
import torch

from optimalmodulationds_b200 import MPPI, LinDS
from optimalmodulationds_b200.sdf.robot_sdf import RobotSdfCollisionNet
from tests.golden_util import full_policy, load_weights

NET_SHAPES = {"planar2": (2, 2), "planar7": (7, 7), "franka": (7, 9)}
# scoring arithmetic of the objects built below: None = the library default (tensor-core split-fp16 when the network
# fits), "ffma" = strict IEEE fp32 on the CUDA cores; test modules parametrise it through the `score_mode` fixture
DEFAULT_SCORE = None


def make_net(name):
    dof, out = NET_SHAPES[name]
    net = RobotSdfCollisionNet(in_channels=dof + 3, out_channels=out, layers=[256] * 4, skips=[])
    W, b = load_weights(name)
    net.load_arrays(W, b)
    return net


def make_mppi(c, device="cpu", H=None, N=None, q_cur=None, pass1="exact", copy_policy=True):
    """device: where the CALLER's tensors live ('cpu' like the reference's scripts, or 'cuda')."""
    dev = torch.device(device)
    t = lambda x: x.to(dev)  # noqa: E731
    N = int(c["N"]) if N is None else N
    H = int(c["H"]) if H is None else H
    net = make_net(c["net"])
    DS = [LinDS(t(c["qf"])), LinDS(t(c["q0"]))]
    m = MPPI(t(c["q0"]), t(c["qf"]), t(c["dh_params"]), t(c["obs"]), float(c["dt"]), H, N, DS, t(c["dh_a"]), net,
             int(c["K"]))
    m.set_pass1_mode(pass1)
    if DEFAULT_SCORE is not None:
        m.set_score_mode(DEFAULT_SCORE)
    m.Policy.p = float(c["p"])
    m.dst_thr = float(c["dst_thr"])
    m.ker_thr = float(c["ker_thr"])
    m.ignored_links = c["ignored_links"].tolist()
    m.Cost.q_min, m.Cost.q_max = t(c["q_min"]), t(c["q_max"])
    nk = int(c["nk"])
    P = m.Policy
    P.n_kernels = nk
    P.mu_c.copy_(t(c["mu_c0"])); P.sigma_c.copy_(t(c["sigma_c0"])); P.alpha_c.copy_(t(c["alpha_c0"]))
    if copy_policy and N == int(c["N"]):
        P.mu_tmp.copy_(t(full_policy(c, "mu_tmp", N)))
        P.sigma_tmp.copy_(t(full_policy(c, "sigma_tmp", N)))
        P.alpha_tmp.copy_(t(full_policy(c, "alpha_tmp", N)))
    m.q_cur = t(c["q_cur"]) if q_cur is None else t(q_cur)
    return m


def make_toy_net():
    net = RobotSdfCollisionNet(in_channels=4, out_channels=1, layers=[256] * 4, skips=[])
    W, b = load_weights("toy2")
    net.load_arrays(W, b)
    return net


def make_toy_mppi(c, device="cpu", H=None, N=None, pass1="exact"):
    """The toy variant (optimalmodulationds_b200.MPPI_toy.MPPI) from a toycase_* golden dictionary, driven like
    standaloneToy2d.py:83-91."""
    from optimalmodulationds_b200.MPPI_toy import MPPI as ToyMPPI
    dev = torch.device(device)
    t = lambda x: x.to(dev)  # noqa: E731
    N = int(c["N"]) if N is None else N
    H = int(c["H"]) if H is None else H
    m = ToyMPPI(t(c["q0"]), t(c["qf"]), torch.zeros(4, 4), t(c["obs"]), float(c["dt"]), H, N, t(c["A"]), 0,
                make_toy_net(), int(c["K"]))
    m.set_pass1_mode(pass1)
    if DEFAULT_SCORE is not None:
        m.set_score_mode(DEFAULT_SCORE)
    m.Policy.p = float(c["p"])
    m.dst_thr = float(c["dst_thr"])
    m.ker_thr = float(c["ker_thr"])
    m.policy_upd_rate = float(c["upd_rate"])
    m.ignored_links = []
    d = c["q0"].shape[0]
    m.Cost.q_min, m.Cost.q_max = -10 * torch.ones(d, device=dev), 10 * torch.ones(d, device=dev)
    nk = int(c["nk"])
    P = m.Policy
    P.n_kernels = nk
    P.mu_c.copy_(t(c["mu_c0"])); P.sigma_c.copy_(t(c["sigma_c0"])); P.alpha_c.copy_(t(c["alpha_c0"]))
    if N == int(c["N"]):
        P.mu_tmp.copy_(t(full_policy(c, "mu_tmp", N)))
        P.sigma_tmp.copy_(t(full_policy(c, "sigma_tmp", N)))
        P.alpha_tmp.copy_(t(full_policy(c, "alpha_tmp", N)))
    m.q_cur = t(c["q_cur"])
    return m
