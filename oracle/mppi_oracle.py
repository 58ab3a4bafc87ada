"""CPU oracle for the ds_mppi MPPI rollout hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A from-scratch fp32 restatement (torch CPU tensors, vectorised over samples) of the reference's
algorithm.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module, and only as the checker / reported CPU baseline.  The product path
(`optimalmodulationds_b200`) never imports it and has no CPU fallback.

Parity pinning: the reference ships no tests or golden vectors for this path, so
this oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, executed in the build container by
`tests/golden/make_golden.py` (unmodified reference classes + the two import shims of SURVEY 8(c)) and
committed as `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function below against
them, plus the known-answer vectors of SURVEY.md section 4.

All `file:line` citations are relative to /root/reference/python_scripts/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch

N_KERNEL_MAX = 50          # ds_mppi/functions/policy.py:18


# ----------------------------------------------------------------------------------------------
# Learned distance network  (mlp_learn/sdf/network_macros_mod.py:137-146, robot_sdf.py:153-158)
# ----------------------------------------------------------------------------------------------
@dataclass
class Net:
    """Weights of the shipped 3(d+3) -> 256 x4 (ReLU) -> O network; W[l]: (out, in), b[l]: (out,)."""
    W: List[torch.Tensor]
    b: List[torch.Tensor]

    @property
    def n_in(self) -> int:            # d + 3
        return self.W[0].shape[1] // 3

    @property
    def n_out(self) -> int:
        return self.W[-1].shape[0]


def encode(x: torch.Tensor) -> torch.Tensor:
    """enc(x) = [x, sin x, cos x]   (network_macros_mod.py:139-140)."""
    return torch.cat((x, torch.sin(x), torch.cos(x)), dim=-1)


def mlp_forward(net: Net, x: torch.Tensor) -> torch.Tensor:
    """Raw network output for rows x = [q, p]  (network_macros_mod.py:137-146, skips=[])."""
    h = encode(x)
    n_layers = len(net.W)
    for l in range(n_layers - 1):
        h = torch.relu(torch.nn.functional.linear(h, net.W[l], net.b[l]))
    return torch.nn.functional.linear(h, net.W[-1], net.b[-1])


def mlp_forward_grad(net: Net, x: torch.Tensor):
    """(z, dz[l*]/dx, l*) with l* = argmin_l z[l]  -- what functorch_vjp returns (robot_sdf.py:153-158).

    Analytic VJP through the ReLU masks and the sin/cos encoding:
      a = W1^T(s1 * W2^T(s2 * W3^T(s3 * W4^T(s4 * W5[l*,:]))));  a = [a0,a1,a2]
      dz/dx = a0 + cos(x)*a1 - sin(x)*a2
    """
    enc = encode(x)
    h = enc
    masks = []
    for l in range(len(net.W) - 1):
        pre = torch.nn.functional.linear(h, net.W[l], net.b[l])
        masks.append(pre > 0)
        h = torch.relu(pre)
    z = torch.nn.functional.linear(h, net.W[-1], net.b[-1])
    lstar = torch.argmin(z, dim=1)
    g = net.W[-1][lstar, :]                                   # (rows, 256)
    for l in range(len(net.W) - 2, -1, -1):
        g = g * masks[l]
        g = g @ net.W[l]                                      # W_l^T g
    n = x.shape[1]
    a0, a1, a2 = g[:, :n], g[:, n:2 * n], g[:, 2 * n:]
    grad = a0 + torch.cos(x) * a1 - torch.sin(x) * a2
    return z, grad, lstar


# ----------------------------------------------------------------------------------------------
# Nominal DS  (ds_mppi/functions/LinDS.py:11-21)
# ----------------------------------------------------------------------------------------------
def lin_ds_velocity(q: torch.Tensor, q_goal: torch.Tensor, lin_thr: float = 0.015) -> torch.Tensor:
    x = q - q_goal
    dst = x.norm(p=2, dim=-1)
    v = -x
    vn = v.norm(p=2, dim=-1, keepdim=True)
    far = dst > lin_thr
    v = torch.where(far.unsqueeze(-1), v / vn, v)
    return v


@dataclass
class SedsParams:
    """Arrays of a SEDS model as the MATLAB files store them (ds_mppi/functions/SEDS.py:12-18)."""
    Mu: torch.Tensor        # (2d, G)
    Sigma: torch.Tensor     # (2d, 2d, G)
    Priors: torch.Tensor    # (G,) or (G, 1)
    seds_thr: float = 1e-2  # SEDS.py:26
    lin_thr: float = 1e-2   # SEDS.py:27


def seds_velocity(x: torch.Tensor, q_goal: torch.Tensor, sp: SedsParams) -> torch.Tensor:
    """SEDS.get_velocity (SEDS.py:61-76) with gaussPDF (:28-34) and GMR (:36-59), restated per state (the
    reference's own batched indexing at :70-71 only broadcasts for a single state).  x: (n, d) -> (n, d)."""
    d = sp.Mu.shape[0] // 2
    G = sp.Sigma.shape[2]
    pri = sp.Priors.reshape(-1)
    out = torch.zeros_like(x)
    inv = [torch.inverse(sp.Sigma[:d, :d, j]) for j in range(G)]
    det = [torch.abs(torch.det(sp.Sigma[:d, :d, j])) for j in range(G)]
    for i in range(x.shape[0]):
        xd = x[i] - q_goal.reshape(-1)                                              # :62
        pxi = torch.zeros(G)
        for j in range(G):
            D = (xd - sp.Mu[:d, j]).reshape(1, d)
            quad = torch.sum((D @ inv[j]) * D, dim=1)
            pxi[j] = pri[j] * (torch.exp(-0.5 * quad) /
                               torch.sqrt((2 * torch.tensor(torch.pi) ** d) * det[j] + torch.tensor(1e-100)))[0]
        beta = torch.clamp((pxi / pxi.sum()).nan_to_num(), min=1e-8)                # :50-52
        y = torch.zeros(d)
        for j in range(G):
            y = y + beta[j] * (sp.Mu[d:, j] + sp.Sigma[d:, :d, j] @ inv[j] @ (xd - sp.Mu[:d, j]))   # :53-59
        dst, y_norm = xd.norm(), y.norm()
        if dst > sp.lin_thr:                                                        # :63,68-71
            y = (-xd / dst) if y_norm < sp.seds_thr else y / y_norm                 # :72-75
        out[i] = y
    return out


def generalized_sigmoid(x, y_min, y_max, x0, x1, k):
    """ds_mppi/functions/MPPI.py:352-353."""
    return y_min + (y_max - y_min) / (1 + torch.exp(k * (-x + (x0 + x1) / 2)))


# ----------------------------------------------------------------------------------------------
# Distance + repulsion gradient  (ds_mppi/functions/MPPI.py:227-282)
# ----------------------------------------------------------------------------------------------
def distance_repulsion(net: Net, q: torch.Tensor, obs: torch.Tensor, n_closest: int,
                       ignored_links: Sequence[int], return_aux: bool = False):
    """q: (n, d), obs: (M, 4)=[x,y,z,r]  ->  distance (n,), nn_grad (n, d).

    Pass 1 (MPPI.py:231-253): raw MLP on all (obstacle, sample) pairs, /100 when O == 9, minus radius,
    ignored links := 1e6, min over links, ascending sort over obstacles, first K.
    Pass 2 (MPPI.py:255-280, robot_sdf.py:153-158): forward + VJP at argmin_l of the RAW output (no
    radius, no ignore mask) on the K selected pairs; softmax(-10 * dist) blend of the K gradients;
    returned distance = rank-0 pair's dist.
    """
    n, d = q.shape
    M = obs.shape[0]
    K = n_closest
    O = net.n_out
    scale = 0.01 if O == 9 else 1.0
    # row j*n + i = [q_i, obs_j]   (build_nn_input, MPPI.py:93-95)
    x_all = torch.hstack((q.tile(M, 1), obs.repeat_interleave(n, 0)))
    y = mlp_forward(net, x_all[:, :-1])
    if O == 9:
        y = y / 100
    y = y - x_all[:, -1:].clone()
    if len(ignored_links) > 0:
        y[:, list(ignored_links)] = 1e6
    mind = y.min(1)[0].reshape(M, n).transpose(0, 1)          # (n, M)
    sort_d, sort_idx = mind.sort(dim=1)
    sel = sort_idx[:, :K]                                     # (n, K)
    rows = (torch.arange(n).unsqueeze(1) + sel * n).flatten()
    x_sel = x_all[rows]                                       # (n*K, d+4), sample-major
    z, grad, lstar = mlp_forward_grad(net, x_sel[:, :-1])
    if O == 9:
        z = z / 100
    z = z - x_sel[:, -1:]
    dsel = z[torch.arange(n * K), lstar].reshape(n, K)
    g = grad[:, :d].reshape(n, K, d)
    w = (-10 * dsel).softmax(dim=-1)
    nn_grad = (g * w.unsqueeze(2)).sum(dim=1)
    distance = dsel[:, 0].clone()
    if return_aux:
        return distance, nn_grad, dict(mind=mind, sel=sel, dsel=dsel, grads=g, lstar=lstar.reshape(n, K))
    return distance, nn_grad


def distance_repulsion_fk(q: torch.Tensor, obs: torch.Tensor, dh_params: torch.Tensor, n_pts: int = 10,
                          return_idx: bool = False):
    """MPPI.distance_repulsion_fk (MPPI.py:306-313): n_pts points on every link of the modified-DH chain
    (numeric_fk_model_vec, fk_num.py:78-89), sphere distances minimised over points, obstacles, links in that
    nesting (dist_tens, fk_num.py:142-160), and the gradient of the minimal distance w.r.t. the joints.  The
    reference evaluates sympy closed forms of that gradient for the planar chain (fk_sym_gen.py:r1..r7,
    lambda_rep_vec :265-269); this restatement uses the geometric Jacobian dP/dq_j = z_j x (P - o_j), the same
    derivative.  q: (n, d), obs: (M, 4) -> distance (n,), gradient (n, d)."""
    n, d = q.shape
    span = torch.linspace(0.01, 1, n_pts)
    T = torch.eye(4).repeat(n, 1, 1)
    frames = []
    pts = torch.zeros(n, d, n_pts, 3)
    for i in range(d):
        T = T @ dh_transform(q[:, i], dh_params[i, 0], dh_params[i, 1], dh_params[i, 2], dh_params[i, 3])
        frames.append(T)
        local = torch.zeros(n_pts, 3)
        local[:, 0] = dh_params[i + 1, 2] * span
        pts[:, i] = torch.einsum('nrc,pc->npr', T[:, :3, :3], local) + T[:, None, :3, 3]
    dst = torch.norm(pts.unsqueeze(2) - obs[:, :3].reshape(1, 1, -1, 1, 3), 2, 4) - obs[:, 3].reshape(1, 1, -1, 1)
    mind_pts, idx_pts = torch.min(dst, 3)                     # (n, d, M)
    mind_obs, idx_obs = torch.min(mind_pts, 2)                # (n, d)
    mind, idx_link = torch.min(mind_obs, 1)                   # (n,)
    ar = torch.arange(n)
    j = idx_obs[ar, idx_link]
    p = idx_pts[ar, idx_link, j]
    P = pts[ar, idx_link, p]                                  # (n, 3)
    u = P - obs[j, :3]
    u = u / u.norm(dim=1, keepdim=True)
    grad = torch.zeros(n, d)
    for k in range(d):
        z = frames[k][:, :3, 2]
        o = frames[k][:, :3, 3]
        dP = torch.cross(z, P - o, dim=1)
        grad[:, k] = torch.where(idx_link >= k, (u * dP).sum(1), torch.zeros(n))
    if return_idx:
        return mind, grad, torch.stack((j, idx_link, p), 1)
    return mind, grad


# ----------------------------------------------------------------------------------------------
# Householder basis  (torch.linalg.qr on [g | e_2 .. e_d], MPPI.py:122-127)
# ----------------------------------------------------------------------------------------------
def householder_basis(g: torch.Tensor) -> torch.Tensor:
    """E = Q of the unblocked LAPACK Householder QR (geqr2 + org2r) of A = I with column 0 := g,
    then column 0 overwritten by g/|g|  (MPPI.py:122-126).  g: (n, d) -> (n, d, d).

    Because A[:, 1:] = I[:, 1:], only the first reflector is non-trivial up to sign: this routine still
    runs the full d-step algorithm so it is an independent check of the closed form used on the GPU.
    """
    n, d = g.shape
    A = torch.eye(d, dtype=g.dtype).repeat(n, 1, 1)
    A[:, :, 0] = g
    taus = torch.zeros(n, d, dtype=g.dtype)
    for k in range(d):
        # slarfg on A[k:, k]
        alpha = A[:, k, k].clone()
        xnorm = A[:, k + 1:, k].norm(dim=1) if k + 1 < d else torch.zeros(n, dtype=g.dtype)
        beta = -torch.copysign(torch.sqrt(alpha * alpha + xnorm * xnorm), alpha)
        trivial = xnorm == 0
        tau = torch.where(trivial, torch.zeros_like(alpha), (beta - alpha) / beta)
        scal = torch.where(trivial, torch.zeros_like(alpha), 1.0 / (alpha - beta))
        if k + 1 < d:
            A[:, k + 1:, k] = A[:, k + 1:, k] * scal.unsqueeze(1)
        A[:, k, k] = torch.where(trivial, alpha, beta)
        taus[:, k] = tau
        # apply H_k = I - tau v v^T to A[k:, k+1:]
        if k + 1 < d:
            v = torch.cat((torch.ones(n, 1, dtype=g.dtype), A[:, k + 1:, k]), dim=1)   # (n, d-k)
            sub = A[:, k:, k + 1:]
            wv = torch.einsum('ni,nij->nj', v, sub)
            A[:, k:, k + 1:] = sub - taus[:, k].view(n, 1, 1) * v.unsqueeze(2) * wv.unsqueeze(1)
    # org2r: Q = H_0 H_1 ... H_{d-1}
    Q = torch.eye(d, dtype=g.dtype).repeat(n, 1, 1)
    for k in range(d - 1, -1, -1):
        v = torch.zeros(n, d, dtype=g.dtype)
        v[:, k] = 1
        if k + 1 < d:
            v[:, k + 1:] = A[:, k + 1:, k]
        wv = torch.einsum('ni,nij->nj', v, Q)
        Q = Q - taus[:, k].view(n, 1, 1) * v.unsqueeze(2) * wv.unsqueeze(1)
    Q[:, :, 0] = g / g.norm(2, 1).unsqueeze(1)
    return Q


# ----------------------------------------------------------------------------------------------
# RBF policy  (ds_mppi/functions/policy.py:186-199)
# ----------------------------------------------------------------------------------------------
def eval_rbf(q: torch.Tensor, mu: torch.Tensor, sigma: torch.Tensor, p=2) -> torch.Tensor:
    num = torch.norm(q[:, None, :] - mu, p=p, dim=2, keepdim=True) ** 2
    return torch.exp(-sigma.unsqueeze(2) * num)


# ----------------------------------------------------------------------------------------------
# One rollout  (ds_mppi/functions/MPPI.py:97-224)
# ----------------------------------------------------------------------------------------------
@dataclass
class RolloutParams:
    dt: float
    dt_H: int
    n_closest_obs: int
    dst_thr: float = 0.5                     # MPPI.py:59
    ignored_links: Sequence[int] = (0, 1, 2) # MPPI.py:62
    lin_thr: float = 0.015                   # LinDS.py:9
    p: float = 2                             # policy.py:41
    with_basis: bool = True                  # materialise norm_basis (Householder E)
    explicit_M: bool = True                  # M = E D E^T like MPPI.py:158-161 (needs with_basis);
                                             # False: rank-1 closed form, what the GPU uses
    # constants the reference compiles into MPPI.py; toy_params() returns the MPPI_toy.py set
    lvel: Sequence[float] = (-1, 0, 10)      # (x0, x1, k) of l_vel                        MPPI.py:132
    dist_sig: Sequence[float] = (0.0, 0.1, 100)   # (dist_low, dist_high, k_sigmoid)       MPPI.py:148-151
    ltau_max: float = 5                      # MPPI.py:152
    goal_act_thr: float = 0.5                # MPPI.py:194
    repulsion: float = 0.1                   # MPPI.py:216
    fold_activation: bool = False            # kernel_val_all *= activation                MPPI_toy.py:178-179
    A: Optional[torch.Tensor] = None         # matrix DS v = (q - qf) @ A                  MPPI_toy.py:89
    fk_dh_params: Optional[torch.Tensor] = None   # set: distances from distance_repulsion_fk (MPPI.py:115)
    seds: Optional["SedsParams"] = None      # set: nominal DS = SEDS.get_velocity (MPPI.py:106 with a SEDS object)


def toy_params(dt, dt_H, n_closest_obs, A, **kw):
    """The constant set of ds_mppi/functions/MPPI_toy.py (:56,89,114,124-133,176-179,199)."""
    base = dict(dst_thr=0.1, ignored_links=(), lvel=(-0.2, 0.0, 100), dist_sig=(0.0, 0.5, 30), ltau_max=3,
                goal_act_thr=0.3, repulsion=0.05, fold_activation=True, A=A)
    base.update(kw)
    return RolloutParams(dt=dt, dt_H=dt_H, n_closest_obs=n_closest_obs, **base)


@dataclass
class RolloutOut:
    all_traj: torch.Tensor
    closest_dist_all: torch.Tensor
    kernel_val_all: torch.Tensor             # (N, H, 50) -- live columns [:nk]
    dot_products: torch.Tensor
    kernel_activations: torch.Tensor
    qdot: torch.Tensor
    nn_grad_all: torch.Tensor                # (N, H, d): blended gradient per state-step
    norm_basis: Optional[torch.Tensor] = None
    step_aux: list = field(default_factory=list)


def rollout(net: Net, q_cur: torch.Tensor, q_goal: torch.Tensor, obs: torch.Tensor,
            mu_tmp: torch.Tensor, sigma_tmp: torch.Tensor, alpha_tmp: torch.Tensor, n_kernels: int,
            prm: RolloutParams, N: int, keep_aux: bool = False) -> RolloutOut:
    """Restates MPPI.propagate (MPPI.py:97-224).  q_cur: (d,) or (N, d); *_tmp: (N, 50, ..)."""
    d = q_goal.shape[-1]
    H = prm.dt_H
    nk = n_kernels
    all_traj = torch.zeros(N, H, d)
    closest = 100 + torch.zeros(N, H)
    kval = torch.zeros(N, H, N_KERNEL_MAX)
    dots = torch.zeros(N, H)
    acts = torch.zeros(N, H)
    grads = torch.zeros(N, H, d)
    basis = torch.zeros(N, H, d, d) if prm.with_basis else None
    qdot = torch.zeros(N, d)
    all_traj[:, 0, :] = q_cur
    aux = []
    for i in range(1, H + 1):
        q = all_traj[:, i - 1, :]
        if prm.seds is not None:
            v = seds_velocity(q, q_goal, prm.seds)                        # MPPI.py:106, SEDS.py:61-76
        elif prm.A is not None:
            v = (q - q_goal) @ prm.A                                      # MPPI_toy.py:89
        else:
            v = lin_ds_velocity(q, q_goal, prm.lin_thr)                   # MPPI.py:106
        vn = v.norm(dim=1).reshape(-1, 1)                                 # :107
        vhat = v / vn                                                     # :108
        if prm.fk_dh_params is not None:
            dist, g = distance_repulsion_fk(q, obs, prm.fk_dh_params)
        elif keep_aux:
            dist, g, a = distance_repulsion(net, q, obs, prm.n_closest_obs, prm.ignored_links, True)
            aux.append(a)
        else:
            dist, g = distance_repulsion(net, q, obs, prm.n_closest_obs, prm.ignored_links)
        dist = dist - prm.dst_thr                                         # :117
        closest[:, i - 1] = dist
        grads[:, i - 1] = g
        e0 = g / g.norm(2, 1).unsqueeze(1)                                # :126
        if prm.with_basis:
            basis[:, i - 1] = householder_basis(g)                        # :122-127
        dot = (e0 * vhat).sum(dim=-1)                                     # :129
        dots[:, i - 1] = dot
        d_lo, d_hi, d_k = prm.dist_sig
        l_vel = generalized_sigmoid(dot, 0, 1, *prm.lvel)                 # :132
        l_n = generalized_sigmoid(dist, 0, 1, d_lo, d_hi, d_k)            # :153
        l_n_vel = l_vel + (1 - l_vel) * l_n                               # :154
        l_tau = generalized_sigmoid(dist, prm.ltau_max, 1, d_lo, d_hi, d_k)   # :155
        if nk > 0:
            kv = eval_rbf(q, mu_tmp[:, :nk], sigma_tmp[:, :nk], prm.p)    # :165  (N, nk, 1)
            policy_value = torch.sum(alpha_tmp[:, :nk] * kv, 1)           # :174-177
            kval[:, i - 1, :nk] = kv.reshape(N, nk)                       # :184
        else:
            policy_value = v * 0                                          # :186
        goal_act = (q - q_goal).norm(p=0.5, dim=1).clamp(0, 1).unsqueeze(1)   # :193
        goal_act = torch.where(goal_act < prm.goal_act_thr, torch.zeros_like(goal_act), goal_act)  # :194
        act = (1 - l_n[:, None]) * (1 - l_vel[:, None]) * goal_act        # :191-195
        acts[:, i - 1] = act.squeeze(1)
        if prm.fold_activation and nk > 0:
            kval[:, i - 1, :nk] *= act                                    # MPPI_toy.py:178-179
        vt = v + act * policy_value * vn                                  # :197-206
        if prm.with_basis and prm.explicit_M:
            E = basis[:, i - 1]
            D = l_tau.repeat_interleave(d).reshape((N, d)).diag_embed(0, 1, 2)   # :158
            D[:, 0, 0] = l_n_vel                                          # :159
            Mmat = E @ D @ E.transpose(1, 2)                              # :161
            m = (Mmat @ vt.unsqueeze(2)).squeeze(2)                       # :209
        else:
            # E orthonormal with column 0 = e0  =>  M = l_tau I + (l_n_vel - l_tau) e0 e0^T
            m = l_tau[:, None] * vt + (l_n_vel - l_tau)[:, None] * e0 * (e0 * vt).sum(1, keepdim=True)
        mnorm = m.norm(dim=-1).reshape(-1, 1)                             # :211
        mnorm = torch.where(mnorm <= 0.5, torch.ones_like(mnorm), mnorm)  # :212
        m = torch.nan_to_num(m / mnorm)                                   # :213
        coll = (dist < 0).unsqueeze(1)
        m = torch.where(coll, 0.1 * m + e0 * vn * prm.repulsion, m)       # :215-217
        if i < H:
            all_traj[:, i, :] = q + prm.dt * m                            # :220-221
        if i == 1:
            qdot = m                                                      # :222-223
    return RolloutOut(all_traj, closest, kval, dots, acts, qdot, grads, basis, aux)


# ----------------------------------------------------------------------------------------------
# Forward kinematics + cost  (ds_mppi/functions/fk_num.py:7-75, cost.py:13-46)
# ----------------------------------------------------------------------------------------------
def kernel_candidates(all_traj, closest_dist_all, dot_products, mu_c, sigma_c, n_kernels, thr_dist, thr_kernel,
                      thr_dot, p=2):
    """TensorPolicyMPPI.check_traj_for_kernels (ds_mppi/functions/policy.py:153-175): state-steps close to an
    obstacle and moving into it, not covered by any kernel; returned in (sample, step) order."""
    near = (closest_dist_all < thr_dist) & (dot_products < thr_dot)
    cand = all_traj[near].reshape(-1, all_traj.shape[-1])
    if n_kernels > 0:
        dist2 = torch.norm(cand[:, None, :] - mu_c[:n_kernels], p, -1) ** 2            # policy.py:201-214
        cover = torch.exp(-sigma_c[:n_kernels] * dist2).max(dim=-1)[0]
        cand = cand[cover < thr_kernel]
    return cand


def dh_transform(q, d, theta, a, alpha):
    """Modified-DH transform, batched over q (n,) -> (n, 4, 4)   (fk_num.py:7-27)."""
    sa, ca = torch.sin(alpha), torch.cos(alpha)
    sq, cq = torch.sin(q + theta), torch.cos(q + theta)
    zero, one = torch.zeros_like(q), torch.ones_like(q)
    rows = [torch.stack((cq, -sq, zero, a + zero), -1),
            torch.stack((sq * ca, cq * ca, -sa + zero, -d * sa + zero), -1),
            torch.stack((sq * sa, cq * sa, ca + zero, d * ca + zero), -1),
            torch.stack((zero, zero, zero, one), -1)]
    return torch.stack(rows, -2)


def link_end_points(q: torch.Tensor, dh_params: torch.Tensor) -> torch.Tensor:
    """P_link(q), link = 0..d-1: last sample point of numeric_fk_model (fk_num.py:50-75, lspan end = 1):
    T_{link+1}[:3,3] + T_{link+1}[:3,0] * a_{link+1}.  q: (n, d) -> (n, d, 3)."""
    n, d = q.shape
    T = torch.eye(4).repeat(n, 1, 1)
    pts = []
    for i in range(d):
        T = T @ dh_transform(q[:, i], dh_params[i, 0], dh_params[i, 1], dh_params[i, 2], dh_params[i, 3])
        pts.append(T[:, :3, 3] + T[:, :3, 0] * dh_params[i + 1, 2])
    return torch.stack(pts, 1)


def evaluate_costs(all_traj, closest_dist_all, q_goal, dh_params, q_min, q_max):
    """Cost.evaluate_costs (cost.py:13-22); rest_cost is computed but unused there."""
    q_T = all_traj[:, -1, :]
    goal = 10 * (q_T - q_goal).norm(p=2, dim=1)                           # :14,24-25
    coll = 100 * (closest_dist_all < 0).sum(dim=1)                        # :15,33-34
    mask = ((all_traj < q_min).sum(dim=1) + (all_traj > q_max).sum(dim=1)).sum(dim=1)
    jl = 100 * ((mask > 0) + 0)                                           # :16,36-39
    dist = (all_traj[:, 0, :] - q_T).norm(2, dim=1)
    stag = 10 * goal * (1 / dist).nan_to_num(0)                           # :17,41-43
    goal_fk = link_end_points(q_goal.reshape(1, -1), dh_params)[0]        # :9
    fk = 10 * (link_end_points(q_T, dh_params) - goal_fk).norm(2, dim=2).sum(dim=1)   # :19,27-31
    return goal + coll + jl + stag + fk


def evaluate_costs_toy(all_traj, closest_dist_all, q_goal):
    """cost_toy.py:13-19: goal + collision + stagnation only."""
    q_T = all_traj[:, -1, :]
    goal = 10 * (q_T - q_goal).norm(p=2, dim=1)                           # :14
    coll = 100 * (closest_dist_all < 0).sum(dim=1)                        # :15
    dist = (all_traj[:, 0, :] - q_T).norm(2, dim=1)
    stag = 10 * goal * (1 / dist).nan_to_num(0)                           # :16
    return goal + coll + stag


# ----------------------------------------------------------------------------------------------
# Policy update  (ds_mppi/functions/MPPI.py:331-345, policy.py:88-113)
# ----------------------------------------------------------------------------------------------
def policy_update(cost, kernel_val_all, kernel_activations, mu_tmp, sigma_tmp, alpha_tmp,
                  mu_c, sigma_c, alpha_c, n_kernels, ker_thr, upd_rate=0.1, toy=False):
    """Returns (mu_c', sigma_c', alpha_c', n_updated, w).  toy=True: MPPI_toy.py:314-324 (max_t of the already
    activation-folded kernel values, no sample-0 base mask)."""
    nk = n_kernels
    beta = cost.mean() / 50
    w = torch.exp(-1 / beta * cost)
    w = w / w.sum()
    mu_c, sigma_c, alpha_c = mu_c.clone(), sigma_c.clone(), alpha_c.clone()
    if nk == 0:
        return mu_c, sigma_c, alpha_c, 0, w
    if toy:
        mask = kernel_val_all[:, :, :nk].max(dim=1)[0].mean(dim=0) > ker_thr          # MPPI_toy.py:318-320
    else:
        max_act = (kernel_val_all[:, :, :nk] * kernel_activations.unsqueeze(-1)).max(dim=1)[0]
        mask = max_act.mean(dim=0) > ker_thr
        mask = (kernel_val_all[0, :, :nk].mean(dim=0) > ker_thr) * mask
    mu_sum = torch.sum(w[:, None, None] * mu_tmp[:, :nk], 0)
    sigma_sum = torch.sum(w[:, None] * sigma_tmp[:, :nk], 0)
    alpha_sum = torch.sum(w[:, None, None] * alpha_tmp[:, :nk], 0)
    upd = upd_rate * torch.ones(nk)
    upd[~mask] = 0.0
    mu_c[:nk] = (1 - upd[:, None]) * mu_c[:nk] + upd[:, None] * mu_sum
    sigma_c[:nk] = (1 - upd) * sigma_c[:nk] + upd * sigma_sum
    alpha_c[:nk] = (1 - upd[:, None]) * alpha_c[:nk] + upd[:, None] * alpha_sum
    return mu_c, sigma_c, alpha_c, int(mask.sum()), w


def policy_update_partials(cost, kernel_val_all, kernel_activations, mu_tmp, sigma_tmp, alpha_tmp,
                           n_kernels, beta, owns_sample0: bool):
    """Per-shard partial sums of the policy update: the packed vector
    [sum w~, sum w~*mu (nk*d), sum w~*sigma (nk), sum w~*alpha (nk*d), sum_i max_t(kv*act) (nk),
     mean_t kv[0] (nk, zeros unless this shard owns global sample 0)] with w~ = exp(-cost/beta)
    (un-normalised).  Summing these over shards and calling policy_update_finalize reproduces
    policy_update()."""
    nk = n_kernels
    d = mu_tmp.shape[-1]
    w = torch.exp(-1 / beta * cost)
    parts = [w.sum().reshape(1),
             torch.sum(w[:, None, None] * mu_tmp[:, :nk], 0).reshape(-1),
             torch.sum(w[:, None] * sigma_tmp[:, :nk], 0).reshape(-1),
             torch.sum(w[:, None, None] * alpha_tmp[:, :nk], 0).reshape(-1),
             (kernel_val_all[:, :, :nk] * kernel_activations.unsqueeze(-1)).max(dim=1)[0].sum(0),
             kernel_val_all[0, :, :nk].mean(dim=0) if owns_sample0 else torch.zeros(nk)]
    return torch.cat(parts)
