"""Modified-DH forward kinematics helpers with the reference's names (ds_mppi/functions/fk_num.py).

Script-facing utilities (plotting a robot, kernel FK for the ZMQ message): vectorised torch, any device.
They are NOT on the MPPI hot path -- the terminal-FK cost runs inside the CUDA cost kernel
(csrc/rollout_kernels.cu: fk_points) -- and they pull in the plotting names exactly like the reference's
`from plots import *` so `from MPPI import *` exposes the same namespace.
"""
import numpy as np  # noqa: F401  (re-exported, scripts use `np` from the star import)
import torch

from .plots import *  # noqa: F401,F403


def dh_transform(q, d, theta, a, alpha):
    """Modified-DH link transform; q may be a scalar tensor or a batch (n,) -> (4,4) or (n,4,4)."""
    q, d, theta, a, alpha = (torch.as_tensor(v, dtype=torch.float32) for v in (q, d, theta, a, alpha))
    sa, ca = torch.sin(alpha), torch.cos(alpha)
    sq, cq = torch.sin(q + theta), torch.cos(q + theta)
    z, o = torch.zeros_like(q), torch.ones_like(q)
    rows = [torch.stack((cq, -sq, z, a + z), -1),
            torch.stack((sq * ca, cq * ca, -sa + z, -d * sa + z), -1),
            torch.stack((sq * sa, cq * sa, ca + z, d * ca + z), -1),
            torch.stack((z, z, z, o), -1)]
    return torch.stack(rows, -2)


def dh_fk(q, dh_params):
    """List of d+1 frames [I, T_1, .., T_d] for one configuration q (d,)."""
    T = [torch.eye(4, device=q.device)]
    for i in range(len(q)):
        T.append(T[-1] @ dh_transform(q[i], *dh_params[i].to(q.device)))
    return T


def numeric_fk_model_vec(q, dh_params, n_pts: int):
    """q (n, d) -> (link_pts (n, d, n_pts, 3), pts_int (n, d, n_pts, 3)): n_pts samples along every link,
    at fractions linspace(0.01, 1, n_pts) of the next link's `a`, in the world and in the link frame."""
    n, d = q.shape
    dh = dh_params.to(q.device, torch.float32)
    span = torch.linspace(0.01, 1, n_pts, device=q.device)
    T = torch.eye(4, device=q.device).repeat(n, 1, 1)
    link_pts = torch.zeros((n, d, n_pts, 3), device=q.device)
    pts_int = torch.zeros((n, d, n_pts, 3), device=q.device)
    for i in range(d):
        T = T @ dh_transform(q[:, i], dh[i, 0], dh[i, 1], dh[i, 2], dh[i, 3])
        local = torch.zeros(n_pts, 3, device=q.device)
        local[:, 0] = dh[i + 1, 2] * span
        link_pts[:, i] = torch.einsum('nrc,pc->npr', T[:, :3, :3], local) + T[:, None, :3, 3]
        pts_int[:, i] = local
    return link_pts, pts_int


def numeric_fk_model(q, dh_params, n_pts: int):
    link_pts, pts_int = numeric_fk_model_vec(q.reshape(1, -1), dh_params, n_pts)
    return link_pts[0], pts_int[0]
