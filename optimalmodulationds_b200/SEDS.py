"""SEDS nominal dynamical system: Gaussian-mixture regression of a velocity field learned from demonstrations.

Host-side mirror of the reference's ds_mppi/functions/SEDS.py:8-76 -- same constructor `SEDS(fname, attr=None)`
reading the MATLAB file's Mu / Sigma / Priors / xT, the same attributes (`Mu, Sigma, Priors, q_goal, dof,
n_gaussians, Sigma_inv, det, seds_thr, lin_thr`) and `get_velocity(x)`.  Inside the rollout the same law is
evaluated by the CUDA step kernel (csrc/step_device.cuh, DSMPPI_DS_SEDS) from the arrays `kernel_arrays()` packs;
`get_velocity` is a vectorised torch restatement kept for callers that query the DS directly.
"""
import math

import torch


class SEDS:
    def __init__(self, fname, attr=None):
        from scipy.io import loadmat
        data = loadmat(fname)
        self._init_from(data['Mu'], data['Sigma'], data['Priors'], data['xT'], attr)

    @classmethod
    def from_arrays(cls, Mu, Sigma, Priors, xT, attr=None):
        """Same object from in-memory arrays (Mu (2d, G), Sigma (2d, 2d, G), Priors (G, 1) or (G,), xT (d, 1))."""
        self = cls.__new__(cls)
        self._init_from(Mu, Sigma, Priors, xT, attr)
        return self

    def _init_from(self, Mu, Sigma, Priors, xT, attr):
        self.dtype = torch.float32
        self.Mu = torch.as_tensor(Mu).to(self.dtype)
        self.Sigma = torch.as_tensor(Sigma).to(self.dtype)
        self.Priors = torch.as_tensor(Priors).to(self.dtype)
        self.q_goal = torch.as_tensor(xT).to(self.dtype)
        if attr is not None:
            self.q_goal = attr
        self.dof = int(self.Mu.shape[0] / 2)
        self.n_gaussians = self.Sigma.shape[2]
        d, G = self.dof, self.n_gaussians
        self.Sigma_inv = torch.zeros([d, d, G], dtype=self.dtype)
        self.det = torch.zeros([G], dtype=self.dtype)
        for j in range(G):
            self.Sigma_inv[:, :, j] = torch.inverse(self.Sigma[:d, :d, j])
            self.det[j] = torch.abs(torch.det(self.Sigma[:d, :d, j]))
        self.seds_thr = 1e-2
        self.lin_thr = 1e-2

    def kernel_arrays(self):
        """The per-Gaussian arrays the CUDA step kernel consumes (include/dsmppi_b200.h: dsmppi_seds), computed with
        the same torch expressions the reference's gaussPDF / GMR use (SEDS.py:28-34,53-56)."""
        d, G = self.dof, self.n_gaussians
        pri = self.Priors.reshape(-1).contiguous()
        den = torch.sqrt((2 * torch.tensor(torch.pi) ** d) * self.det + torch.tensor(1e-100)).contiguous()
        mu_x = self.Mu[:d].t().contiguous()
        mu_y = self.Mu[d:].t().contiguous()
        sinv = torch.stack([torch.inverse(self.Sigma[:d, :d, j]) for j in range(G)]).contiguous()
        A = torch.stack([self.Sigma[d:, :d, j] @ self.Sigma_inv[:d, :d, j] for j in range(G)]).contiguous()
        return pri, den, mu_x, mu_y, sinv, A

    def GMR(self, x):
        """x: (d, n) offsets from the goal -> (d, n) regressed velocities."""
        pri, den, mu_x, mu_y, sinv, A = (t.to(x.device) for t in self.kernel_arrays())
        D = x.t().unsqueeze(0) - mu_x.unsqueeze(1)                       # (G, n, d)
        quad = ((D @ sinv) * D).sum(-1)                                  # (G, n)
        pxi = (pri.unsqueeze(1) * torch.exp(-0.5 * quad) / den.unsqueeze(1)).t()   # (n, G)
        beta = (pxi / pxi.sum(dim=1, keepdim=True)).nan_to_num().clamp(min=1e-8)
        y_each = mu_y.unsqueeze(1) + D @ A.transpose(1, 2)               # (G, n, d)
        return (beta.t().unsqueeze(-1) * y_each).sum(0).t()

    def get_velocity(self, x):
        goal = torch.as_tensor(self.q_goal).to(x.device).reshape(-1, 1)
        x_dif = x.transpose(-1, -2) - goal
        dst = x_dif.norm(p=2, dim=0)
        far = dst > self.lin_thr
        y = self.GMR(x_dif)
        y_norm = y.norm(p=2, dim=0)
        y_lin = -x_dif
        unit = torch.where(far, y_norm, torch.ones_like(y_norm))
        unit_lin = torch.where(far, dst, torch.ones_like(dst))
        y = y / unit
        y_lin = y_lin / unit_lin
        weak = (y_norm < self.seds_thr) & far
        y = torch.where(weak.unsqueeze(0), y_lin, y)
        return y[:self.dof, :].transpose(-2, -1)


__all__ = ["SEDS", "math"]
