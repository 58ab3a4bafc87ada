"""RBF-kernel navigation policy sampled by MPPI.

Host-side mirror of the reference's ds_mppi/functions/policy.py (class TensorPolicyMPPI :12-175, eval_rbf
:186-199, eval_rbf_simple :201-214): same attribute names, shapes and call semantics, so the reference's
scripts can rebind / poke them (standalonePlanar2d.py:181-188).  Sampling and kernel bookkeeping are
O(n_kernels) host logic and stay in torch; the per-state RBF evaluation inside the rollout and the weighted
update are done by the CUDA kernels (`MPPI.propagate`, `MPPI.shift_policy_means`).
"""
import torch


class TensorPolicyMPPI:
    def __init__(self, n_traj, n_dof, tensor_params):
        self.n_dof = n_dof
        self.n_traj = n_traj
        self.params = tensor_params
        self.n_kernels = 0
        self.N_KERNEL_MAX = 50
        self.sigma_c_nominal = 0.2
        K = self.N_KERNEL_MAX
        # policy means
        self.mu_c = torch.zeros((K, n_dof), **tensor_params)
        self.sigma_c = torch.zeros((K,), **tensor_params)
        self.alpha_c = torch.zeros((K, n_dof), **tensor_params)
        # sampling widths
        self.mu_s = torch.tensor(0, **tensor_params)
        self.sigma_s = torch.tensor(0.0, **tensor_params)
        self.alpha_s = torch.tensor(0.0, **tensor_params)
        # per-sample draws
        self.mu_tmp = torch.zeros((n_traj, K, n_dof), **tensor_params)
        self.sigma_tmp = torch.zeros((n_traj, K), **tensor_params)
        self.alpha_tmp = torch.zeros((n_traj, K, n_dof), **tensor_params)
        self.q_min = torch.tensor([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973])
        self.q_max = torch.tensor([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973])
        self.rest = self.q_min + (self.q_max - self.q_min) * 0.5
        self.kernel_gammas = torch.zeros(K, **tensor_params)
        self.kernel_obstacle_bases = torch.zeros((K, n_dof, n_dof), **tensor_params)
        self.p = 2
        self._backend = None          # set by the MPPI object that owns this policy

    def reset_policy(self):
        self.n_kernels = 0
        for t in (self.mu_c, self.sigma_c, self.alpha_c, self.kernel_gammas, self.kernel_obstacle_bases):
            t.zero_()

    def sample_policy(self):
        """Gaussian draws around the means for the live kernels; sample 0 keeps the mean weights.
        RNG call order (mu, sigma, alpha) matches the reference so equal seeds give equal draws."""
        nk = self.n_kernels
        for tmp, std, mean in ((self.mu_tmp, self.mu_s, self.mu_c), (self.sigma_tmp, self.sigma_s, self.sigma_c),
                               (self.alpha_tmp, self.alpha_s, self.alpha_c)):
            tmp.zero_()
            view = tmp[:, :nk]
            view.normal_(mean=0, std=float(std))
            view += mean[:nk]
        self.alpha_tmp[0, :nk] = self.alpha_c[:nk]

    def update_policy(self, w, upd_rate, update_mask=None):
        """Weighted EMA of the means on host tensors (the MPPI object normally does this on the GPU via
        shift_policy_means; this method keeps the reference's standalone call working)."""
        nk = self.n_kernels
        if nk == 0:
            return
        w = w.to(self.mu_tmp.device)
        rate = upd_rate * torch.ones(nk, **self.params)
        if update_mask is not None:
            rate[~update_mask.to(rate.device)] = 0.0
        for cur, tmp in ((self.mu_c, self.mu_tmp), (self.sigma_c, self.sigma_tmp), (self.alpha_c, self.alpha_tmp)):
            wsum = torch.tensordot(w, tmp[:, :nk], dims=([0], [0]))
            r = rate.reshape((nk,) + (1,) * (cur.dim() - 1))
            cur[:nk] = (1 - r) * cur[:nk] + r * wsum

    def update_with_data(self, data):
        if data is None:
            return
        nk = self.n_kernels = data['n_kernels']
        self.mu_c[:nk] = data['mu_c']
        self.alpha_c[:nk] = data['alpha_c']
        self.sigma_c[:nk] = data['sigma_c']
        self.kernel_obstacle_bases[:nk] = data['norm_basis']
        for t in (self.mu_c, self.alpha_c, self.sigma_c, self.kernel_obstacle_bases):
            t[nk:] = 0

    def add_kernel(self, q, kernel_gamma, kernel_obstacle_basis):
        nk = self.n_kernels
        if nk >= self.N_KERNEL_MAX:
            print('Not adding new kernel at: maximum number of kernels reached', q)
            return
        self.mu_c[nk, :] = q
        self.sigma_c[nk] = self.sigma_c_nominal
        if nk > 0:
            nearest = torch.argmin(torch.norm(self.mu_c[:nk, :] - q, 2, 1))
            self.alpha_c[nk, :] = self.alpha_c[nearest, :]
        else:
            self.alpha_c[nk, :] = 0
        self.kernel_gammas[nk] = kernel_gamma
        self.kernel_obstacle_bases[nk] = kernel_obstacle_basis
        self.n_kernels = nk + 1

    def check_traj_for_kernels(self, all_traj, closests_dist_all, dotproducts_all, thr_dist, thr_kernel, thr_dot):
        """States that are close to an obstacle, moving into it, and not yet covered by a kernel
        (policy.py:153-175), in trajectory-major order.  Runs as one CUDA pass over the (N, H) state-steps of
        the MPPI object that owns this policy (dsmppi_kernel_candidates)."""
        if self._backend is None:
            raise RuntimeError("check_traj_for_kernels needs the CUDA backend of the MPPI object that owns this "
                               "policy (there is no CPU path)")
        return self._backend.kernel_candidates(self, all_traj, closests_dist_all, dotproducts_all, thr_dist,
                                               thr_kernel, thr_dot)


def eval_policy(rbf_val, alphas):
    return torch.sum(alphas * rbf_val, 1)


def eval_rbf(q, mu, sigma, p=2):
    """exp(-sigma * ||q - mu||_p^2) with a trailing singleton axis: (n, nk, 1)."""
    dist2 = torch.norm(q[:, None, :] - mu, p=p, dim=2, keepdim=True) ** 2
    return torch.exp(-sigma.unsqueeze(2) * dist2)


def eval_rbf_simple(q, mu, sigma, p=2):
    dist2 = torch.norm(q[:, None, :] - mu, p, -1) ** 2
    return torch.exp(-sigma * dist2)
