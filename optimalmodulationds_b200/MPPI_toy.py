"""Drop-in `MPPI` object of the reference's 2-D point "toy" variant (ds_mppi/functions/MPPI_toy.py).

Same CUDA kernels as optimalmodulationds_b200.MPPI with the other constant set the reference compiles into
MPPI_toy.py -- l_vel = sigmoid(dot; 0, 1, -0.2, 0, 100) (:114), distance sigmoids (0, 0.5, k = 30) and
l_tau_max = 3 (:124-133), goal-activation threshold 0.3 (:176), activations folded into kernel_val_all
(:178-179), repulsion 0.05 (:199) -- the matrix nominal DS v = (q - qf) @ A (:89), the three-term cost of
cost_toy.py:13-19 and the un-masked policy update of MPPI_toy.py:314-324.  API differences kept exactly:
the constructor takes `A` where MPPI.py takes `DS_ARRAY` (:22-23), `propagate()` returns a 4-tuple without
kernel_activations (:207), `shift_policy_means()` returns 0 (:324), `dst_thr` defaults to 0.1 (:56).
The toy network takes [q (2), obstacle x, y] (in_channels = DOF + 2, standaloneToy2d.py:30) and obstacles
are (M, 3) = [x, y, r] (:57-64).
"""
import torch

from . import _capi
from .MPPI import *  # noqa: F401,F403  (the same star-import namespace as MPPI.py:1-8)
from .MPPI import MPPI as _BaseMPPI
from .cost_toy import Cost


class MatrixDS:
    """Nominal dynamics v = (x - q_goal) @ A of MPPI_toy.py:89 as a DS object (q_goal, A, get_velocity)."""

    def __init__(self, q_goal, A):
        self.q_goal = torch.as_tensor(q_goal)
        self.A = torch.as_tensor(A)

    def get_velocity(self, x):
        return (x - self.q_goal.to(x.device)) @ self.A.to(x.device)


class MPPI(_BaseMPPI):
    _TOY = True
    _COST_TERMS = 0
    _UPDATE_VARIANT = 1
    _DEFAULT_DST_THR = 0.1
    _COST_CLASS = Cost

    def __init__(self, q0: torch.Tensor, qf: torch.Tensor, dh_params: torch.Tensor, obs: torch.Tensor, dt: float,
                 dt_H: int, N_traj: int, A: torch.Tensor, dh_a, nn_model, n_closest_obs):
        super().__init__(q0, qf, dh_params, obs, dt, dt_H, N_traj, [MatrixDS(qf, A)], dh_a, nn_model, n_closest_obs)
        self.qf = qf                       # MPPI_toy.py:28 keeps the caller's tensor (no squeeze)

    @property
    def A(self):
        return self.DS.A

    @A.setter
    def A(self, value):
        self.DS.A = value

    def propagate(self):
        all_traj, closest, kval, dots, _ = super().propagate()
        return all_traj, closest, kval, dots

    def shift_policy_means(self):
        super().shift_policy_means()
        return 0
