"""Trajectory cost of one MPPI iteration.

Host-side mirror of the reference's ds_mppi/functions/cost.py:4-46 (class Cost: goal, collision count,
joint-limit flag, stagnation, terminal-FK terms; `rest_cost` is computed but unused there and is skipped
here).  `evaluate_costs` runs the CUDA cost kernel (csrc/rollout_kernels.cu: cost_kernel) through the
backend of the MPPI object that owns this Cost; `q_min`/`q_max`/`qf` are plain attributes read at call
time, as the reference's scripts assign them after construction (standalonePlanar2d.py:128-129).
"""
import torch

from .fk_num import numeric_fk_model


class Cost:
    def __init__(self, q_f, dh_params, backend=None):
        self.qf = q_f
        self.COLL_WEIGHT = 500
        self.dh_params = dh_params
        self.goal_fk = numeric_fk_model(torch.as_tensor(q_f).detach().cpu().float(), dh_params.detach().cpu(), 2)[0]
        self.q_min = torch.tensor([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973])
        self.q_max = torch.tensor([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973])
        self.rest = self.q_min + (self.q_max - self.q_min) * 0.5
        self._backend = backend

    def evaluate_costs(self, all_traj, closest_dist_all):
        if self._backend is None:
            raise RuntimeError("Cost.evaluate_costs needs the CUDA backend of an MPPI object "
                               "(there is no CPU path); use MPPI.get_cost()")
        return self._backend.evaluate_costs(self, all_traj, closest_dist_all)
