"""Nominal linear DS with unit speed outside a small ball around the goal.

Host-side mirror of the reference's ds_mppi/functions/LinDS.py:6-21 (same constructor, attributes and
`get_velocity`).  Inside the rollout the same law is evaluated by the CUDA step kernel from `q_goal` and
`lin_thr`; `get_velocity` is kept for callers that query the DS directly.
"""
import torch


class LinDS:
    def __init__(self, q_goal):
        self.q_goal = torch.as_tensor(q_goal).clone()
        self.lin_thr = 0.015
        self.dof = self.q_goal.shape[0]

    def get_velocity(self, x):
        goal = self.q_goal.to(x.device)
        vel = -(x - goal)
        speed = vel.norm(p=2, dim=-1, keepdim=True)
        far = speed.squeeze(-1) > self.lin_thr
        return torch.where(far.unsqueeze(-1), vel / speed, vel)
