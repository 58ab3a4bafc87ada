"""The learned-distance MLP container (mirror of mlp_learn/sdf/network_macros_mod.py:106-146).

Only what the MPPI path needs: a module whose parameters carry the reference's state_dict keys
(`layers.<block>.<i>.0.{weight,bias}`) so shipped checkpoints load, and a torch `forward` for callers that
evaluate the net directly (plots).  The rollout itself reads the weights once and runs them in CUDA.
"""
import torch
from torch import nn


def _block(widths, act_fn, last_linear):
    mods = []
    n = len(widths) - 1
    for i in range(n):
        if last_linear and i == n - 1:
            mods.append(nn.Sequential(nn.Linear(widths[i], widths[i + 1])))
        else:
            mods.append(nn.Sequential(nn.Linear(widths[i], widths[i + 1]), act_fn()))
    return nn.Sequential(*mods)


class MLPRegression(nn.Module):
    def __init__(self, input_dims=10, output_dims=1, mlp_layers=(256, 256, 256, 256), skips=(), act_fn=nn.ReLU,
                 nerf=True):
        super().__init__()
        if len(skips) > 0:
            raise NotImplementedError("skip connections are not used by any shipped checkpoint")
        self.nerf = nerf
        in_dim = 3 * input_dims if nerf else input_dims
        self.layers = nn.ModuleList([_block([in_dim, *mlp_layers, output_dims], act_fn, last_linear=True)])

    def forward(self, x):
        h = torch.cat((x, torch.sin(x), torch.cos(x)), dim=-1) if self.nerf else x
        return self.layers[0](h)
