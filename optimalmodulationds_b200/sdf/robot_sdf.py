"""RobotSdfCollisionNet: loader/holder of the learned distance network.

Mirror of mlp_learn/sdf/robot_sdf.py:13-51 (constructor and load_weights signatures, `model`,
`in_channels`, `out_channels`, `model_jit`, `update_aot_lambda`, `allocate_gradients`): the MPPI object only
reads `.model`'s weights and the channel counts.  The direct-evaluation helpers run the torch module on
whatever device it is on; they are not used by the rollout.
"""
import torch
from torch import nn

from .network_macros_mod import MLPRegression


class RobotSdfCollisionNet:
    def __init__(self, in_channels, out_channels, skips, layers):
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.model = MLPRegression(in_channels, out_channels, list(layers), list(skips), act_fn=nn.ReLU, nerf=True)
        self.model_jit = self.model
        self.order = list(range(out_channels))
        self.norm_dict = None
        self.tensor_args = {'device': 'cpu', 'dtype': torch.float32}

    def set_link_order(self, order):
        self.order = order

    def load_weights(self, f_name, tensor_args):
        try:
            chk = torch.load(f_name, map_location=torch.device('cpu'), weights_only=False)
            self.model.load_state_dict(chk["model_state_dict"])
            self.norm_dict = chk.get("norm")
            print('Weights loaded!')
        except Exception as exc:  # noqa: BLE001 - same forgiving behaviour as the reference
            print('WARNING: Weights not loaded')
            print(exc)
        self.model = self.model.to(**tensor_args)
        self.model_jit = self.model
        self.tensor_args = tensor_args
        self.model.eval()

    def load_arrays(self, weights, biases):
        """Load from plain (out, in) arrays (tests/golden/weights/*.npz)."""
        lin = [m for m in self.model.modules() if isinstance(m, nn.Linear)]
        assert len(lin) == len(weights)
        with torch.no_grad():
            for m, W, b in zip(lin, weights, biases):
                m.weight.copy_(torch.as_tensor(W))
                m.bias.copy_(torch.as_tensor(b))
        self.model.eval()

    def update_aot_lambda(self):
        """The reference compiles a functorch VJP here (robot_sdf.py:164-166); the CUDA path has its own
        analytic VJP, so this is a no-op kept for script compatibility."""
        self.aot_lambda = self.functorch_vjp
        return 0

    def allocate_gradients(self, N, tensor_args):
        self.maxInputSize = N

    def compute_signed_distance(self, q):
        with torch.no_grad():
            return self.model(q)[:, self.order].detach()

    def functorch_vjp(self, points):
        dists, vjp_fn = torch.func.vjp(self.model.forward, points)
        idx = torch.argmin(dists, dim=1)
        seed = torch.zeros_like(dists)
        seed[torch.arange(points.shape[0]), idx] = 1
        return dists.detach(), vjp_fn(seed)[0].detach(), idx.detach()

    def dist_grad_closest_aot(self, q):
        return self.functorch_vjp(q)
