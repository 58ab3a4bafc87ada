"""optimalmodulationds_b200 -- B200-native ds_mppi MPPI rollout behind the reference's Python MPPI API.

The compute lives in libdsmppi_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/dsmppi_b200.h); this package is the host-side mirror of the reference interface
(ds_mppi/functions/{MPPI,policy,cost,LinDS,fk_num}.py, mlp_learn/sdf/robot_sdf.py).
There is no CPU fallback: constructing `MPPI` without the library or without a B200 raises.
"""
from .MPPI import MPPI, generalized_sigmoid  # noqa: F401
from .policy import TensorPolicyMPPI, eval_rbf, eval_rbf_simple  # noqa: F401
from .cost import Cost  # noqa: F401
from .LinDS import LinDS  # noqa: F401
from .SEDS import SEDS  # noqa: F401

__all__ = ["MPPI", "TensorPolicyMPPI", "Cost", "LinDS", "SEDS", "eval_rbf", "eval_rbf_simple", "generalized_sigmoid"]
