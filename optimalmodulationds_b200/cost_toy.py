"""Trajectory cost of the 2-D point "toy" variant.

Host-side mirror of the reference's ds_mppi/functions/cost_toy.py:4-19: the same class as cost.py but
`evaluate_costs` keeps only goal + collision + stagnation (no joint-limit flag, no terminal FK).  It runs
the same CUDA cost kernel (csrc/rollout_kernels.cu: cost_kernel) with the optional terms switched off
(dsmppi_cost_args.terms = 0).
"""
from .cost import Cost as _Cost


class Cost(_Cost):
    pass
