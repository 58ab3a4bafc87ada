"""Sample-sharded MPPI across the GPUs of one node.

Samples never interact inside propagate/get_cost, so rank r simply owns N/G of them; obstacles, network
weights and policy means are replicated.  The policy update needs exactly two tiny exchanges per iteration,
both plain NCCL all-reduces over NVLink/NVSwitch on the packed vectors produced by the CUDA reduction
kernels:
  1. cost statistics  [sum cost, min cost, argmin, N]  -> global beta = mean(cost)/50
  2. packed weighted sums (<= 3.2 KB)                  -> every rank applies the identical EMA
The functions below operate on torch tensors on any device so the same logic is exercised with the gloo
backend on CPU in tests/test_parallel_gloo.py.
"""
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world):
    """Contiguous, balanced [lo, hi) of global sample indices owned by `rank`; sample 0 lives on rank 0."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_cost_stats(stats, group=None):
    """stats = [sum cost, min cost, argmin (local), N] per rank -> global sum / min / N in place.
    Returns the global sample count."""
    sums = torch.stack((stats[0], stats[3]))
    mn = stats[1:2].clone()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
    stats[0], stats[3], stats[1] = sums[0], sums[1], mn[0]
    return int(round(float(sums[1])))


def allreduce_packed(packed, group=None):
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed


def finalize_from_packed(packed, nk, d, n_global, ker_thr, upd_rate, mu_c, sigma_c, alpha_c):
    """Host restatement of the finalize kernel's arithmetic on the all-reduced packed vector; used by the
    gloo test to check that sharding + summation reproduces the unsharded update."""
    o_mu, o_sg = 1, 1 + nk * d
    o_al = o_sg + nk
    o_mx = o_al + nk * d
    o_b0 = o_mx + nk
    wsum = packed[0]
    on = (packed[o_mx:o_mx + nk] / n_global > ker_thr) & (packed[o_b0:o_b0 + nk] > ker_thr)
    r = torch.where(on, torch.full((nk,), float(upd_rate)), torch.zeros(nk))
    mu_c, sigma_c, alpha_c = mu_c.clone(), sigma_c.clone(), alpha_c.clone()
    mu_c[:nk] = (1 - r[:, None]) * mu_c[:nk] + r[:, None] * (packed[o_mu:o_sg].reshape(nk, d) / wsum)
    sigma_c[:nk] = (1 - r) * sigma_c[:nk] + r * (packed[o_sg:o_al] / wsum)
    alpha_c[:nk] = (1 - r[:, None]) * alpha_c[:nk] + r[:, None] * (packed[o_al:o_mx].reshape(nk, d) / wsum)
    return mu_c, sigma_c, alpha_c, int(on.sum())
