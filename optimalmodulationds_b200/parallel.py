"""Sample-sharded MPPI across the GPUs of one node.

Samples never interact inside propagate/get_cost, so rank r simply owns N/G of them; obstacles, network
weights and policy means are replicated.  The policy update needs exactly two tiny exchanges per iteration,
both plain NCCL all-reduces over NVLink/NVSwitch on the packed vectors produced by the CUDA reduction
kernels:
  1. cost statistics  SUM of [sum cost, N]             -> global beta = mean(cost)/50
  2. packed weighted sums (<= 3.4 KB), SUM             -> every rank applies the identical EMA
Neither needs the host: the global sample count is known when sharding is enabled, so an iteration enqueues two
collectives and never synchronises.
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
    """stats = [sum cost, N, min cost, argmin (local)] per rank (cost_stats_kernel): sums the first two entries over
    the ranks, in place, with ONE collective.  Entries 2..3 stay per-shard (see allreduce_best)."""
    dist.all_reduce(stats[0:2], op=dist.ReduceOp.SUM, group=group)
    return stats


def allreduce_best(stats, rank_offset, group=None):
    """Global (min cost, global sample index) from per-rank stats -- only get_qdot('best') on a sharded job needs it.
    Returns a 2-vector on stats' device; ties go to the lowest global index like torch.argmin."""
    world = dist.get_world_size(group)
    mine = torch.stack((stats[2], stats[3] + float(rank_offset)))
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine, group=group)
    allv = torch.stack(allv)
    order = torch.argsort(allv[:, 1])
    allv = allv[order]
    return allv[torch.argmin(allv[:, 0])]


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
