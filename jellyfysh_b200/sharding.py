"""Multi-GPU layout: independent Markov chains shard over the ranks of one node.

One process per GPU (torch.distributed: NCCL on the GPUs, gloo in the CPU tests). Rank r owns the chains
[r * chains_per_rank, (r + 1) * chains_per_rank) -- start configurations and random streams are numbered by the global
chain index, so a run is reproducible for any number of ranks. There is NO data-path collective: the only exchanges
are sum-reductions of event counters / sample histograms and the max-reduction of timings."""
import numpy as np

COUNTER_KEYS = ("events", "pair_events", "veto_events", "veto_accepted", "boundary_events", "end_of_chain_events",
                "candidates", "bound_violations", "capacity_errors", "bond_events", "factor_pair_events", "pair_targets")


def chain_shard(rank, world_size, chains_per_rank):
    """(first global chain index, number of chains) of a rank; weak scaling: every rank owns chains_per_rank chains."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    return rank * chains_per_rank, chains_per_rank


def split_chains(total_chains, world_size):
    """Strong-scaling alternative: total_chains split as evenly as possible; returns [(first, count)] per rank."""
    base, extra = divmod(total_chains, world_size)
    shards, first = [], 0
    for rank in range(world_size):
        count = base + (1 if rank < extra else 0)
        shards.append((first, count))
        first += count
    return shards


def _all_reduce(array, op_name, device):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return array
    tensor = torch.as_tensor(array, device=device)
    dist.all_reduce(tensor, op=getattr(dist.ReduceOp, op_name))
    return tensor.cpu().numpy()


def reduce_counters(stats, device="cpu"):
    """Sum the EcmcStats dictionaries of all ranks (int64 all-reduce)."""
    values = np.array([int(stats.get(key, 0)) for key in COUNTER_KEYS], dtype=np.int64)
    total = _all_reduce(values, "SUM", device)
    return {key: int(value) for key, value in zip(COUNTER_KEYS, total)}


def reduce_histogram(histogram, device="cpu"):
    """Sum a sample histogram (e.g. pair separations) over all ranks: the estimator reduction of north_star."""
    return _all_reduce(np.ascontiguousarray(histogram, dtype=np.int64), "SUM", device)


def reduce_max(values, device="cpu"):
    """Max over ranks of timings (device times are reported as the slowest rank's)."""
    return _all_reduce(np.ascontiguousarray(values, dtype=np.float64), "MAX", device)
