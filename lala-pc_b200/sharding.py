"""Multi-GPU driver logic of the batched mode: stores are independent units, so the path shards with
no data-path collective. Rank r owns the subproblem-id range [r * n, (r + 1) * n); the propagator table is replicated;
the only exchange is one all-reduce of the 4 x int64 reduction record {n_solution, n_bot, n_unknown, best_bound}
(SUM over the three counters, MIN over the bound) after the last kernel of a step.

Pure torch.distributed plumbing (NCCL on the GPUs, gloo in the CPU tests); no compute of its own.
"""


def decision_bits(world, base_bits=16):
    """Weak scaling: 2**base_bits subproblems per rank, so the id space needs log2(world) more decision variables."""
    return base_bits + max(0, (world - 1).bit_length())


def shard_first_id(rank, stores_per_rank):
    return rank * stores_per_rank


def allreduce_record(red, dist=None):
    """In-place all-reduce of the reduction record (an int64 tensor of 4: three counters, then the bound)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return red
    dist.all_reduce(red[:3], op=dist.ReduceOp.SUM)
    dist.all_reduce(red[3:], op=dist.ReduceOp.MIN)
    return red
