"""Multi-GPU driver logic of the batched mode: stores are independent units, so the path shards with
no data-path collective. The id space [0, world * n) is cut into one uniform pseudo-random sample of n ids per rank
(`shard_ids`: a bijective bit mixer applied to the rank's contiguous range), because a contiguous range fixes the top
decision variables of a rank and sub-cubes of an EPS decomposition differ in difficulty by tens of percent - measured
as 86 % "scaling efficiency" at 8 GPUs with every rank's step taking the same 15 ms; the propagator table is replicated;
the only exchange is one all-reduce of the 4 x int64 reduction record {n_solution, n_bot, n_unknown, best_bound}
(SUM over the three counters, MIN over the bound) after the last kernel of a step.

Pure torch.distributed plumbing (NCCL on the GPUs, gloo in the CPU tests); no compute of its own.
"""


def decision_bits(world, base_bits=16):
    """Weak scaling: 2**base_bits subproblems per rank, so the id space needs log2(world) more decision variables."""
    return base_bits + max(0, (world - 1).bit_length())


def shard_first_id(rank, stores_per_rank):
    return rank * stores_per_rank


def mix_ids(x, bits):
    """A bijection of [0, 2**bits) that spreads every input bit over the whole word (xorshift / odd multiply rounds on
    `bits`-bit words), vectorised over a uint64 array."""
    import numpy as np
    mask = np.uint64((1 << bits) - 1)
    x = np.asarray(x, dtype=np.uint64) & mask
    if bits < 2:
        return x.astype(np.int64)
    sh = np.uint64(max(1, bits // 2))
    for mul in (0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB):
        x = x ^ (x >> sh)
        x = (x * np.uint64(mul | 1)) & mask
    x = x ^ (x >> sh)
    return x.astype(np.int64)


def shard_ids(rank, world, stores_per_rank):
    """The subproblem ids of `rank`: consecutive for a single rank (the id order of BASELINE.json's configs[3]), a uniform
    sample of the `world * stores_per_rank` ids otherwise. Requires a power-of-two id space when world > 1."""
    import numpy as np
    lo = rank * stores_per_rank
    ids = np.arange(lo, lo + stores_per_rank, dtype=np.int64)
    if world == 1:
        return ids
    total = world * stores_per_rank
    bits = (total - 1).bit_length()
    assert total == 1 << bits, "uniform sharding needs a power-of-two number of subproblems"
    return mix_ids(ids, bits)


def allreduce_record(red, dist=None):
    """In-place all-reduce of the reduction record (an int64 tensor of 4: three counters, then the bound)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return red
    dist.all_reduce(red[:3], op=dist.ReduceOp.SUM)
    dist.all_reduce(red[3:], op=dist.ReduceOp.MIN)
    return red
