"""Multi-GPU driver logic of the batched mode: stores are independent units, so the path shards with
no data-path collective. The id space [0, world * n) is cut into one uniform pseudo-random sample of n ids per rank
(`shard_ids`: a bijective bit mixer applied to the rank's contiguous range), because a contiguous range fixes the top
decision variables of a rank and sub-cubes of an EPS decomposition differ in difficulty by tens of percent - measured
as 86 % "scaling efficiency" at 8 GPUs with every rank's step taking the same 15 ms; the propagator table is replicated;
the only exchange is ONE all-reduce (SUM) per step of the 3 + world int64 payload the last block of the batch kernel
fills (include/lpc.h: lpc_eps_set_rank / lpc_batch_set_rank): the three counters {n_solution, n_bot, n_unknown}, then
every rank's best bound in its own slot (zeros elsewhere), so that SUM delivers all bounds and MIN is taken on the host
(`fold_payload`).

Two ways to cut the work: `weak` - every rank 2**16 subproblems of an id space that grows with the number
of ranks; `strong` - the 2**16 subproblems of BASELINE.json's configs[3] dealt to the ranks, 2**16 / world each.

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


def strong_shard_ids(rank, world, total=65536):
    """Strong scaling: rank r's share [r * total / world, (r + 1) * total / world) of ONE batch of `total` subproblems
    (SURVEY.md 8e), taken after the same bit mixing so that every rank gets a uniform sample of the decomposition."""
    import numpy as np
    assert total % world == 0
    per = total // world
    ids = np.arange(rank * per, (rank + 1) * per, dtype=np.int64)
    if world == 1:
        return ids
    bits = (total - 1).bit_length()
    assert total == 1 << bits, "uniform sharding needs a power-of-two number of subproblems"
    return mix_ids(ids, bits)


def allreduce_payload(payload, dist=None):
    """The one collective of a step: in-place SUM all-reduce of the int64 payload [3 + world]."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return payload
    dist.all_reduce(payload, op=dist.ReduceOp.SUM)
    return payload


def fold_payload(payload):
    """Host side of the collective: [n_solution, n_bot, n_unknown, best_bound] from the all-reduced payload."""
    p = [int(v) for v in payload]
    return p[:3] + [min(p[3:])]


def allreduce_record(red, dist=None):
    """In-place all-reduce of the reduction record (an int64 tensor of 4: three counters, then the bound)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return red
    dist.all_reduce(red[:3], op=dist.ReduceOp.SUM)
    dist.all_reduce(red[3:], op=dist.ReduceOp.MIN)
    return red
