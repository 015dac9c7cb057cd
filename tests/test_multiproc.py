"""The N > 1 path on CPU: world_size 2 over gloo. Each rank takes its shard of the EPS subproblem-id range, runs the
fixpoints of its shard (with the oracle here: the GPU kernel is exercised by the -m gpu tests), fills the 4 x int64
reduction record exactly like the batch kernel does, and the sharding helper all-reduces it. The result must equal
the single-process reduction over the union of the shards."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

N_PER_RANK = 128


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _shard_record(rank, world, n_per_rank):
    import lala_pc_b200  # noqa: F401  (package import path)
    from lala_pc_b200 import sharding, workloads as W
    from oracle import oracle as O
    net = W.config4_base()
    root, _ = O.pir_fixpoint(net.store, net.records)
    dec, obj = W.eps_decisions(net.records, root, n=24)
    dec = dec[:sharding.decision_bits(world, base_bits=7)]
    ids = sharding.shard_ids(rank, world, n_per_rank)
    stores = W.eps_stores(root, dec, 0, n_per_rank, ids=ids)
    out, flags, _, _, _ = O.pir_batch_fixpoint(stores, net.records, threads=2)
    bot = (flags & 1) != 0
    sol = ((flags & 2) != 0) & ~bot
    best = int(out[~bot][:, obj, 0].min()) if (~bot).any() else 2**31 - 1
    return [int(sol.sum()), int(bot.sum()), int((~bot & ~sol).sum()), best]


def _worker(rank, world, port, n_per_rank, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lala_pc_b200 import sharding
    rec = _shard_record(rank, world, n_per_rank)
    red = torch.tensor(rec, dtype=torch.int64)
    sharding.allreduce_record(red, dist)
    # the packed form the kernels fill (include/lpc.h: lpc_eps_set_rank): ONE SUM all-reduce, MIN folded on the host
    payload = torch.zeros(3 + world, dtype=torch.int64)
    payload[:3] = torch.tensor(rec[:3])
    payload[3 + rank] = rec[3]
    sharding.allreduce_payload(payload, dist)
    assert sharding.fold_payload(payload.tolist()) == red.tolist()
    q.put((rank, red.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_batch_reduction_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N_PER_RANK, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference over the union of both shards
    parts = [_shard_record(r, world, N_PER_RANK) for r in range(world)]
    want = [sum(p[i] for p in parts) for i in range(3)] + [min(p[3] for p in parts)]
    assert got[0] == want and got[1] == want
    assert want[0] + want[1] + want[2] == world * N_PER_RANK
    assert parts[0] != parts[1], "shards must be distinct subproblem ranges"


def test_sharding_helpers():
    from lala_pc_b200 import sharding
    assert [sharding.decision_bits(w) for w in (1, 2, 4, 8)] == [16, 17, 18, 19]
    assert sharding.shard_first_id(3, 65536) == 3 * 65536
    assert sharding.shard_ids(0, 1, 1000).tolist() == list(range(1000))       # a single rank keeps the id order
    for world in (2, 4, 8):
        parts = [sharding.shard_ids(r, world, 4096) for r in range(world)]
        allids = np.concatenate(parts)
        assert sorted(allids.tolist()) == list(range(world * 4096))           # a partition of the id space
        for p in parts:                                                        # ... into uniform samples: no fixed bit
            for j in range((world * 4096 - 1).bit_length()):
                assert 0.4 < float(((p >> j) & 1).mean()) < 0.6
    for world in (1, 2, 8):                                                   # strong scaling: ONE batch of 65,536 dealt out
        parts = [sharding.strong_shard_ids(r, world) for r in range(world)]
        assert all(len(p) == 65536 // world for p in parts)
        assert sorted(np.concatenate(parts).tolist()) == list(range(65536))
    assert sharding.fold_payload([5, 6, 7, 40, 30, 50]) == [5, 6, 7, 30]
    red = torch.tensor([1, 2, 3, 4], dtype=torch.int64)
    assert sharding.allreduce_record(red, None).tolist() == [1, 2, 3, 4]
