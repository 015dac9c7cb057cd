"""Worker of tests/test_gpu_multi.py, one process per GPU (torchrun): the ranks solve disjoint shards of one EPS batch with
the record exchange fused into the kernel (lpc_eps_peer_*), and every rank must end up with the record of the whole batch -
which rank 0 also computes alone, without peers, on all the ids."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import lala_pc_b200 as L  # noqa: E402
from lala_pc_b200 import sharding, workloads as W  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("gloo")   # plumbing only: the handles travel through it, the records do not
L.device_init(torch.cuda.current_device())

net = W.pir_network(2_000, 10_000, W.SEED_BASE + 99, window=256, value_range=1024)
table = L.Table(net.records, net.nvars)
s = L.Store(values=net.store)
L.fixpoint(table, s)
root = s.read()
dec, obj = W.eps_decisions(net.records, root, n=12)
total = 1 << len(dec)
ids = sharding.strong_shard_ids(rank, world, total)



class DevView:   # a device int64 vector of the library, viewed as a torch tensor (no copy)
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}


eps = L.Eps(table, len(ids))
eps.set_rank(rank, world)
handles = [None] * world
dist.all_gather_object(handles, eps.peer_export())
eps.peer_connect(rank, world, handles)
payload = torch.as_tensor(DevView(*eps.payload), device="cuda")
dist.barrier()
eps.upload(root, dec, ids=ids)
records = []
for step in range(3):   # three calls: the two slot sets of the inbox alternate
    r = eps.run(objective_var=obj)
    rec = sharding.fold_payload(payload.cpu().tolist()[:3 + world])
    records.append((rec, r))
own = [int(r.n_solution), int(r.n_bot), int(r.n_unknown), int(r.best_bound)]
gathered = [None] * world
dist.all_gather_object(gathered, own)
want = [sum(g[0] for g in gathered), sum(g[1] for g in gathered), sum(g[2] for g in gathered), min(g[3] for g in gathered)]
for rec, _ in records:
    assert [int(x) for x in rec] == want, (rank, rec, want)
if rank == 0:   # the same batch on one GPU, no peers
    alone = L.Eps(table, total)
    alone.upload(root, dec, ids=np.arange(total, dtype=np.int64))
    ra = alone.run(objective_var=obj)
    assert want == [int(ra.n_solution), int(ra.n_bot), int(ra.n_unknown), int(ra.best_bound)], (want, ra.as_dict())
    assert want[0] + want[1] + want[2] == total
    alone.close()
eps.peer_disconnect()
dist.barrier()
eps.close()
dist.destroy_process_group()
print("rank %d ok: %s" % (rank, want))
