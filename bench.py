#!/usr/bin/env python
"""bench.py — propagator deductions/s of the fixpoint hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload auto|c1|c2|c4]

One "step" = one pass of the hot path over one batch of synthetic input:
  N == 1  -> BASELINE.json configs[1]: ONE fixpoint of the 1M-variable / 5M-propagator PIR network (dense-sweep mode,
             the mode whose work unit equals the reference's: every sweep evaluates every propagator). The batched
             mode (configs[3]) is measured too and reported under "batched".
  N  > 1  -> BASELINE.json configs[3]: batched EPS, 65,536 subproblem stores of the 2k-variable / 10k-propagator model
             PER GPU (weak scaling, disjoint subproblem-id ranges), one thread block per store, no inter-GPU traffic
             inside the fixpoints and one NCCL all-reduce of the 32-byte reduction record per step.
`value` is measured with inputs resident in HBM (CUDA events on the launching stream, L2 flushed between steps);
`e2e` goes through the C-ABI's host-buffer entry points (pinned host memory, H2D + fixpoint + D2H inside the timed
region). `--impl reference` times the CPU restatement of the reference's Gauss-Seidel path (oracle/) on the host.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "propagator deductions/sec"
UNIT = "deductions/s"
BYTES_PER_DEDUCTION = 40          # 16 B record + 3 x 8 B bounds (SURVEY.md §8d / BASELINE.md §3)
STORES_PER_GPU = 65536


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v for v in vis.split(",") if v.strip().isdigit()]
        self.index = int(ids[index]) if index < len(ids) else index
        self.rows, self.stop, self.t = [], threading.Event(), None

    def _run(self):
        # In-process NVML (nvidia_ml_py): spawning nvidia-smi every few ms perturbs the measurement itself (a 27 ms
        # batched launch was seen to take 90 ms under a 50 ms nvidia-smi polling loop).
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = [nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown,
                    nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap]
            while not self.stop.is_set():
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.rows.append([str(sm), str(mx)] + ["Active" if rs & b else "Not Active" for b in bits])
                self.stop.wait(0.02)
            return
        except Exception:
            pass
        # fallback: one nvidia-smi query (not a polling loop)
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                  "-i", str(self.index)], capture_output=True, text=True, timeout=10).stdout
            self.rows.append([c.strip() for c in out.strip().split(",")])
        except Exception:
            pass

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=10)

    def summary(self):
        sm = [int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------------
def build_c2(scale):
    from lala_pc_b200 import workloads as W
    return W.config2(scale)


def build_c4():
    """Base model, its root fixpoint (computed with the product itself), decisions, objective."""
    import lala_pc_b200 as L
    from lala_pc_b200 import workloads as W
    net = W.config4_base()
    table = L.Table(net.records, net.nvars)
    s = L.Store(values=net.store)
    r = L.fixpoint(table, s)
    assert not r.is_bot
    root = s.read()
    dec, obj = W.eps_decisions(net.records, root, n=24)
    return net, table, root, dec, obj


def own_single(args, rank):
    """N == 1: one fixpoint of the single-store network per step."""
    import torch
    import lala_pc_b200 as L
    from lala_pc_b200 import workloads as W
    net = W.config1() if args.workload == "c1" else build_c2(args.scale)
    name = ("PIR single fixpoint: %d vars / %d ternary propagators (BASELINE.json configs[%d])"
            % (net.nvars, len(net.records), 0 if args.workload == "c1" else 1))
    table = L.Table(net.records, net.nvars)
    P = len(net.records)
    n_slots = args.steps + args.warmup
    stores = [L.Store(values=net.store) for _ in range(n_slots)]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    mode = L.MODE_SWEEP
    launches0 = L.launch_count()
    results = []
    torch.cuda.synchronize()
    with ClockSampler(rank) as clk:
        for i in range(n_slots):
            flush.zero_()
            if i == args.warmup:
                torch.cuda.synchronize()
                launches0 = L.launch_count()
            L.fixpoint_async(table, stores[i], mode=mode)
            r = L.fixpoint_collect(stores[i])
            if i >= args.warmup:
                results.append(r.as_dict())
        torch.cuda.synchronize()
    launches = L.launch_count() - launches0
    dev_ms = [r["device_ms"] for r in results]
    ded = [r["deductions"] for r in results]
    total_ms = float(np.sum(dev_ms))
    value = float(np.sum(ded)) / (total_ms * 1e-3)
    # latency of the other modes (same input, untimed warm-up then best of 3)
    latency = {"sweep_ms": float(np.mean(dev_ms)), "sweeps": results[0]["sweeps"]}
    for mname, m in (("auto", dict(mode=L.MODE_AUTO)), ("worklist", dict(mode=L.MODE_WORKLIST))):
        best = None
        for _ in range(4):
            s = L.Store(values=net.store)
            flush.zero_()
            r = L.fixpoint(table, s, **m)
            best = r if best is None or r.device_ms < best.device_ms else best
        latency[mname + "_ms"] = float(best.device_ms)
        latency[mname + "_deductions"] = int(best.deductions)
        latency[mname + "_iterations"] = int(best.sweeps)
    # end to end through the host-buffer entry point
    pinned = torch.empty((net.nvars, 2), dtype=torch.int32).pin_memory()
    src = torch.from_numpy(net.store)
    e2e_t, e2e_d = 0.0, 0
    for i in range(args.warmup + args.steps):
        pinned.copy_(src)
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = L.fixpoint_host(table, pinned.data_ptr(), mode=mode)
        t1 = time.perf_counter()
        if i >= args.warmup:
            e2e_t += t1 - t0
            e2e_d += r.deductions
    peak, peak_src = measured_peak()
    per_launch_ms = float(np.mean(dev_ms))
    achieved = BYTES_PER_DEDUCTION * float(np.mean(ded)) / (per_launch_ms * 1e-3) / 1e9
    # the working set (65 MB table + 8 MB store) lives in L2: the on-chip ceiling, measured by the library's own copy kernel
    # on an L2-resident 32 MB buffer pair (the driver provides an HBM figure only)
    l2_gbs = float(L.measure_l2_copy_gbs(32 << 20, 20))
    out = {
        "value": value, "ms_per_step": per_launch_ms, "gpu_launches": int(launches),
        "config": {"workload": name, "mode": "dense sweeps (LPC_MODE_SWEEP)", "vars": net.nvars, "propagators": P,
                   "sweeps_per_fixpoint": results[0]["sweeps"], "timing": "CUDA events around each fixpoint launch",
                   "l2": "flushed between steps (512 MiB write)", "seed": net.meta["seed"]},
        "e2e": {"value": e2e_d / e2e_t, "unit": UNIT, "h2d_bytes_per_step": net.nvars * 8,
                "d2h_bytes_per_step": net.nvars * 8 + 64, "ms_per_step": e2e_t / args.steps * 1e3,
                "what": "lpc_fixpoint_host: pinned host store -> device, fixpoint, store -> host (table resident)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic("k_pir_fixpoint"), "peak_source": peak_src, "kernel": "k_pir_fixpoint",
                     "algorithmic_bytes_per_launch": BYTES_PER_DEDUCTION * float(np.mean(ded)),
                     "l2_copy_peak": l2_gbs, "frac_of_l2_copy_peak": achieved / l2_gbs if l2_gbs > 0 else None,
                     "l2_copy_peak_source": "lpc_measure_l2_copy_gbs: read + write GB/s of a 32 MB buffer pair resident in L2"},
        "latency": latency, "clocks": clk.summary(),
    }
    return out, net, table


def own_batched(args, rank, world, first_id=None, n_stores=STORES_PER_GPU, steps=None, warmup=None, e2e=True):
    """Batched EPS on this rank: n_stores subproblems, one block per store; one all-reduce per step when world > 1."""
    import torch
    import lala_pc_b200 as L
    steps = args.steps if steps is None else steps
    warmup = args.warmup if warmup is None else warmup
    from lala_pc_b200 import sharding
    net, table, root, dec, obj = build_c4()
    nbits = sharding.decision_bits(world)
    dec = dec[:nbits]
    # rank r works on a uniform sample of the world * n_stores subproblem ids (sharding.shard_ids), not on a sub-cube
    ids = sharding.shard_ids(rank, world, n_stores) if first_id is None else first_id + np.arange(n_stores, dtype=np.int64)
    batch = L.Batch(table, n_stores)

    class _Red:   # the 4 x int64 reduction record of the library, viewed as a torch tensor (no copy)
        def __init__(self, ptr):
            self.__cuda_array_interface__ = {"shape": (4,), "typestr": "<i8", "data": (ptr, False), "version": 2}

    red = torch.as_tensor(_Red(batch.reduction_device_ptr), device="cuda")
    dist = None
    if world > 1:
        import torch.distributed as dist
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    res = None
    launches0 = L.launch_count()
    with ClockSampler(rank) as clk:
        for i in range(warmup + steps):
            batch.init_split(root, dec, ids=ids)             # untimed: inputs resident in HBM (1 GiB > L2)
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            if i == warmup:
                launches0 = L.launch_count()
            if i >= warmup:
                ev[i - warmup][0].record()
            batch.fixpoint_async(objective_var=obj, mode=L.MODE_SWEEP)   # dense: the reference's work unit
            sharding.allreduce_record(red, dist)
            if i >= warmup:
                ev[i - warmup][1].record()
            res = batch.collect()
            torch.cuda.synchronize()
    launches = L.launch_count() - launches0
    ms = [a.elapsed_time(b) for a, b in ev]
    total = torch.tensor([float(np.sum(ms)), float(res.deductions), float(res.sweeps_total)], dtype=torch.float64, device="cuda")
    tmax = total.clone()
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)          # time: max over ranks
        dist.all_reduce(total, op=dist.ReduceOp.SUM)         # work: sum over ranks
    step_ms = float(tmax[0]) / steps
    ded_per_step = float(total[1])
    value = ded_per_step / (step_ms * 1e-3)
    reduced = [int(v) for v in red.cpu().tolist()]
    out = {
        "value": value, "ms_per_step": step_ms, "gpu_launches": int(launches),
        "config": {"workload": "batched EPS: %d subproblem stores per GPU of a %d-var / %d-propagator PIR model, one "
                               "block per store (BASELINE.json configs[3])" % (n_stores, net.nvars, len(net.records)),
                   "stores_total": n_stores * world, "decision_bits": nbits, "timing": "CUDA events, max over ranks",
                   "sharding": "consecutive ids" if world == 1 else "each rank a uniform sample of the id space (sharding.shard_ids)",
                   "l2": "inputs larger than L2 (1 GiB of stores per GPU)", "collective":
                   "one NCCL all-reduce pair (SUM, MIN) over the 32-byte reduction record per step" if world > 1 else "none",
                   "seed": net.meta["seed"]},
        "batch_result": {"n_solution": reduced[0], "n_bot": reduced[1], "n_unknown": reduced[2], "best_bound": reduced[3],
                         "sweeps_total": float(total[2]), "max_sweeps_seen": res.max_sweeps_seen},
        "clocks": clk.summary(),
    }
    # time to result of the same batch in the change-driven modes (same fixpoints, fewer deductions)
    lat = {"sweep_ms": float(np.mean(ms)), "sweep_deductions": int(res.deductions)}
    for mname, seeds in (("change_driven", None), ("change_driven_seeded", dec)):
        batch.set_seeds(seeds)
        best = None
        for _ in range(3):
            batch.init_split(root, dec, ids=ids)
            r = batch.fixpoint(objective_var=obj, mode=L.MODE_WORKLIST)
            best = r if best is None or r.device_ms < best.device_ms else best
        lat[mname + "_ms"] = float(best.device_ms)
        lat[mname + "_deductions"] = int(best.deductions)
    batch.set_seeds(None)
    out["latency"] = lat
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_DEDUCTION * float(res.deductions) / (float(np.mean(ms)) * 1e-3) / 1e9
    out["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                       "traffic": ncu_traffic("k_pir_batch"), "peak_source": peak_src, "kernel": "k_pir_batch",
                       "algorithmic_bytes_per_launch": BYTES_PER_DEDUCTION * float(res.deductions),
                       "note": "per GPU; stores and table are shared-memory resident, so the algorithmic rate may "
                               "exceed the HBM copy peak"}
    if e2e:
        pinned = torch.empty((n_stores, net.nvars, 2), dtype=torch.int32).pin_memory()
        batch.init_split(root, dec, ids=ids)
        src = torch.from_numpy(batch.read())
        n_e2e = max(1, min(steps, 3))
        t_e2e, d_e2e = 0.0, 0
        for i in range(1 + n_e2e):
            pinned.copy_(src)
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = batch.fixpoint_host(pinned.data_ptr(), objective_var=obj, mode=L.MODE_SWEEP)
            if dist is not None:
                sharding.allreduce_record(red, dist)
                torch.cuda.synchronize()
            t1 = time.perf_counter()
            if i >= 1:
                t_e2e += t1 - t0
                d_e2e += r.deductions
        tt = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
        dd = torch.tensor([float(d_e2e)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(dd, op=dist.ReduceOp.SUM)
        sb = n_stores * net.nvars * 8
        out["e2e"] = {"value": float(dd[0]) / float(tt[0]), "unit": UNIT, "h2d_bytes_per_step": sb, "d2h_bytes_per_step": sb + 64,
                      "ms_per_step": float(tt[0]) / n_e2e * 1e3, "steps": n_e2e,
                      "what": "lpc_batch_fixpoint_host: pinned host stores -> device, fixpoints, stores -> host (8 chunks on three streams)"}
        # The EPS call sequence as a solver makes it: the subproblems are GENERATED on the device from the root store,
        # the decision list and the subproblem ids (lpc_batch_init_split_ids), so a step moves the root store and the ids
        # in, and the per-store flags + the reduction record out - not 1 GiB of store images each way. Reported next to
        # `e2e` (which keeps the conservative reading: every store image crosses the link twice).
        t_sp, d_sp = 0.0, 0
        for i in range(1 + n_e2e):
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            batch.init_split(root, dec, ids=ids)
            r = batch.fixpoint(objective_var=obj, mode=L.MODE_SWEEP)
            fl = batch.flags()
            if dist is not None:
                sharding.allreduce_record(red, dist)
                torch.cuda.synchronize()
            t1 = time.perf_counter()
            if i >= 1:
                t_sp += t1 - t0
                d_sp += r.deductions
        tt = torch.tensor([t_sp], dtype=torch.float64, device="cuda")
        dd = torch.tensor([float(d_sp)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(dd, op=dist.ReduceOp.SUM)
        out["e2e_split"] = {"value": float(dd[0]) / float(tt[0]), "unit": UNIT,
                            "h2d_bytes_per_step": int(net.nvars * 8 + 4 * len(dec) + 8 * n_stores), "d2h_bytes_per_step": int(len(fl) + 64),
                            "ms_per_step": float(tt[0]) / n_e2e * 1e3, "steps": n_e2e,
                            "what": "lpc_batch_init_split_ids (root store + decisions + ids from the host) -> fixpoints -> "
                                    "lpc_batch_flags + reduction record to the host"}
    batch.close()
    return out


def own_pc(args, cpu=True):
    """PC configs (BASELINE.json configs[2] and [4]): one fixpoint of the flattened n-ary propagators per step, on an
    interval store and (config 5) on an NBitset<64> store. Algorithmic bytes per deduction = 16 B header + per term
    8 B {coef, var} + 8 B domain."""
    import torch
    import lala_pc_b200 as L
    from lala_pc_b200 import workloads as W
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    peak, _ = measured_peak()
    out = {}
    for name, net, bitset in (("c3_interval", W.config3(), False), ("c5_interval", W.config5(), False),
                              ("c5_bitset", None, True)):
        if net is None:
            net = prev
        prev = net
        t = L.PcTable(net.props, net.terms, net.nvars)
        cells = L.nbit_from_intervals(net.store) if bitset else None
        ms, res = [], None
        n_steps = max(3, min(args.steps, 10))
        for i in range(3 + n_steps):
            s = L.Store(values=net.store)
            if bitset:
                s.write_bits(cells)
            flush.zero_()
            torch.cuda.synchronize()
            res = t.fixpoint(s, bitset=bitset)
            if i >= 3:
                ms.append(res.device_ms)
        P, T = len(net.props), len(net.terms)
        bytes_per_sweep = 16 * P + 16 * T + (8 * int((net.props[:, 0] == 2).sum()))
        m = float(np.mean(ms))
        achieved = bytes_per_sweep * res.sweeps / (m * 1e-3) / 1e9
        e = {"ms_per_fixpoint": m, "sweeps": int(res.sweeps), "propagators": P, "terms": T, "vars": net.nvars,
             "value": res.deductions / (m * 1e-3), "unit": UNIT,
             "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                          "kernel": "k_pc_fixpoint<%s>" % ("true" if bitset else "false"),
                          "algorithmic_bytes_per_launch": bytes_per_sweep * res.sweeps}}
        if cpu:
            from oracle import oracle as O
            mdl = O.PCModel(net.formulas())
            _, st = mdl.fixpoint_bits(cells) if bitset else mdl.fixpoint(net.store)
            e["cpu_baseline"] = {"value": st.deductions / st.seconds, "unit": UNIT, "cores": 1, "kind": "port",
                                 "fixpoint_ms": st.seconds * 1e3, "sweeps": int(st.sweeps),
                                 "sample": "the whole workload: one Gauss-Seidel fixpoint of the tree-walking restatement"}
        out[name] = e
    return out


def own_search(args, cpu=True, n_split=16384, max_nodes=64):
    """the config-4 model; the EPS subproblems that survive their root fixpoint are searched
    depth-first, each in its block (propagate, branch by bisection on the widest variables, snapshot / restore on the
    device), under a node budget."""
    import lala_pc_b200 as L
    from lala_pc_b200 import sharding
    net, table, root, dec, obj = build_c4()
    dec = dec[:sharding.decision_bits(1)]
    width = root[:, 1].astype(np.int64) - root[:, 0]
    bv = [int(v) for v in np.argsort(-width, kind="stable")[:64]]
    split = L.Batch(table, n_split)
    split.init_split(root, dec, 0)
    split.set_seeds(dec)
    split.fixpoint(objective_var=obj)
    alive = np.flatnonzero((split.flags() & 1) == 0)
    roots = np.ascontiguousarray(split.read()[alive])
    split.close()
    batch = L.Batch(table, len(alive))
    batch.write(roots)
    best, best_cd = None, None
    for _ in range(4):
        r, _ = batch.search(bv, objective_var=obj, max_nodes=max_nodes, max_depth=48, want_per_store=False, change_driven=False)
        best = r if best is None or r.device_ms < best.device_ms else best
        r, _ = batch.search(bv, objective_var=obj, max_nodes=max_nodes, max_depth=48, want_per_store=False, change_driven=True)
        best_cd = r if best_cd is None or r.device_ms < best_cd.device_ms else best_cd
    out = {"workload": "config-4 model: the %d of %d EPS subproblems alive after their root fixpoint, DFS with a budget of "
                       "%d nodes each" % (len(alive), n_split, max_nodes),
           "ms": best_cd.device_ms, "nodes": int(best_cd.n_nodes), "solutions": int(best_cd.n_solutions),
           "fails": int(best_cd.n_fails), "incomplete": int(best_cd.n_incomplete), "max_depth": int(best_cd.max_depth_seen),
           "nodes_per_s": best_cd.n_nodes / (best_cd.device_ms * 1e-3), "deductions": int(best_cd.deductions),
           "dense_nodes_ms": best.device_ms, "dense_nodes_deductions": int(best.deductions),
           "value": best.deductions / (best.device_ms * 1e-3), "unit": UNIT}
    if cpu:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        sample = min(len(alive), 4 * cores)
        t0 = time.perf_counter()
        want = O.pir_search(roots[:sample], net.records, bv, objective_var=obj, max_nodes=max_nodes, max_depth=48, threads=cores)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"nodes_per_s": float(want[:, 1].sum()) / dt, "cores": cores, "kind": "port",
                               "sample": "%d of the %d subproblems" % (sample, len(alive))}
    batch.close()
    return out


def cpu_baseline_single(net, max_seconds=30.0):
    """The oracle's Gauss-Seidel fixpoint (restated reference CPU path) on the same network, 1 thread."""
    from oracle import oracle as O
    t0 = time.perf_counter()
    _, st = O.pir_fixpoint(net.store, net.records)
    return {"value": st.deductions / st.seconds, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "the whole workload: one Gauss-Seidel fixpoint (%d sweeps, %.3f s)" % (st.sweeps, st.seconds),
            "fixpoint_ms": st.seconds * 1e3, "sweeps": int(st.sweeps),
            "label": "restated reference CPU path (oracle/pir_oracle.cpp) - upstream lala-core unavailable offline"}


def reference_arm(args, rank, world):
    """--impl reference: the reference's CPU path (restated: oracle/) on the host cores."""
    if rank != 0:
        return None
    from oracle import oracle as O
    from lala_pc_b200 import workloads as W
    cores = os.cpu_count() or 1
    if world == 1 and args.workload in ("c2", "c1"):
        net = W.config1() if args.workload == "c1" else build_c2(args.scale)
        secs, ded, sweeps = [], [], 0
        for i in range(args.warmup + args.steps):
            _, st = O.pir_fixpoint(net.store, net.records)
            if i >= args.warmup:
                secs.append(st.seconds)
                ded.append(st.deductions)
                sweeps = int(st.sweeps)
        value = float(np.sum(ded) / np.sum(secs))
        cfg = {"workload": "PIR single fixpoint: %d vars / %d ternary propagators (BASELINE.json configs[%d])"
                           % (net.nvars, len(net.records), 0 if args.workload == "c1" else 1),
               "mode": "sequential Gauss-Seidel sweeps", "sweeps_per_fixpoint": sweeps}
        cpu = {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "the whole workload per step (Gauss-Seidel is sequential: 1 thread)"}
        ms = float(np.mean(secs)) * 1e3
    else:
        net = W.config4_base()
        root, st = O.pir_fixpoint(net.store, net.records)
        from lala_pc_b200 import sharding
        nbits = sharding.decision_bits(world)
        dec, obj = W.eps_decisions(net.records, root, n=24)
        sample = 4096
        secs, ded = [], []
        for i in range(args.warmup + args.steps):
            ids = sharding.shard_ids(0, world, STORES_PER_GPU)[(i * sample) % STORES_PER_GPU:][:sample]
            stores = W.eps_stores(root, dec[:nbits], 0, len(ids), ids=ids)
            _, _, _, d, sec = O.pir_batch_fixpoint(stores, net.records, threads=cores)
            if i >= args.warmup:
                secs.append(sec)
                ded.append(d)
        value = float(np.sum(ded) / np.sum(secs))
        cfg = {"workload": "batched EPS: %d-var / %d-propagator PIR model (BASELINE.json configs[3])" % (net.nvars, len(net.records)),
               "mode": "one store per host thread, Gauss-Seidel each"}
        cpu = {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d of the %d subproblem stores per step" % (sample, STORES_PER_GPU * world)}
        ms = float(np.mean(secs)) * 1e3
    return {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": cfg, "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "restated reference CPU path (oracle/): the reference itself needs un-vendored lala-core and cannot be built offline"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "c1", "c2", "c4"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink config 2 (development only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pc", action="store_true", help="skip the PC configs (3 and 5) section")
    args = ap.parse_args()
    # exactly ONE line on stdout: keep a private handle to it and point fd 1 at stderr, so that anything a C library
    # prints (e.g. NCCL's version banner) cannot end up in front of the JSON line
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "auto":
        args.workload = "c2" if world == 1 else "c4"
    args.warmup = max(args.warmup, 3) if args.impl == "own" else args.warmup

    if args.impl == "reference":
        line = reference_arm(args, rank, world)
        if line is not None:
            real_stdout.write(json.dumps(line) + "\n")
            real_stdout.flush()
        return 0

    import torch
    import lala_pc_b200 as L
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    L.device_init(local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    line = {"metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "impl": "own"}
    if args.workload in ("c1", "c2") and world == 1:
        out, net, table = own_single(args, rank)
        line.update(out)
        b = own_batched(args, rank, 1, steps=min(args.steps, 5), warmup=3)
        line["batched"] = {k: b[k] for k in ("value", "ms_per_step", "config", "batch_result", "roofline", "e2e", "e2e_split", "latency") if k in b}
        line["batched"]["unit"] = UNIT
        line["gpu_launches"] += b["gpu_launches"]
        if args.workload == "c2" and args.scale == 1.0 and not args.no_pc:
            l0 = L.launch_count()
            line["pc"] = own_pc(args, cpu=not args.no_cpu_baseline)
            line["search"] = own_search(args, cpu=not args.no_cpu_baseline)
            line["gpu_launches"] += L.launch_count() - l0
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_single(net)
    else:
        out = own_batched(args, rank, world)
        line.update(out)
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            ref = reference_arm(argparse.Namespace(**{**vars(args), "steps": 3, "warmup": 1}), 0, 1)
            line["cpu_baseline"] = ref["cpu_baseline"]
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    return 0


if __name__ == "__main__":
    sys.exit(main())
