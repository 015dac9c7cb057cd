#!/usr/bin/env python
"""bench.py — propagator deductions/s of the fixpoint hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic input. The line the driver scales (top-level `value`,
the same workload at EVERY N, so that v_N / (N * v_1) means something) is BASELINE.json configs[3], the one configuration
the metric quotes at 1/2/4/8 GPUs: batched EPS, 65,536 subproblem stores of the 2k-variable / 10k-propagator PIR model
PER GPU (weak scaling; each rank a uniform sample of the subproblem ids), one thread group per store, no inter-GPU
traffic inside the fixpoints; the ranks' reduction records (32 B each) are exchanged by the kernel itself over NVLink
peer memory (fallback: ONE NCCL all-reduce of 3 + N int64) per step. Next to it, in the same line:
  strong            the 65,536 subproblems of configs[3] in total, 65,536 / N per rank
  single_fixpoint   (N = 1) BASELINE.json configs[1]: one fixpoint of the 1M-variable / 5M-propagator network - dense
                    sweeps with their HBM roofline fraction, time to fixpoint of the change-driven mode with
                    entailment-driven elimination, end to end from host buffers, the CPU baseline
  resident_images   (N = 1) lpc_batch_fixpoint on 1 GiB of store images resident in HBM (round 1's headline kernel shape)
  pc, search        (N = 1) configs[2] and [4] (PC, interval and bitset stores), in-kernel search
`value` is measured with the inputs (root store, decision list, subproblem ids, table) resident in HBM, CUDA events on
the launching stream, every sweep evaluating every propagator (LPC_MODE_SWEEP: the reference's work unit). `e2e` is the
same work through the C-ABI call a solver makes, lpc_eps_solve_host, with pinned HOST buffers: root + decisions + ids in,
flags + reduction record + the compacted non-failed stores out, copies inside the timed region. `time_to_result` reports
the default mode (propagators entailed on the root dropped). `--impl reference` times the CPU restatement of the
reference's Gauss-Seidel path (oracle/) on the host cores, on the same workload.
"""
import argparse
import ctypes
import importlib.util
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
STORES = 65536                    # BASELINE.json configs[3]
SURVIVOR_CAP = 8192               # non-failed stores kept per call (about 4 % of the subproblems survive)


def host_module(name):
    """workloads.py / sharding.py are pure numpy host logic: the reference arm loads them by path, WITHOUT importing the
    package (whose __init__ loads liblpc.so), so that no product code is mapped while the CPU path is timed."""
    spec = importlib.util.spec_from_file_location("_lpc_host_" + name, os.path.join(ROOT, "lala-pc_b200", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_counters(kernel):
    """Per-launch counters of a kernel from the committed ncu capture of this round (profiles/ncu_counters.json, written by
    tools/profile.sh from the .ncu-rep files it captures - never typed in by hand): DRAM bytes, instructions executed and
    the deductions of the profiled launch."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_counters.json")) as f:
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
def c4_config(world, scaling):
    """The `config` object of BOTH arms (own and reference): identical dicts for identical work."""
    nbits = 16 + (max(0, (world - 1).bit_length()) if scaling == "weak" else 0)
    per = STORES if scaling == "weak" else STORES // world
    return {"workload": "batched EPS: %d subproblem stores %s of a 2000-var / 10000-propagator PIR model, one thread "
                        "group per store (BASELINE.json configs[3])" % (STORES, "per GPU" if scaling == "weak" else "in total"),
            "scaling": scaling, "stores_per_gpu": per, "stores_total": per * world, "decision_bits": nbits,
            "mode": "every sweep evaluates every propagator (the reference's work unit)",
            "l2": "GPU arm: L2 flushed between timed steps (a 512 MiB write); the step's inputs are the 16 KB root store, the "
                  "subproblem ids and the 80 KB table"}


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


class _DevView:   # a device int64 vector of the library, viewed as a torch tensor (no copy)
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def _rank_stats(x, dist, world):
    """min / max / mean over ranks of a per-rank float."""
    import torch
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    if dist is None:
        return {"min": x, "max": x, "mean": x}
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    v = [float(o[0]) for o in out]
    return {"min": min(v), "max": max(v), "mean": float(np.mean(v)), "per_rank": v}


def own_eps(args, rank, world, scaling, ctx, steps=None, warmup=None, check=True):
    """Batched EPS on this rank through the EPS-native call; one all-reduce per step when world > 1.
    Returns value (device-timed, inputs resident), e2e (host buffers), time_to_result (default mode), checks."""
    import torch
    import lala_pc_b200 as L
    from lala_pc_b200 import sharding, workloads as W
    steps = args.steps if steps is None else steps
    warmup = args.warmup if warmup is None else warmup
    net, table, root, dec24, obj = ctx
    cfg = c4_config(world, scaling)
    n = cfg["stores_per_gpu"]
    dec = dec24[:cfg["decision_bits"]]
    ids = sharding.shard_ids(rank, world, n) if scaling == "weak" else sharding.strong_shard_ids(rank, world, STORES)
    dist = None
    if world > 1:
        import torch.distributed as dist
    eps = L.Eps(table, n, survivor_cap=SURVIVOR_CAP)
    eps.set_rank(rank, world)
    # The record exchange: fused into the batch kernel over NVLink peer memory when the GPUs can reach each other (CUDA IPC
    # handles gathered once through torch.distributed), else one NCCL all-reduce per step. LPC_P2P=0 forces NCCL.
    p2p = False
    if dist is not None and os.environ.get("LPC_P2P", "1") != "0":
        handles = [None] * world
        dist.all_gather_object(handles, eps.peer_export())
        try:
            eps.peer_connect(rank, world, handles)
            ok = 1
        except L.LpcError:
            ok = 0
        agree = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        p2p = bool(int(agree[0]))
        assert p2p or not ok, "peer access on some ranks only"
    payload = torch.as_tensor(_DevView(*eps.payload), device="cuda")
    flush = ctx_flush()
    eps.upload(root, dec, ids=ids)
    out = {"config": cfg}

    def timed(mode, k_steps, k_warm):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k_steps)]
        res, l0, kms = None, L.launch_count(), []
        for i in range(k_warm + k_steps):
            flush.zero_()
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            if i == k_warm:
                l0 = L.launch_count()
            if i >= k_warm:
                ev[i - k_warm][0].record()
            eps.run_async(objective_var=obj, mode=mode)     # with peers connected the exchange is part of the call
            if not p2p:
                sharding.allreduce_payload(payload, dist)
            if i >= k_warm:
                ev[i - k_warm][1].record()
            res = eps.collect()
            if i >= k_warm:
                kms.append(res.device_ms)
            torch.cuda.synchronize()
        ms = [a.elapsed_time(b) for a, b in ev]
        return ms, res, L.launch_count() - l0, float(np.mean(kms))

    with ClockSampler(rank) as clk:
        ms, res, launches, kernel_ms = timed(L.MODE_SWEEP, steps, warmup)
    my_ms = float(np.mean(ms))
    tot = torch.tensor([float(np.sum(ms)), float(res.deductions), float(res.sweeps_total)], dtype=torch.float64, device="cuda")
    tmax = tot.clone()
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)          # time: max over ranks
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)           # work: sum over ranks
    step_ms = float(tmax[0]) / steps
    reduced = sharding.fold_payload(payload.cpu().tolist())
    out.update({"value": float(tot[1]) / (step_ms * 1e-3), "ms_per_step": step_ms, "gpu_launches": int(launches),
                "rank_ms": _rank_stats(my_ms, dist, world), "rank_kernel_ms": _rank_stats(kernel_ms, dist, world),
                "clocks": clk.summary(),
                "batch_result": {"n_solution": reduced[0], "n_bot": reduced[1], "n_unknown": reduced[2], "best_bound": reduced[3],
                                 "sweeps_total": float(tot[2]), "max_sweeps_seen": res.max_sweeps_seen,
                                 "deductions_per_step": float(tot[1])},
                "measurement": {"timing": "CUDA events on the launching stream around lpc_eps_run_async + the all-reduce, max over ranks",
                                "resident": "root, decision list, ids, the record table and what is derived from them alone (the packed "
                                            "table, built by the first warm-up run after the upload) stay in HBM across steps; every "
                                            "step generates and solves all subproblems anew; the e2e figure uploads and packs per call",
                                "l2": "flushed between steps (512 MiB write); the kernel's inputs are the 16 KB root store, the "
                                      "subproblem ids and the 80 KB table",
                                "sharding": "consecutive ids" if world == 1 else "each rank a uniform sample of the id space (sharding.py)",
                                "collective": "none" if world == 1 else
                                              ("fused into the batch kernel: its last block writes the rank's 32-byte record into every "
                                               "peer's inbox over NVLink (CUDA IPC peer memory), a one-thread kernel folds the inbox; no "
                                               "collective library call per step" if p2p else
                                               "one NCCL all-reduce (SUM) of the %d x int64 payload per step" % (3 + world)),
                                "seed": net.meta["seed"]}})
    # ---- run-time checks (outside the timed region) -----------------------------------------------------------------------
    if check:
        flags, surv, sidx = eps.download()
        bot = (flags & 1) != 0
        sol = ((flags & 2) != 0) & ~bot
        mine = [int(sol.sum()), int(bot.sum()), int((~bot & ~sol).sum())]
        assert mine == [res.n_solution, res.n_bot, res.n_unknown], "rank %d: reduction record != sum of its own flags" % rank
        assert res.n_survivors == int((~bot).sum()), "rank %d: survivor count != non-failed flags" % rank
        best = int(surv[:, obj, 0].min()) if len(surv) and res.n_survivors <= SURVIVOR_CAP else res.best_bound
        assert best == res.best_bound, "rank %d: best bound != min over its survivors" % rank
        if dist is not None:   # the all-reduced payload equals the fold of the per-rank records
            rec = torch.tensor(mine + [res.best_bound], dtype=torch.int64, device="cuda")
            allrec = [torch.zeros_like(rec) for _ in range(world)]
            dist.all_gather(allrec, rec)
            a = torch.stack(allrec).cpu().numpy()
            assert reduced == [int(a[:, 0].sum()), int(a[:, 1].sum()), int(a[:, 2].sum()), int(a[:, 3].min())], \
                "rank %d: all-reduced record != fold of the per-rank records" % rank
        if rank == 0:          # a sample of this rank's shard against the CPU checker
            from oracle import oracle as O
            k = 256
            want, wflags, _, _, _ = O.pir_batch_fixpoint(W.eps_stores(root, dec, 0, k, ids=ids[:k]), net.records, threads=os.cpu_count() or 8)
            assert np.array_equal(flags[:k], wflags), "rank 0: flags of the first %d subproblems differ from the oracle" % k
            sel = sidx < k
            assert np.array_equal(surv[sel], want[sidx[sel]]), "rank 0: surviving stores differ from the oracle"
            if res.n_survivors <= SURVIVOR_CAP:
                assert sorted(sidx[sel].tolist()) == np.flatnonzero((wflags & 1) == 0).tolist()
        out["checks"] = "per-rank record == sum of own flags; all-reduced record == fold of per-rank records; rank 0: " \
                        "256-subproblem sample == oracle (flags, surviving stores)"
    # ---- end to end: the call a solver makes, host buffers in and out ------------------------------------------------------
    nv = net.nvars
    root_p = torch.from_numpy(np.ascontiguousarray(root, dtype=np.int32)).pin_memory()
    ids_p = torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int64)).pin_memory()
    flags_p = torch.empty(n, dtype=torch.uint8).pin_memory()
    surv_p = torch.empty((SURVIVOR_CAP, nv, 2), dtype=torch.int32).pin_memory()
    sidx_p = torch.empty(SURVIVOR_CAP, dtype=torch.int32).pin_memory()

    def e2e(mode, k_steps):
        t_sum, d_sum, nw = 0.0, 0, 0
        for i in range(2 + k_steps):
            flush.zero_()
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r, nw = eps.solve_host(root_p.data_ptr(), dec, ids=ids_p.data_ptr(), n=n, objective_var=obj, flags=flags_p.data_ptr(),
                                   survivors=surv_p.data_ptr(), survivor_index=sidx_p.data_ptr(), mode=mode)
            if dist is not None and not p2p:
                sharding.allreduce_payload(payload, dist)
                torch.cuda.synchronize()
            t1 = time.perf_counter()
            if i >= 2:
                t_sum += t1 - t0
                d_sum += r.deductions
        tt = torch.tensor([t_sum], dtype=torch.float64, device="cuda")
        dd = torch.tensor([float(d_sum)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(dd, op=dist.ReduceOp.SUM)
        return float(tt[0]) / k_steps * 1e3, float(dd[0]) / float(tt[0]), nw

    ms_e, v_e, nw = e2e(L.MODE_SWEEP, steps)
    out["e2e"] = {"value": v_e, "unit": UNIT, "ms_per_step": ms_e, "steps": steps,
                  "h2d_bytes_per_step": int(nv * 8 + 4 * len(dec) + 8 * n), "d2h_bytes_per_step": int(n + nw * (nv * 8 + 4) + 1024),
                  "what": "lpc_eps_solve_host per rank: root store + decision list + subproblem ids from pinned host memory -> "
                          "subproblems generated on the chip -> fixpoints -> flags, reduction record and the compacted "
                          "non-failed stores (%d here) back to pinned host memory; wall clock, max over ranks" % nw}
    # ---- time to result in the default mode: propagators entailed on the root are dropped -------------
    ms_a, res_a, _, _ = timed(L.MODE_AUTO, min(steps, 10), 2)
    ms_ae, v_ae, _ = e2e(L.MODE_AUTO, min(steps, 10))
    ta = torch.tensor([float(np.mean(ms_a))], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(ta, op=dist.ReduceOp.MAX)
    out["time_to_result"] = {"dense_ms": step_ms, "dense_e2e_ms": ms_e, "auto_ms": float(ta[0]), "auto_e2e_ms": ms_ae,
                             "auto_deductions_per_rank": int(res_a.deductions), "dense_deductions_per_rank": int(res.deductions),
                             "auto_live_propagators": int(res_a.n_live_records), "propagators": len(net.records),
                             "auto_value": float(res_a.deductions) * world / (float(ta[0]) * 1e-3), "auto_e2e_value": v_ae,
                             "what": "LPC_MODE_AUTO: same flags, stores and record (tests/test_gpu_eps.py), fewer deductions"}
    if p2p:   # nobody may still have an inbox mapped when its owner frees it
        eps.peer_disconnect()
        dist.barrier()
    eps.close()
    return out


_FLUSH = None


def ctx_flush():
    global _FLUSH
    if _FLUSH is None:
        import torch
        _FLUSH = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    return _FLUSH


def issue_roofline(value_per_gpu, clocks):
    """The batch kernel never leaves the SM (table and stores in shared memory), so its ceiling is instruction issue: one
    warp instruction per scheduler and clock. achieved = warp instructions per deduction (ncu, committed capture) x
    deductions/s; peak = 148 SMs x 4 schedulers x SM clock."""
    c = ncu_counters("k_pir_group")
    mhz = (clocks or {}).get("sm_mhz") or 1965
    peak = 148 * 4 * mhz * 1e6
    r = {"bound": "issue", "unit": "warp-inst/s", "peak": peak, "kernel": "k_pir_group<false, 8, true>",
         "peak_source": "148 SMs x 4 schedulers x %d MHz (SM clock sampled during the timed region)" % mhz,
         "achieved": None, "frac": None, "traffic": None}
    if c and c.get("inst_executed") and c.get("deductions"):
        ipd = c["inst_executed"] / c["deductions"]
        r.update({"achieved": ipd * value_per_gpu, "frac": ipd * value_per_gpu / peak, "warp_inst_per_deduction": ipd,
                  "thread_inst_per_deduction": c.get("thread_inst_executed", 0) / c["deductions"],
                  "traffic": c.get("dram_bytes"), "smem_wavefronts_per_deduction": (c.get("smem_wavefronts") or 0) / c["deductions"],
                  "counters_source": c.get("source")})
    peak_hbm, src = measured_peak()
    r["hbm"] = {"achieved": BYTES_PER_DEDUCTION * value_per_gpu / 1e9, "peak": peak_hbm, "unit": "GB/s",
                "frac": BYTES_PER_DEDUCTION * value_per_gpu / 1e9 / peak_hbm, "peak_source": src,
                "note": "algorithmic bytes (40 B per deduction) over the HBM copy peak: NOT a bound for this kernel - records "
                        "and bounds are read from shared memory, real DRAM traffic per launch is `traffic`"}
    return r


def own_resident(args, ctx, steps=5, warmup=3):
    """Round 1's headline shape, kept for continuity: lpc_batch_fixpoint over 1 GiB of store images resident in HBM."""
    import torch
    import lala_pc_b200 as L
    net, table, root, dec24, obj = ctx
    dec = dec24[:16]
    batch = L.Batch(table, STORES)
    out = {}
    for name, mode in (("dense", L.MODE_SWEEP), ("auto", L.MODE_AUTO)):
        ms, res = [], None
        for i in range(warmup + steps):
            batch.init_split(root, dec, 0)                  # untimed: 1 GiB of images, larger than L2
            torch.cuda.synchronize()
            res = batch.fixpoint(objective_var=obj, mode=mode)
            if i >= warmup:
                ms.append(res.device_ms)
        m = float(np.mean(ms))
        out[name] = {"ms_per_step": m, "value": res.deductions / (m * 1e-3), "deductions": int(res.deductions),
                     "sweeps_total": int(res.sweeps_total)}
    out["what"] = "lpc_batch_init_split (untimed) + lpc_batch_fixpoint on 65,536 store images (1 GiB) resident in HBM; failed " \
                  "stores are not written back"
    batch.close()
    return out


def own_single(args, rank):
    """BASELINE.json configs[1] (and [0]): one fixpoint of the single-store network per step."""
    import torch
    import lala_pc_b200 as L
    from lala_pc_b200 import workloads as W
    net = W.config2(args.scale)
    name = "PIR single fixpoint: %d vars / %d ternary propagators (BASELINE.json configs[1])" % (net.nvars, len(net.records))
    table = L.Table(net.records, net.nvars)
    P = len(net.records)
    steps, warmup = min(args.steps, 20), 3
    flush = ctx_flush()

    def run_mode(mode, k):
        res = []
        for i in range(warmup + k):
            s = L.Store(values=net.store)
            flush.zero_()
            torch.cuda.synchronize()
            r = L.fixpoint(table, s, mode=mode)
            if i >= warmup:
                res.append(r.as_dict())
            s.close()
        return res

    l0 = L.launch_count()
    dense = run_mode(L.MODE_SWEEP, steps)
    auto = run_mode(L.MODE_AUTO, steps)
    dev_ms = [r["device_ms"] for r in dense]
    ded = [r["deductions"] for r in dense]
    per_launch_ms = float(np.mean(dev_ms))
    value = float(np.sum(ded)) / (float(np.sum(dev_ms)) * 1e-3)
    auto_ms = float(np.mean([r["device_ms"] for r in auto]))
    # end to end through the host-buffer entry point (default mode: what a caller gets)
    pinned = torch.empty((net.nvars, 2), dtype=torch.int32).pin_memory()
    src = torch.from_numpy(net.store)
    e2e = {}
    for mname, mode in (("dense", L.MODE_SWEEP), ("auto", L.MODE_AUTO)):
        t_sum, d_sum = 0.0, 0
        for i in range(warmup + steps):
            pinned.copy_(src)
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = L.fixpoint_host(table, pinned.data_ptr(), mode=mode)
            t1 = time.perf_counter()
            if i >= warmup:
                t_sum += t1 - t0
                d_sum += r.deductions
        e2e[mname] = {"value": d_sum / t_sum, "ms_per_step": t_sum / steps * 1e3}
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_DEDUCTION * float(np.mean(ded)) / (per_launch_ms * 1e-3) / 1e9
    l2_gbs = float(L.measure_l2_copy_gbs(32 << 20, 20))
    c = ncu_counters("k_pir_fixpoint") or {}
    # config 1 (the reference's CPU-runnable case): time to fixpoint
    net1 = W.config1()
    t1 = L.Table(net1.records, net1.nvars)
    c1 = {}
    for mname, mode in (("dense", L.MODE_SWEEP), ("auto", L.MODE_AUTO)):
        best = None
        for _ in range(6):
            s = L.Store(values=net1.store)
            r = L.fixpoint(t1, s, mode=mode)
            best = r if best is None or r.device_ms < best.device_ms else best
        c1[mname + "_ms"] = float(best.device_ms)
        c1[mname + "_sweeps"] = int(best.sweeps)
    out = {
        "workload": name, "value": value, "unit": UNIT, "ms_per_step": per_launch_ms, "steps": steps,
        "mode": "dense sweeps (LPC_MODE_SWEEP): every sweep evaluates every propagator", "sweeps_per_fixpoint": dense[0]["sweeps"],
        "l2": "flushed between steps (512 MiB write)", "seed": net.meta["seed"],
        "time_to_fixpoint": {"dense_ms": per_launch_ms, "dense_deductions": int(np.mean(ded)),
                             "auto_ms": auto_ms, "auto_deductions": int(np.mean([r["deductions"] for r in auto])),
                             "auto_iterations": auto[0]["sweeps"], "auto_dense_iterations": auto[0]["dense_sweeps"],
                             "auto_value": float(np.mean([r["deductions"] for r in auto])) / (auto_ms * 1e-3),
                             "what": "LPC_MODE_AUTO: change-driven sweeps + entailment-driven elimination (groups of 64 "
                                     "propagators that PIR::ask entails are struck for the rest of the fixpoint)",
                             "config1": c1},
        "e2e": {"value": e2e["dense"]["value"], "unit": UNIT, "ms_per_step": e2e["dense"]["ms_per_step"],
                "auto_ms_per_step": e2e["auto"]["ms_per_step"], "auto_value": e2e["auto"]["value"],
                "h2d_bytes_per_step": net.nvars * 8, "d2h_bytes_per_step": net.nvars * 8 + 64,
                "what": "lpc_fixpoint_host: pinned host store -> device, fixpoint, store -> host (table resident)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": c.get("dram_bytes"), "traffic_source": c.get("source"), "peak_source": peak_src,
                     "kernel": "k_pir_fixpoint<false, false, 2, 3, false>",
                     "algorithmic_bytes_per_launch": BYTES_PER_DEDUCTION * float(np.mean(ded)),
                     "l2_copy_peak": l2_gbs, "frac_of_l2_copy_peak": achieved / l2_gbs if l2_gbs > 0 else None,
                     "l2_copy_peak_source": "lpc_measure_l2_copy_gbs: read + write GB/s of a 32 MB buffer pair resident in L2"},
        "gpu_launches": int(L.launch_count() - l0),
    }
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_single(net)
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
        n_steps = max(3, min(args.steps, 10))
        runs = {}
        for mode in (L.MODE_SWEEP, L.MODE_AUTO):
            ms, res = [], None
            for i in range(3 + n_steps):
                s = L.Store(values=net.store)
                if bitset:
                    s.write_bits(cells)
                flush.zero_()
                torch.cuda.synchronize()
                res = t.fixpoint(s, bitset=bitset, mode=mode)
                if i >= 3:
                    ms.append(res.device_ms)
            runs[mode] = (float(np.mean(ms)), res, s.read())
        P, T = len(net.props), len(net.terms)
        bytes_per_sweep = 16 * P + 16 * T + (8 * int((net.props[:, 0] == 2).sum()))
        m, res, final = runs[L.MODE_SWEEP]
        ma, resa, finala = runs[L.MODE_AUTO]
        assert np.array_equal(final, finala), "PC: the two modes disagree"
        achieved = bytes_per_sweep * res.sweeps / (m * 1e-3) / 1e9
        e = {"ms_per_fixpoint": m, "sweeps": int(res.sweeps), "propagators": P, "terms": T, "vars": net.nvars,
             "value": res.deductions / (m * 1e-3), "unit": UNIT,
             "time_to_fixpoint": {"dense_ms": m, "auto_ms": ma, "dense_deductions": int(res.deductions),
                                  "auto_deductions": int(resa.deductions),
                                  "what": "LPC_MODE_AUTO skips the lane tiles whose operands have not moved since their "
                                          "last evaluation: same store", "auto_sweeps": int(resa.sweeps)},
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
    out["backtracking"] = own_search_complete(cpu)
    return out


def own_search_complete(cpu=True, nbits=12, max_nodes=2048):
    """A search that runs to the bottom: a 200-variable / 500-propagator planted-solution model with narrow domains,
    4,096 EPS subproblems, depth-first until every leaf is a solution or a failure (restore on every failure). The
    per-subproblem records of a sample are compared with the CPU restatement at run time."""
    import lala_pc_b200 as L
    from lala_pc_b200 import workloads as W
    net = W.pir_network(200, 500, seed=77, width=6)
    table = L.Table(net.records, net.nvars)
    s = L.Store(values=net.store)
    L.fixpoint(table, s)
    root = s.read()
    dec, obj = W.eps_decisions(net.records, root, n=nbits, min_degree=2)
    stores = W.eps_stores(root, dec, 0, 1 << len(dec))
    width = root[:, 1].astype(np.int64) - root[:, 0]
    bv = [int(v) for v in np.argsort(-width, kind="stable")]
    batch = L.Batch(table, len(stores))
    runs = {}
    for cd in (False, True, None):   # dense nodes, change-driven nodes, the library's default (by table size)
        for _ in range(3):
            batch.write(stores)
            r, per = batch.search(bv, objective_var=obj, max_nodes=max_nodes, max_depth=96, change_driven=cd)
            if cd not in runs or r.device_ms < runs[cd][0].device_ms:
                runs[cd] = (r, per)
    (best, got), dense, chd = runs[None], runs[False][0], runs[True][0]
    out = {"workload": "200-var / 500-propagator model, %d EPS subproblems, DFS to the leaves (budget %d nodes each)"
                       % (len(stores), max_nodes),
           "ms": best.device_ms, "nodes": int(best.n_nodes), "solutions": int(best.n_solutions), "fails": int(best.n_fails),
           "incomplete": int(best.n_incomplete), "best_bound": int(best.best_bound), "max_depth": int(best.max_depth_seen),
           "nodes_per_s": best.n_nodes / (best.device_ms * 1e-3), "deductions": int(best.deductions),
           "dense_nodes_ms": dense.device_ms, "dense_nodes_deductions": int(dense.deductions),
           "change_driven_nodes_ms": chd.device_ms, "change_driven_nodes_deductions": int(chd.deductions)}
    assert best.n_fails > 0 and best.n_solutions > 0, "the backtracking workload neither failed nor solved anything"
    assert (dense.n_nodes, dense.n_solutions, dense.n_fails) == (chd.n_nodes, chd.n_solutions, chd.n_fails)
    if cpu:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        sample = min(len(stores), 512)
        t0 = time.perf_counter()
        want = O.pir_search(stores[:sample], net.records, bv, objective_var=obj, max_nodes=max_nodes, max_depth=96, threads=cores)
        dt = time.perf_counter() - t0
        assert np.array_equal(got[:sample], want), "search records differ from the CPU restatement"
        out["cpu_baseline"] = {"nodes_per_s": float(want[:, 1].sum()) / dt, "cores": cores, "kind": "port",
                               "sample": "%d of the %d subproblems" % (sample, len(stores))}
        out["checked"] = "per-subproblem {solutions, nodes, fails, best, incomplete} of %d subproblems == CPU restatement" % sample
    batch.close()
    return out


def cpu_baseline_single(net):
    """The oracle's Gauss-Seidel fixpoint (restated reference CPU path) on the same network, 1 thread."""
    from oracle import oracle as O
    _, st = O.pir_fixpoint(net.store, net.records)
    return {"value": st.deductions / st.seconds, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "the whole workload: one Gauss-Seidel fixpoint (%d sweeps, %.3f s)" % (st.sweeps, st.seconds),
            "fixpoint_ms": st.seconds * 1e3, "sweeps": int(st.sweeps),
            "label": "restated reference CPU path (oracle/pir_oracle.cpp) - upstream lala-core unavailable offline"}


def cpu_batched(world, scaling, steps, warmup, sample=4096):
    """The reference's CPU path on the batched workload: one store per host thread, Gauss-Seidel each, on a bounded sample
    of the subproblems per step. Imports nothing of the product (host_module)."""
    from oracle import oracle as O
    W, sharding = host_module("workloads"), host_module("sharding")
    cores = os.cpu_count() or 1
    cfg = c4_config(world, scaling)
    net = W.config4_base()
    root, _ = O.pir_fixpoint(net.store, net.records)
    dec, obj = W.eps_decisions(net.records, root, n=24)
    dec = dec[:cfg["decision_bits"]]
    n = cfg["stores_per_gpu"]
    all_ids = sharding.shard_ids(0, world, n) if scaling == "weak" else sharding.strong_shard_ids(0, world, STORES)
    sample = min(sample, n)
    secs, ded = [], []
    for i in range(warmup + steps):
        ids = all_ids[(i * sample) % n:][:sample]
        stores = W.eps_stores(root, dec, 0, len(ids), ids=ids)
        _, _, _, d, sec = O.pir_batch_fixpoint(stores, net.records, threads=cores)
        if i >= warmup:
            secs.append(sec)
            ded.append(d)
    value = float(np.sum(ded) / np.sum(secs))
    cpu = {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": "%d of the %d subproblem stores per step, one store per host thread" % (sample, cfg["stores_total"]),
           "label": "restated reference CPU path (oracle/pir_oracle.cpp) - upstream lala-core unavailable offline"}
    return cfg, cpu, value, float(np.mean(secs)) * 1e3


def reference_arm(args, rank, world):
    """--impl reference: the reference's CPU path (restated: oracle/) on the host cores, same workload as the own arm."""
    if rank != 0:
        return None
    cfg, cpu, value, ms = cpu_batched(world, "weak", args.steps, args.warmup)
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
    ap.add_argument("--scale", type=float, default=1.0, help="shrink config 2 (development only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="N = 1: skip the single_fixpoint / resident_images / pc / search sections")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling section")
    args = ap.parse_args()
    # exactly ONE line on stdout: keep a private handle to it and point fd 1 at stderr, so that anything a C library
    # prints (e.g. NCCL's version banner) cannot end up in front of the JSON line
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        line = reference_arm(args, rank, world)
        if line is not None:
            real_stdout.write(json.dumps(line) + "\n")
            real_stdout.flush()
        return 0

    args.warmup = max(args.warmup, 3)
    import torch
    import lala_pc_b200 as L
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    L.device_init(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    line = {"metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "impl": "own"}
    ctx = build_c4()
    out = own_eps(args, rank, world, "weak", ctx)
    line.update(out)
    line["roofline"] = issue_roofline(out["value"] / world, out["clocks"])
    if world > 1 and not args.no_strong:
        s = own_eps(args, rank, world, "strong", ctx, check=True)
        line["strong"] = {k: s[k] for k in ("config", "value", "ms_per_step", "rank_ms", "rank_kernel_ms", "e2e", "time_to_result", "batch_result", "checks") if k in s}
        line["strong"]["unit"] = UNIT
        line["gpu_launches"] += s["gpu_launches"]
    if world == 1:
        line["strong"] = {"note": "at one GPU the strong and the weak workload coincide (65,536 subproblems)"}
        if not args.no_extras:
            l0 = L.launch_count()
            line["resident_images"] = own_resident(args, ctx)
            line["single_fixpoint"] = own_single(args, rank)
            if args.scale == 1.0:
                line["pc"] = own_pc(args, cpu=not args.no_cpu_baseline)
                line["search"] = own_search(args, cpu=not args.no_cpu_baseline)
            line["gpu_launches"] += L.launch_count() - l0
        if not args.no_cpu_baseline:
            _, cpu, _, _ = cpu_batched(1, "weak", 3, 1)
            line["cpu_baseline"] = cpu
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    return 0


if __name__ == "__main__":
    sys.exit(main())
