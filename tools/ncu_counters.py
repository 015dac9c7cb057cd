"""profiles/ncu_counters.json from the ncu captures of tools/profile_r02.sh: per kernel the counters of ONE launch and the
deductions of that same launch (tools/prof_one.py), so that bench.py's roofline objects quote measured numbers.

  python tools/ncu_counters.py gpurun_out > gpurun_out/ncu_counters.json
"""
import csv
import json
import os
import subprocess
import sys

# capture name -> (key in the json, prof_one workload)
CAPS = {"eps_dense": "k_pir_group", "eps_auto": "k_pir_group_auto", "c2_dense": "k_pir_fixpoint", "c2_auto": "k_pir_dirty",
        "pc_c3": "k_pc_fixpoint_auto", "pc_c3_dense": "k_pc_fixpoint", "pc_c5": "k_pc_fixpoint_c5", "pc_c5_bits": "k_pc_fixpoint_c5_bits"}
M = {"duration_ns": "gpu__time_duration.sum", "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum",
     "inst_executed": "smsp__inst_executed.sum", "thread_inst_ratio": "smsp__thread_inst_executed_per_inst_executed.ratio",
     "smem_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
     "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts_bytes": "lts__t_bytes.sum",
     "l1_hit_pct": "l1tex__t_sector_hit_rate.pct", "lts_hit_pct": "lts__t_sector_hit_rate.pct",
     "registers": "launch__registers_per_thread", "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active"}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9,
         "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main(d):
    out = {}
    for cap, key in CAPS.items():
        rep, meta = os.path.join(d, "prof_%s.ncu-rep" % cap), os.path.join(d, "prof_%s.json" % cap)
        if not (os.path.exists(rep) and os.path.exists(meta)):
            continue
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units, r = rows[0], rows[1], rows[-1]
        e = {"kernel": r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else None}
        for k, name in M.items():
            if name in hdr:
                i = hdr.index(name)
                v = num(r[i])
                e[k] = v * SCALE.get(units[i], 1) if v is not None else None
        if e.get("dram_read") is not None and e.get("dram_write") is not None:
            e["dram_bytes"] = e["dram_read"] + e["dram_write"]
        if e.get("inst_executed") and e.get("thread_inst_ratio"):
            e["thread_inst_executed"] = e["inst_executed"] * e["thread_inst_ratio"]
        m = json.load(open(meta))
        e["deductions"] = m.get("deductions")
        e["sweeps"] = m.get("sweeps", m.get("sweeps_total"))
        e["source"] = "ncu --set full --clock-control none, one launch of `tools/prof_one.py %s` (tools/profile_r02.sh)" % cap
        out[key] = e
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out")
