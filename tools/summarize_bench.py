"""Print the numbers of a bench.py line that matter at a glance."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])


def g(x, *ks):
    for k in ks:
        if not isinstance(x, dict) or k not in x:
            return None
        x = x[k]
    return x


def f(x, p=3):
    return "-" if x is None else (round(x, p) if isinstance(x, float) else x)


print("N", d.get("n_gpus"), "| c4 weak: value", f((d.get("value") or 0) / 1e9, 1), "G ded/s, step", f(d.get("ms_per_step")), "ms, e2e",
      f(g(d, "e2e", "ms_per_step")), "ms (", f((g(d, "e2e", "value") or 0) / 1e9, 1), "G/s ), rank_kernel_ms", {k: f(v) for k, v in (d.get("rank_kernel_ms") or {}).items() if k != "per_rank"})
print("  time_to_result", {k: f(v) for k, v in (d.get("time_to_result") or {}).items() if k.endswith("_ms") or k.endswith("propagators")})
print("  roofline", {k: f(v) for k, v in (d.get("roofline") or {}).items() if k in ("bound", "frac", "warp_inst_per_deduction", "thread_inst_per_deduction", "traffic")},
      "hbm-alg frac", f(g(d, "roofline", "hbm", "frac")))
print("  result", d.get("batch_result"), "launches", d.get("gpu_launches"), "clocks", d.get("clocks"))
if isinstance(d.get("strong"), dict) and "value" in d["strong"]:
    s = d["strong"]
    print("  strong: value", f(s["value"] / 1e9, 1), "G ded/s, step", f(s["ms_per_step"]), "ms, e2e", f(g(s, "e2e", "ms_per_step")), "ms, auto",
          f(g(s, "time_to_result", "auto_ms")), "ms, rank_kernel_ms", {k: f(v) for k, v in (s.get("rank_kernel_ms") or {}).items() if k != "per_rank"})
r = d.get("resident_images")
if r:
    print("  resident images: dense", f(g(r, "dense", "ms_per_step")), "ms", f((g(r, "dense", "value") or 0) / 1e9, 1), "G/s | auto", f(g(r, "auto", "ms_per_step")), "ms")
s = d.get("single_fixpoint")
if s:
    print("c2 dense:", f(s["value"] / 1e9, 1), "G ded/s,", f(s["ms_per_step"], 4), "ms, frac hbm", f(g(s, "roofline", "frac")), "frac l2", f(g(s, "roofline", "frac_of_l2_copy_peak")),
          "| e2e", f(g(s, "e2e", "ms_per_step")), "ms auto-e2e", f(g(s, "e2e", "auto_ms_per_step")), "ms")
    print("  time_to_fixpoint", {k: f(v, 4) for k, v in s["time_to_fixpoint"].items() if k != "what"})
    print("  cpu", f(g(s, "cpu_baseline", "fixpoint_ms"), 1), "ms")
if d.get("pc"):
    print("pc (dense ms, sweeps, hbm frac, default-mode ms):", {k: (f(v["ms_per_fixpoint"], 4), v["sweeps"], f(g(v, "roofline", "frac")), f(g(v, "time_to_fixpoint", "auto_ms"), 4)) for k, v in d["pc"].items()})
print("cpu_baseline", d.get("cpu_baseline"))
if d.get("search"):
    s = d["search"]
    print("search:", f(s.get("ms")), "ms", s.get("nodes"), "nodes; fails", s.get("fails"), "solutions", s.get("solutions"))
    b = s.get("backtracking")
    if b:
        print("  backtracking:", f(b.get("ms")), "ms (dense", f(b.get("dense_nodes_ms")), "cd", f(b.get("change_driven_nodes_ms")), ")", b.get("nodes"), "nodes; fails",
              b.get("fails"), "solutions", b.get("solutions"), "incomplete", b.get("incomplete"), "| cpu nodes/s", f(g(b, "cpu_baseline", "nodes_per_s"), 0),
              "gpu nodes/s", f(b.get("nodes_per_s"), 0))
