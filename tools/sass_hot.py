"""Aggregate an ncu SASS source page (csv): instruction mix and the hottest instructions."""
import csv, sys, re
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ins = []
for i, r in enumerate(rows[2:]):
    if len(r) <= iex:
        if ins: break          # next kernel's header: keep the first captured launch only
        continue
    if not (r[isamp] or '0').isdigit(): break
    ins.append((i, r[isrc].strip(), int(r[isamp] or 0), int(r[iex] or 0)))
tot_s = sum(x[2] for x in ins); tot_e = sum(x[3] for x in ins)
byop_e, byop_s = defaultdict(int), defaultdict(int)
for i, s, sm, ex in ins:
    op = re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0]
    byop_e[op] += ex; byop_s[op] += sm
print(f"total samples {tot_s}, warp-instructions executed {tot_e}, SASS lines {len(ins)}")
print("opcode mix (executed %, samples %):")
for op, e in sorted(byop_e.items(), key=lambda kv: -kv[1])[:22]:
    print(f"  {op:12s} {e / tot_e * 100:6.2f}%  {byop_s[op] / max(1, tot_s) * 100:6.2f}%")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
print(f"top {n} instructions by samples:")
for i, s, sm, ex in sorted(ins, key=lambda x: -x[2])[:n]:
    print(f"  #{i:5d} {sm / max(1, tot_s) * 100:5.2f}%  ex={ex:>12d}  {s[:90]}")
# cumulative samples by 100-instruction windows
print("samples by window of 100 SASS lines:")
for w in range(0, len(ins), 100):
    sm = sum(x[2] for x in ins[w:w + 100]); ex = sum(x[3] for x in ins[w:w + 100])
    print(f"  [{w:5d},{w + 100:5d})  samples {sm / max(1, tot_s) * 100:5.1f}%  executed {ex / tot_e * 100:5.1f}%")
