"""ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file X) -> per-kernel
summary CSV (launches, time, share of the step, DRAM bytes).  usage: python profiles/summarize_launch_list.py in.csv out.csv "note" """
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
iK, iM, iV, iI = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = collections.defaultdict(lambda: [set(), 0.0, 0.0])
for r in rows[1:]:
    name = re.sub(r"^void ", "", r[iK]).split("(")[0]
    a = per[name]
    a[0].add(r[iI])
    v = float(r[iV].replace(",", ""))
    if r[iM] == "gpu__time_duration.sum": a[1] += v / 1e6            # ns -> ms
    elif r[iM].startswith("dram__bytes"): a[2] += v / 1e9
tot = sum(a[1] for a in per.values())
with open(sys.argv[2], "w") as f:
    f.write("# " + (sys.argv[3] if len(sys.argv) > 3 else "") + "\n")
    f.write("kernel,launches,time_ms,share,dram_GB,dram_GB_per_s\n")
    for k, a in sorted(per.items(), key=lambda t: -t[1][1]):
        f.write(f"{k},{len(a[0])},{a[1]:.3f},{a[1] / tot:.4f},{a[2]:.3f},{a[2] / a[1] * 1e3 if a[1] else 0:.1f}\n")
    f.write(f"TOTAL,{sum(len(a[0]) for a in per.values())},{tot:.3f},1.0,{sum(a[2] for a in per.values()):.3f},\n")
print(open(sys.argv[2]).read())
