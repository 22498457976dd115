"""Phase breakdown of the GEMM launches of one step from the I2IT_TRACE timeline (profiles/r01b_gemm_timeline_trace.txt).
usage: python profiles/trace_breakdown.py [trace.txt] [GHz]"""
import collections
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "profiles/r01b_gemm_timeline_trace.txt"
ghz = float(sys.argv[2]) if len(sys.argv) > 2 else 1.85
rows = []
for l in open(path):
    if not l.startswith("TRACE"):
        continue
    head, vals = l.split("|")
    shape = head.strip()
    v = vals.split()
    mx = int(v[v.index("max") + 1])
    s = [None] + [int(x) if x != "-" else None for x in v[:v.index("max")]]
    tiles, grid = int(re.search(r"tiles=(\d+)", shape).group(1)), int(re.search(r"grid=(\d+)", shape).group(1))
    rows.append((shape, s, mx, tiles / grid))


def report(sel, title):
    tot = collections.Counter()
    for shape, s, mx, per in sel:
        if None in (s[1], s[4], s[8], s[9], s[10], s[12]):
            continue
        tot["prologue (barriers, TMEM alloc, cluster sync)"] += s[1]
        tot["first operands landed"] += s[4] - s[1]
        tot["first tile's mainloop"] += s[8] - s[4]
        tot["first tile's epilogue"] += s[9] - s[8]
        tot["remaining tiles"] += s[10] - s[9]
        tot["join + TMEM free"] += s[12] - s[10]
        tot["= median CTA"] += s[12]
        tot["slowest CTA"] += mx
    print(f"{title}: {len(sel)} launches")
    for k, v in tot.items():
        print(f"  {k:48s} {v / ghz / 1e6:7.2f} ms")


report(rows, f"all GEMM launches of one step (cycles -> ms at {ghz} GHz)")
report([r for r in rows if r[3] < 8], "launches with < 8 tiles per CTA")
