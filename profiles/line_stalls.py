"""Per-source-line warp-stall summary of an .ncu-rep captured with `--set full --import-source on` (kernels built -lineinfo).
usage: python profiles/line_stalls.py rep.ncu-rep [top_n] > out.txt      (reads the report here with `ncu -i`, no GPU needed)"""
import csv
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True,
                     text=True).stdout
launch, cur_file, hdr, lines, fn = -1, None, None, {}, ""


def flush():
    if not lines:
        return
    tot = sum(v[0] for v in lines.values()) or 1
    print(f"== launch {launch}: {fn[:90]}  total_samples {tot}")
    for (f, ln), (n, src, reasons) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        rs = ", ".join(f"{k}={v}" for k, v in sorted(reasons.items(), key=lambda kv: -kv[1])[:3])
        print(f"  {100.0 * n / tot:5.1f}%  {f.split('/')[-1]}:{ln:<5} {src.strip()[:100]}   [{rs}]")


seen_files = set()
for r in csv.reader(raw.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        if r[1] in seen_files:          # the file list restarts with every profiled launch
            flush(); lines, seen_files = {}, set()
        if not seen_files:
            launch += 1
        seen_files.add(r[1]); cur_file = r[1]
        continue
    if r[0] == "Function Name":
        fn = r[1]; continue
    if r[0] == "Line No":
        hdr = r; i_samp = hdr.index("# Samples")
        stall_cols = [j for j, h in enumerate(hdr) if h.startswith("stall_") and "(Not" in h]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":                       # a CUDA source line: ncu already aggregated its SASS rows
        n = int(r[i_samp]) if r[i_samp].isdigit() else 0
        if n:
            reasons = {hdr[j].split(" ")[0].replace("stall_", ""): int(r[j]) for j in stall_cols if r[j].isdigit() and int(r[j])}
            key = (cur_file, r[0])
            if key in lines:
                lines[key][0] += n
            else:
                lines[key] = [n, r[1], reasons]
flush()
