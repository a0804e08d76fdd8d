"""Per-source-line view of an .ncu-rep captured with --import-source on (kernels compiled with -lineinfo):
warp instructions, share of the kernel, average active threads and stall samples per CUDA source line.
python tools/ncu_lines.py REPORT.ncu-rep [top N]   (needs `ncu` on PATH; no GPU)"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, check=True).stdout.decode(errors="replace")
rows = list(csv.reader(raw.splitlines()))
cur, hdr, lines = "?", None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-":
        d = dict(zip(hdr, r))
        try:
            lines.append((int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), int(d["# Samples"] or 0),
                          cur, int(r[0]), r[1].strip()))
        except ValueError:
            pass
tot = sum(x[0] for x in lines) or 1
tth = sum(x[1] for x in lines)
tsm = sum(x[2] for x in lines) or 1
print(f"total warp instructions {tot}, thread instructions {tth}, threads per instruction {tth / tot:.2f}, samples {tsm}")
for wi, ti, sm, f, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{100 * wi / tot:5.1f}% inst {100 * sm / tsm:5.1f}% smp  thr/inst {ti / max(wi, 1):5.1f}  {f}:{ln:<4d} {src[:110]}")
