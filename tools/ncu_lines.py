"""Aggregate an ncu report's warp-stall samples by CUDA source line.
Usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur = None; agg = {}; tot = 0
for r in csv.reader(io.StringIO(out)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) < 8 or not r[0].strip().isdigit(): continue
    try: s = int(r[4]); ins = int(r[7])
    except ValueError: continue
    a = agg.setdefault((cur, int(r[0])), [0, 0, r[1].strip()[:100]]); a[0] += s; a[1] += ins; tot += s
print("total samples", tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:4d} {v[0]:6d} ({100*v[0]/max(tot,1):4.1f}%) inst={v[1]:9d}  {v[2]}")
