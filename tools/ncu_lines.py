"""Aggregate an ncu report's warp-stall samples / executed instructions by CUDA source line.
Usage: python tools/ncu_lines.py report.ncu-rep [top_n] [inst|samples]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
by = 1 if (len(sys.argv) > 3 and sys.argv[3] == "inst") else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur = None; agg = {}; tot = [0, 0]
for r in csv.reader(io.StringIO(out)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) < 8 or not r[0].strip().isdigit(): continue
    try: s = int(r[4]); ins = int(r[7])
    except ValueError: continue
    a = agg.setdefault((cur, int(r[0])), [0, 0, r[1].strip()[:95]]); a[0] += s; a[1] += ins; tot[0] += s; tot[1] += ins
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h = rows[0]
for k in ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "dram__bytes_read.sum", "dram__bytes_write.sum"]:
    if k in h: print(k, rows[2][h.index(k)], rows[1][h.index(k)])
print("total samples", tot[0], "total warp-inst", tot[1])
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][by])[:top]:
    print(f"{k[0]}:{k[1]:4d} smp={v[0]:6d} ({100*v[0]/max(tot[0],1):4.1f}%) inst={v[1]:9d} ({100*v[1]/max(tot[1],1):4.1f}%)  {v[2]}")
