"""Per-source-line instruction counts and stall samples of one kernel in an ncu report (needs -lineinfo + --import-source on).
python profiles/ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
files = {}
cur = None
hdr = None
agg = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = {n: i for i, n in enumerate(r)}; continue
    if hdr is None or r[2] != '-': continue          # '-' address = the per-line aggregate row
    try:
        n = int(r[hdr['Instructions Executed']]); s = int(r[hdr['Warp Stall Sampling (All Samples)']])
    except Exception:
        continue
    agg[(cur, int(r[0]))] = (n, s, r[1].strip()[:105])
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("total warp instructions %d, stall samples %d" % (ti, ts))
for (f, l), (n, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% samp %5.1f%% inst  %s:%d  %s" % (100.0 * s / max(ts, 1), 100.0 * n / max(ti, 1), f, l, src))
