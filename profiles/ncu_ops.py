"""Executed-instruction histogram by SASS opcode for one kernel of an ncu report (weighted by executions)."""
import collections, csv, subprocess, sys
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None
h = collections.Counter(); st = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == 'Address': hdr = {n: i for i, n in enumerate(r)}; continue
    if hdr is None: continue
    try:
        n = int(r[hdr['Instructions Executed']]); s = int(r[hdr['Warp Stall Sampling (All Samples)']])
    except Exception:
        continue
    ins = r[hdr['Source']].split()
    op = ins[1] if ins[0].startswith('@') else ins[0]
    key = '.'.join(op.split('.')[:2]) if op.startswith(('LD', 'ST', 'RED', 'BAR')) else op.split('.')[0]
    h[key] += n; st[key] += s
tot = sum(h.values()); ts = sum(st.values())
print("total %d warp instructions, %d samples" % (tot, ts))
for k, v in h.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 28):
    print('%-14s %5.1f%% inst %5.1f%% samples' % (k, 100 * v / tot, 100 * st[k] / max(ts, 1)))
