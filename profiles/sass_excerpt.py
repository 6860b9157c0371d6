"""SASS evidence per hot kernel: counts of the memory / TMA / sync mnemonics in the built library.
Usage: python profiles/sass_excerpt.py > profiles/r2_sass_excerpt.txt   (runs cuobjdump -sass on lib/libroi3d_b200.so)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "3d-mask-r-cnn_b200", "lib", "libroi3d_b200.so")
PAT = re.compile(r"\b(LDG\.E[.\w]*|STG\.E[.\w]*|REDG?\.E[.\w]*|RED\.E[.\w]*|ATOMG?[.\w]*|LDS[.\w]*|STS[.\w]*|LDGSTS[.\w]*|UBLKCP[.\w]*|"
                 r"UTMALDG[.\w]*|UTMASTG[.\w]*|UTMAPF[.\w]*|SYNCS[.\w]*|BAR\.[.\w]*|CCTL[.\w]*|MATCH[.\w]*|VOTE[.\w]*|REDUX[.\w]*|"
                 r"SHFL[.\w]*|ACQBULK|FFMA2?|FADD2?|FMUL2?|MUFU[.\w]*|UTCBAR[.\w]*|ERRBAR|MEMBAR[.\w]*|NANOSLEEP)\b")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = {}
names = re.findall(r"Function : (\S+)", sass)
if names:
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, out))
cur, counts, total = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        total[cur] = 0
        continue
    if cur is None or "/*" not in line:
        continue
    ins = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][\w.]*)", line)
    if not ins:
        continue
    total[cur] += 1
    op = ins.group(1)
    if PAT.match(op):
        counts[cur][op] += 1
want = sys.argv[1:] or ["car3d_fwd_plane_kernel", "car3d_fwd_direct", "car3d_grad_image_plane_kernel", "car3d_grad_image_os",
                        "car3d_fwd_plane_tma", "car3d_fwd_box", "zero_fill", "nms_mask_kernel", "nms_scan_kernel", "nms_rank_sort",
                        "nms_keys", "topk", "mask_targets"]
print("# cuobjdump -sass 3d-mask-r-cnn_b200/lib/libroi3d_b200.so (sm_100a), instruction-site counts per kernel")
print("# vector global accesses: LDG.E.128 / STG.E.128 / REDG.E.ADD.F32x4-style vector REDs; TMA: UBLKCP (cp.async.bulk),")
print("# UTMALDG (cp.async.bulk.tensor); mbarrier: SYNCS.*; warp collectives: MATCH / VOTE / REDUX / SHFL")
for fn, c in counts.items():
    name = demangle.get(fn, fn)
    if not any(w in name for w in want):
        continue
    print("\n== %s\n   %d SASS instructions" % (name[:150], total[fn]))
    groups = collections.OrderedDict()
    for op, n in sorted(c.items()):
        groups.setdefault(op.split(".")[0], []).append("%s x%d" % (op, n))
    for k, v in groups.items():
        print("   " + ", ".join(v))
