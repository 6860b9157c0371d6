"""NMS3D profiling driver.  usage: prof_nms.py <n> <max_out> <dense 0|1> [iters]
cfg1 = 6000 1000 0, cfg3 = 20000 2000 0; dense = clusters of 64 near-duplicates (the regime a trained RPN produces:
the scan must go 4-5x deeper before max_out boxes are kept).  Run plain first, then under ncu."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import roi3d_b200 as rb, roi3d_synth
n, mo, dense = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda", 0)
vol = (256, 256, 256) if n > 6000 else (128, 128, 128)
kw = dict(cluster=64, jitter=0.05) if dense else {}
b, s = roi3d_synth.nms_boxes(n, vol, **kw)
db, ds = torch.from_numpy(b).to(dev), torch.from_numpy(s).to(dev)
for _ in range(iters):
    keep = rb.non_max_suppression_3d(db, ds, mo, 0.7)
torch.cuda.synchronize()
print(n, mo, dense, len(keep))
