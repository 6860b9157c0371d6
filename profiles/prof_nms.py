"""NMS3D profiling driver: cfg1 (6000 -> 1000 @0.7) and cfg3 (20000 -> 2000 @0.7)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import roi3d_b200 as rb, roi3d_synth
dev = torch.device("cuda", 0)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for n, mo, vol in ((6000, 1000, (128, 128, 128)), (20000, 2000, (256, 256, 256))):
    b, s = roi3d_synth.nms_boxes(n, vol)
    db, ds = torch.from_numpy(b).to(dev), torch.from_numpy(s).to(dev)
    for _ in range(iters):
        keep = rb.non_max_suppression_3d(db, ds, mo, 0.7)
    torch.cuda.synchronize()
    print(n, len(keep))
