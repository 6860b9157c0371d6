"""Profiling driver: forward variant 5 (TMA gather4) at cfg2 P2, one pool size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import roi3d_b200 as rb, roi3d_synth
c, tgt = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device('cuda', 0)
vol, B = (128, 128, 128), 2
boxes, bidx, _ = roi3d_synth.pyramid_rois(128, B, vol, seed=2002)[2]
shape = roi3d_synth.level_shape(vol, 2, batch=B)
torch.manual_seed(0)
image = torch.randn(shape, device=dev)
tb, ti = torch.from_numpy(boxes).to(dev), torch.from_numpy(bidx).to(dev)
rb.set_option("car_fwd_variant", 5); rb.set_option("car_ctas_per_sm_target", tgt)
for _ in range(3):
    out = rb.crop_and_resize_3d(image, tb, ti, (c, c, c))
torch.cuda.synchronize()
print("ok", tuple(out.shape))
