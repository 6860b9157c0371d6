"""Profiling driver for the row-walk kernels: cfg2 P2, one pool size, chosen knobs.  python profiles/prof_sep.py fwd|bwd crop V rows ring target"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import roi3d_b200 as rb, roi3d_synth
what, c, V, rc, ns, tgt = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
dev = torch.device('cuda', 0)
vol, B = (128, 128, 128), 2
boxes, bidx, _ = roi3d_synth.pyramid_rois(128, B, vol, seed=2002)[2]
shape = roi3d_synth.level_shape(vol, 2, batch=B)
torch.manual_seed(0)
image = torch.randn(shape, device=dev)
tb, ti = torch.from_numpy(boxes).to(dev), torch.from_numpy(bidx).to(dev)
for k, v in dict(car_fwd_variant=4, car_bwd_variant=4, car_lanes_v=V, car_sep_rows=rc, car_sep_ring=ns, car_ctas_per_sm_target=tgt).items():
    rb.set_option(k, v)
g = torch.randn((len(boxes), c, c, c, shape[4]), device=dev)
for _ in range(3):
    if what == "fwd":
        out = rb.crop_and_resize_3d(image, tb, ti, (c, c, c))
    else:
        out = rb.crop_and_resize_3d_grad_image(g, tb, ti, shape)
torch.cuda.synchronize()
print("ok", tuple(out.shape))
