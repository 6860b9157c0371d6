"""Profiling driver for the output-stationary grad-image kernel: cfg2 P2, one crop size, chosen options.
usage: prof_os.py <crop> <debug> <tz> <stage KiB> <ring stages> [iters] [variant]   (run plain first, then under ncu)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import roi3d_b200 as rb, roi3d_synth
c, V, tz, shape_id, cpc = (int(v) for v in sys.argv[1:6])
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 3
variant = int(sys.argv[7]) if len(sys.argv) > 7 else 3
dev = torch.device("cuda", 0)
vol, B = (128, 128, 128), 2
boxes, bidx, _ = roi3d_synth.pyramid_rois(128, B, vol, seed=2002)[2]
shape = roi3d_synth.level_shape(vol, 2, batch=B)
tb, ti = torch.from_numpy(boxes).to(dev), torch.from_numpy(bidx).to(dev)
torch.manual_seed(0)
g = torch.randn((len(boxes), c, c, c, shape[4]), device=dev)
rb.set_option("car_bwd_variant", variant); rb.set_option("car_os_debug", V); rb.set_option("car_os_tile_depth", tz)
rb.set_option("car_os_stage_kib", shape_id); rb.set_option("car_os_ring_stages", cpc)  # r2: stage KiB / ring depth / (V = debug mode)
for _ in range(iters):
    gi = rb.crop_and_resize_3d_grad_image(g, tb, ti, shape)
torch.cuda.synchronize()
print("ok", tuple(gi.shape))
