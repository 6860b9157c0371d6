"""Grid sizing of the plane kernels: CTAs per SM targeted by the depth-sample split (car_ctas_per_sm_target)."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import roi3d_b200 as rb, roi3d_synth
dev = torch.device('cuda', 0)
vol, B = (128, 128, 128), 2
routed = roi3d_synth.pyramid_rois(128, B, vol, seed=2002)
boxes, bidx, _ = routed[2]
shape = roi3d_synth.level_shape(vol, 2, batch=B)
torch.manual_seed(0)
image = torch.randn(shape, device=dev)
tb, ti = torch.from_numpy(boxes).to(dev), torch.from_numpy(bidx).to(dev)
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)
rb.set_option("car_fwd_variant", 2); rb.set_option("car_bwd_variant", 2)
for c in (7, 14):
    g = torch.randn((len(boxes), c, c, c, shape[4]), device=dev)
    for tgt in (6, 8, 12, 16, 20, 24, 32, 48):
        rb.set_option("car_ctas_per_sm_target", tgt)
        tf = timeit(lambda: rb.crop_and_resize_3d(image, tb, ti, (c, c, c)))
        tw = timeit(lambda: rb.crop_and_resize_3d_grad_image(g, tb, ti, shape))
        print("crop %2d target %2d CTAs/SM: fwd %.4f ms  bwd %.4f ms" % (c, tgt, tf, tw), flush=True)
