"""Backward plane kernel with its grads slices staged by TMA tensor tile copies (car_bwd_variant 4) vs the production kernel (2), cfg2 P2 + a 28^3 case."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import roi3d_b200 as rb, roi3d_synth
dev = torch.device('cuda', 0)
vol, B = (128, 128, 128), 2
boxes, bidx, _ = roi3d_synth.pyramid_rois(128, B, vol, seed=2002)[2]
shape = roi3d_synth.level_shape(vol, 2, batch=B)
tb, ti = torch.from_numpy(boxes).to(dev), torch.from_numpy(bidx).to(dev)
def timeit(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)
for c, n in ((14, 256), (7, 256), (28, 64)):
    g = torch.randn((n, c, c, c, shape[4]), device=dev)
    bb, ii = tb[:n], ti[:n]
    rb.set_option("car_bwd_variant", 2); rb.set_option("car_experiment", 0); rb.set_option("car_bwd_stage_kib", 0)
    ref = rb.crop_and_resize_3d_grad_image(g, bb, ii, shape)
    print("crop %2d n %3d production (LDG staging): %.4f ms" % (c, n, timeit(lambda: rb.crop_and_resize_3d_grad_image(g, bb, ii, shape))), flush=True)
    rb.set_option("car_bwd_variant", 4)
    for ex in (0, 2):                       # FULL instantiation off / on for the plain kernel (car_experiment bit 2 flips it)
        for kib in (0, 36, 100):
            for tgt in (16, 24):
                rb.set_option("car_experiment", ex); rb.set_option("car_bwd_stage_kib", kib); rb.set_option("car_ctas_per_sm_target", tgt)
                out = rb.crop_and_resize_3d_grad_image(g, bb, ii, shape)
                err = float((out - ref).abs().max() / ref.abs().max())
                t = timeit(lambda: rb.crop_and_resize_3d_grad_image(g, bb, ii, shape))
                print("crop %2d n %3d TMA staging  full=%d stage %3d KiB target %2d: %.4f ms  err %.1e" % (c, n, ex == 2, kib, tgt, t, err), flush=True)
    rb.set_option("car_ctas_per_sm_target", 0)
rb.set_option("car_bwd_variant", 0); rb.set_option("car_experiment", 0); rb.set_option("car_bwd_stage_kib", 0)
