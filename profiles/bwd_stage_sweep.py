"""Backward plane kernel: staged grads slice capacity (car_bwd_stage_kib) x V x CTAs/SM target, cfg2 P2 and a 28^3 case."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import roi3d_b200 as rb, roi3d_synth
dev = torch.device('cuda', 0)
vol, B = (128, 128, 128), 2
boxes, bidx, _ = roi3d_synth.pyramid_rois(128, B, vol, seed=2002)[2]
shape = roi3d_synth.level_shape(vol, 2, batch=B)
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
for c, n in ((14, 256), (7, 256), (28, 64)):
    g = torch.randn((n, c, c, c, shape[4]), device=dev)
    bb, ii = tb[:n], ti[:n]
    rb.set_option("car_lanes_v", 0); rb.set_option("car_bwd_stage_kib", 0); rb.set_option("car_ctas_per_sm_target", 0)
    ref = rb.crop_and_resize_3d_grad_image(g, bb, ii, shape)
    print("crop %2d n %3d default: %.4f ms" % (c, n, timeit(lambda: rb.crop_and_resize_3d_grad_image(g, bb, ii, shape))), flush=True)
    for V in (2, 1):
        for kib in (32, 50, 64, 72, 100, 110, 200):
            for tgt in (16, 24):
                rb.set_option("car_lanes_v", V); rb.set_option("car_bwd_stage_kib", kib); rb.set_option("car_ctas_per_sm_target", tgt)
                try:
                    out = rb.crop_and_resize_3d_grad_image(g, bb, ii, shape)
                    err = float((out - ref).abs().max() / ref.abs().max())
                    t = timeit(lambda: rb.crop_and_resize_3d_grad_image(g, bb, ii, shape))
                    print("crop %2d V=%d stage %3d KiB target %2d: %.4f ms  err %.1e" % (c, V, kib, tgt, t, err), flush=True)
                except Exception as e:
                    print("crop %2d V=%d stage %3d KiB target %2d: %s" % (c, V, kib, tgt, str(e)[:60]), flush=True)
rb.set_option("car_lanes_v", 0); rb.set_option("car_bwd_stage_kib", 0); rb.set_option("car_ctas_per_sm_target", 0)
