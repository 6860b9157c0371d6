import os, sys, statistics
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
for c in (14, 7):
    g = torch.randn((len(boxes), c, c, c, shape[4]), device=dev)
    for ex in (0, 2):
        for kib in (0, 48, 36):
            for tgt in (16, 24, 32):
                rb.set_option("car_experiment", ex); rb.set_option("car_bwd_stage_kib", kib); rb.set_option("car_ctas_per_sm_target", tgt)
                print("crop %2d exp %d stage %2d KiB target %2d: bwd %.4f ms" % (c, ex, kib, tgt, timeit(lambda: rb.crop_and_resize_3d_grad_image(g, tb, ti, shape))), flush=True)
