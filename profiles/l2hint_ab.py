"""A/B: L2 evict-first hint on the TMA-staged grads slices of the grad-image kernel (car_experiment bit 32 = no hint), cfg2 P2 + cfg4."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import roi3d_b200 as rb, roi3d_synth
dev = torch.device('cuda', 0)
vol = (128, 128, 128)
def timeit(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)
for name, R, B in (("cfg2", 128, 2), ("cfg4", 1000, 1)):
    boxes, bidx, _ = roi3d_synth.pyramid_rois(R, B, vol, seed=2002)[2]
    shape = roi3d_synth.level_shape(vol, 2, batch=B)
    tb, ti = torch.from_numpy(boxes).to(dev), torch.from_numpy(bidx).to(dev)
    for c in ((14, 7) if name == "cfg2" else (14,)):
        g = torch.randn((len(boxes), c, c, c, shape[4]), device=dev)
        for rep in range(2):
            for ex in (0, 32):
                rb.set_option("car_experiment", ex)
                t = timeit(lambda: rb.crop_and_resize_3d_grad_image(g, tb, ti, shape))
                print("%s crop %2d  L2 evict-first hint %-3s: %.4f ms" % (name, c, "off" if ex else "on", t), flush=True)
        del g
rb.set_option("car_experiment", 0)
