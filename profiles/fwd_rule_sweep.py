"""Direct vs plane-staged forward across channel counts, pool sizes and ROI counts: data for the auto-selection rule."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import roi3d_b200 as rb, roi3d_synth
dev = torch.device('cuda', 0)
vol = (128, 128, 128)
def timeit(fn, reps=12):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)
for iso in (False, True):
    for C in (32, 64, 128, 256):
        shape = roi3d_synth.level_shape(vol, 2, batch=1, channels=C, isotropic=iso)
        image = torch.randn(shape, device=dev)
        for p in (5, 7, 10):
            for n in (64, 256, 1024):
                boxes = roi3d_synth.rois(n, vol, seed=n + p)
                tb = torch.from_numpy(boxes).to(dev); ti = torch.zeros(n, dtype=torch.int32, device=dev)
                t = {}
                for var in (1, 2):
                    rb.set_option("car_fwd_variant", var)
                    t[var] = timeit(lambda: rb.crop_and_resize_3d(image, tb, ti, (p, p, p)))
                print("%s C %3d pool %2d n %4d: direct %.4f plane %.4f  -> %s" % ("iso" if iso else "aniso", C, p, n, t[1], t[2], "plane" if t[2] < t[1] else "direct"), flush=True)
