"""A/B: forward plane kernel with 512-thread CTAs (car_experiment bit 128) vs 256, cfg2 P2 + cfg4 + a 28^3 case; bit-equality checked."""
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
for name, R, B, crops in (("cfg2", 128, 2, (14, 7, 28)), ("cfg4", 1000, 1, (14,))):
    boxes, bidx, _ = roi3d_synth.pyramid_rois(R, B, vol, seed=2002)[2]
    shape = roi3d_synth.level_shape(vol, 2, batch=B)
    torch.manual_seed(0)
    image = torch.randn(shape, device=dev)
    for c in crops:
        n = len(boxes) if c < 28 else 64
        tb, ti = torch.from_numpy(boxes[:n]).to(dev), torch.from_numpy(bidx[:n]).to(dev)
        rb.set_option("car_experiment", 0)
        ref = rb.crop_and_resize_3d(image, tb, ti, (c, c, c))
        for rep in range(2):
            for ex in (0, 128):
                for tgt in ((0,) if ex == 0 else (0, 8, 12)):
                    rb.set_option("car_experiment", ex); rb.set_option("car_ctas_per_sm_target", tgt)
                    out = rb.crop_and_resize_3d(image, tb, ti, (c, c, c))
                    t = timeit(lambda: rb.crop_and_resize_3d(image, tb, ti, (c, c, c)))
                    print("%s crop %2d n %4d  %s threads, target %2d: %.4f ms  bit-equal %s" % (name, c, n, "512" if ex else "256", tgt, t, torch.equal(out, ref)), flush=True)
    del image
rb.set_option("car_experiment", 0); rb.set_option("car_ctas_per_sm_target", 0)
