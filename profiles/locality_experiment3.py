"""Ascending vs descending (image, y) order for the grad-image call: the zero-fill sweeps addresses upwards, so its last
~60 MB (high image index, high y) are still dirty in L2 when the scatter starts."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import roi3d_b200 as rb, roi3d_synth
dev = torch.device('cuda', 0)
vol = (128, 128, 128)
rb.set_option("car_experiment", 16)
def timeit(fn, reps=40):
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
    yc = (boxes[:, 0] + boxes[:, 3]) / 2
    asc = np.lexsort((yc, bidx))
    orders = {"(image, y) ascending": asc, "(image, y) descending": asc[::-1].copy()}
    for c in (14, 7):
        g0 = torch.randn((len(boxes), c, c, c, shape[4]), device=dev)
        for rep in range(2):
            for oname, order in orders.items():
                o = torch.from_numpy(order).to(dev)
                tb, ti = torch.from_numpy(boxes[order]).to(dev), torch.from_numpy(bidx[order]).to(dev)
                g = g0[o].contiguous()
                print("%s crop %2d  %-24s bwd %.4f ms" % (name, c, oname, timeit(lambda: rb.crop_and_resize_3d_grad_image(g, tb, ti, shape))), flush=True)
                del g
        del g0
rb.set_option("car_experiment", 0)
