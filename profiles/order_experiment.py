"""Does CTA load imbalance matter?  Same cfg2 P2 ROIs in three orders: as generated, heaviest first (LPT), lightest first."""
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
def timeit(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)
cost = (boxes[:, 3] - boxes[:, 0]) * (boxes[:, 4] - boxes[:, 1])          # ~ footprint plane size
for name, order in (("as generated", np.arange(len(boxes))), ("heaviest first", np.argsort(-cost)), ("lightest first", np.argsort(cost))):
    tb, ti = torch.from_numpy(boxes[order]).to(dev), torch.from_numpy(bidx[order]).to(dev)
    for c in (14, 7):
        g = torch.randn((len(boxes), c, c, c, shape[4]), device=dev)
        out = torch.empty_like(g)
        tf = timeit(lambda: rb.crop_and_resize_3d(image, tb, ti, (c, c, c)))
        tg = timeit(lambda: rb.crop_and_resize_3d_grad_image(g, tb, ti, shape))
        print("%-15s crop %2d  fwd %.4f ms  bwd %.4f ms" % (name, c, tf, tg))
