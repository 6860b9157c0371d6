"""Forward fed by TMA gather4 row copies (car_fwd_variant 5) vs the production plane kernel (2) and the UBLKCP-fed one (3), cfg2 P2."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import roi3d_b200 as rb, roi3d_synth
dev = torch.device('cuda', 0)
vol, B = (128, 128, 128), 2
boxes, bidx, _ = roi3d_synth.pyramid_rois(128, B, vol, seed=2002)[2]
shape = roi3d_synth.level_shape(vol, 2, batch=B)
torch.manual_seed(0)
image = torch.randn(shape, device=dev)
tb, ti = torch.from_numpy(boxes).to(dev), torch.from_numpy(bidx).to(dev)
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)
for c in (14, 7, 28):
    n = len(boxes) if c < 28 else 64
    bb, ii = tb[:n], ti[:n]
    rb.set_option("car_fwd_variant", 2)
    ref = rb.crop_and_resize_3d(image, bb, ii, (c, c, c))
    for v, name in ((2, "plane-staged (LDG)"), (3, "TMA bulk copies (UBLKCP, 256 B rows)"), (5, "TMA gather4 (UTMALDG.2D.GATHER4)")):
        for tgt in ((16,) if v == 2 else (16, 24, 32)):
            rb.set_option("car_fwd_variant", v); rb.set_option("car_ctas_per_sm_target", tgt)
            out = rb.crop_and_resize_3d(image, bb, ii, (c, c, c))
            t = timeit(lambda: rb.crop_and_resize_3d(image, bb, ii, (c, c, c)))
            print("crop %2d n %3d variant %d %-40s target %2d: %.4f ms  bit-equal %s" % (c, n, v, name, tgt, t, torch.equal(out, ref)), flush=True)
rb.set_option("car_fwd_variant", 0); rb.set_option("car_ctas_per_sm_target", 0)
