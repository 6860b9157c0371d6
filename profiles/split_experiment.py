"""Scatter grad-image kernel: whole-batch zero-fill + scatter vs image-by-image (option car_bwd_image_split 2 / 1), cfg2 P2."""
import ctypes, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import roi3d_b200 as rb, roi3d_synth
dev = torch.device('cuda', 0)
lib = rb._lib.load(); vp = ctypes.c_void_p
vol = (128, 128, 128)
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)
B, R = 2, 128
boxes, bidx, _ = roi3d_synth.pyramid_rois(R, B, vol, seed=2002)[2]
shape = roi3d_synth.level_shape(vol, 2, batch=B)
tb, ti = torch.from_numpy(boxes).to(dev), torch.from_numpy(bidx).to(dev)
out = torch.empty(shape, device=dev)
rb.set_option("car_bwd_variant", 2)
for c in (7, 14):
    g = torch.randn((len(boxes), c, c, c, shape[4]), device=dev)
    def call():
        Bq, H, W, D, C = shape
        rb._lib.check(lib.roi3d_car3d_grad_image(vp(g.data_ptr()), vp(tb.data_ptr()), vp(ti.data_ptr()), len(boxes), c, c, c,
                                                 Bq, H, W, D, C, 0, vp(out.data_ptr()), vp(torch.cuda.current_stream().cuda_stream)))
    res = {}
    for mode in (2, 1):
        rb.set_option("car_bwd_image_split", mode)
        t = timeit(call); res[mode] = out.clone()
        print("crop %2d image_split=%s  %.4f ms" % (c, "off" if mode == 2 else "on", t))
    print("  max diff", float((res[1] - res[2]).abs().max()))
rb.set_option("car_bwd_image_split", 0)
