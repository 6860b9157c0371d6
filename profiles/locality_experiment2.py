"""More ROI orders (host-side permutations, library ordering switched off): bands along y with a snake along x inside,
2-D Morton, at cfg2 (128 ROIs / image) and cfg4 (1000 ROIs on one image)."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import roi3d_b200 as rb, roi3d_synth
dev = torch.device('cuda', 0)
vol = (128, 128, 128)
rb.set_option("car_experiment", 16)                     # the library processes the ROIs in the order given
def timeit(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)
def band_snake(yc, xc, nb, nxb):
    yb = np.clip((yc * nb).astype(int), 0, nb - 1)
    xb = np.clip((xc * nxb).astype(int), 0, nxb - 1)
    xs = np.where(yb % 2 == 0, xb, nxb - 1 - xb)
    return yb * nxb + xs
def morton2(a, b, bits=4):
    q = lambda v: np.clip((v * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    a, b = q(a), q(b); out = np.zeros_like(a)
    for i in range(bits):
        out |= ((a >> i) & 1) << (2 * i + 1) | ((b >> i) & 1) << (2 * i)
    return out
for name, R, B in (("cfg2", 128, 2), ("cfg4", 1000, 1)):
    boxes, bidx, _ = roi3d_synth.pyramid_rois(R, B, vol, seed=2002)[2]
    shape = roi3d_synth.level_shape(vol, 2, batch=B)
    yc = (boxes[:, 0] + boxes[:, 3]) / 2; xc = (boxes[:, 1] + boxes[:, 4]) / 2; zc = (boxes[:, 2] + boxes[:, 5]) / 2
    orders = {
        "as given": np.arange(len(boxes)),
        "image, y (64 buckets)": np.lexsort((np.arange(len(boxes)), (yc * 64).astype(int), bidx)),
        "image, y exact": np.lexsort((yc, bidx)),
        "image, 4 y-bands x snake(16)": np.lexsort((np.arange(len(boxes)), band_snake(yc, xc, 4, 16), bidx)),
        "image, 8 y-bands x snake(8)": np.lexsort((np.arange(len(boxes)), band_snake(yc, xc, 8, 8), bidx)),
        "image, 16 y-bands x snake(4)": np.lexsort((np.arange(len(boxes)), band_snake(yc, xc, 16, 4), bidx)),
        "image, Morton(y,x) 4 bits": np.lexsort((np.arange(len(boxes)), morton2(yc, xc), bidx)),
        "image, 8 y-bands x snake z(8)": np.lexsort((np.arange(len(boxes)), band_snake(yc, zc, 8, 8), bidx)),
    }
    c = 14
    g0 = torch.randn((len(boxes), c, c, c, shape[4]), device=dev)
    img = torch.randn(shape, device=dev)
    for oname, order in orders.items():
        o = torch.from_numpy(order).to(dev)
        tb, ti = torch.from_numpy(boxes[order]).to(dev), torch.from_numpy(bidx[order]).to(dev)
        g = g0[o].contiguous()
        tbw = timeit(lambda: rb.crop_and_resize_3d_grad_image(g, tb, ti, shape))
        tfw = timeit(lambda: rb.crop_and_resize_3d(img, tb, ti, (c, c, c)))
        print("%s crop 14  %-32s bwd %.4f ms   fwd %.4f ms" % (name, oname, tbw, tfw), flush=True)
        del g
    del g0, img
rb.set_option("car_experiment", 0)
