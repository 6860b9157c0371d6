"""Does the processing ORDER of the ROIs matter for the grad-image scatter?  The REDs of overlapping footprints hit L2 only if
they are close in time.  Same cfg2 P2 ROIs, same kernel, ROIs passed in different orders (host-side permutation)."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import roi3d_b200 as rb, roi3d_synth
dev = torch.device('cuda', 0)
vol, B = (128, 128, 128), 2
boxes, bidx, _ = roi3d_synth.pyramid_rois(128, B, vol, seed=2002)[2]
shape = roi3d_synth.level_shape(vol, 2, batch=B)
def timeit(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)
zc = (boxes[:, 2] + boxes[:, 5]) / 2; yc = (boxes[:, 0] + boxes[:, 3]) / 2; xc = (boxes[:, 1] + boxes[:, 4]) / 2
def morton(a, b, c, bits=5):
    q = lambda v: np.clip((v * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    a, b, c = q(a), q(b), q(c); out = np.zeros_like(a)
    for i in range(bits):
        out |= ((a >> i) & 1) << (3 * i + 2) | ((b >> i) & 1) << (3 * i + 1) | ((c >> i) & 1) << (3 * i)
    return out
rng = np.random.default_rng(0)
orders = {
    "as generated (image-major, random inside)": np.arange(len(boxes)),
    "image, then z centre": np.lexsort((zc, bidx)),
    "image, then y centre": np.lexsort((yc, bidx)),
    "image, then Morton(y,x,z) of the centre": np.lexsort((morton(yc, xc, zc), bidx)),
    "images interleaved, random": rng.permutation(len(boxes)),
}
for c in (14, 7):
    g0 = torch.randn((len(boxes), c, c, c, shape[4]), device=dev)
    for name, order in orders.items():
        o = torch.from_numpy(order).to(dev)
        tb, ti = torch.from_numpy(boxes[order]).to(dev), torch.from_numpy(bidx[order]).to(dev)
        g = g0[o].contiguous()
        t = timeit(lambda: rb.crop_and_resize_3d_grad_image(g, tb, ti, shape))
        img = torch.randn(shape, device=dev)
        tf = timeit(lambda: rb.crop_and_resize_3d(img, tb, ti, (c, c, c)))
        print("crop %2d  %-45s bwd %.4f ms   fwd %.4f ms" % (c, name, t, tf), flush=True)
        del img
