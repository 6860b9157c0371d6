"""Zero-fill grid size (car_fill_ctas_per_sm) vs the grad-image call time: per-op P2 call and the fused pyramid grad call, cfg2."""
import ctypes, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import roi3d_b200 as rb, roi3d_synth
dev = torch.device('cuda', 0)
lib = rb._lib.load(); vp = ctypes.c_void_p
vol, B, R = (128, 128, 128), 2, 128
boxes, bidx, _ = roi3d_synth.pyramid_rois(R, B, vol, seed=2002)[2]
shape = roi3d_synth.level_shape(vol, 2, batch=B)
tb, ti = torch.from_numpy(boxes).to(dev), torch.from_numpy(bidx).to(dev)
boxes_br = torch.from_numpy(np.stack([roi3d_synth.rois(R, vol, 2002 * 131 + b) for b in range(B)])).to(dev)
shapes = [roi3d_synth.level_shape(vol, lv, batch=B) for lv in roi3d_synth.LEVELS]
gms = [torch.empty(s, device=dev) for s in shapes]
gm_ptrs = (vp * 4)(*[g.data_ptr() for g in gms])
lshapes = (ctypes.c_int * 12)(*[int(d) for s in shapes for d in s[1:4]])
ishape = (ctypes.c_float * 3)(*[float(v) for v in vol])
C = shape[4]
def timeit(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)
stream = lambda: vp(torch.cuda.current_stream().cuda_stream)
for c in (14, 7):
    g = torch.randn((len(boxes), c, c, c, C), device=dev)
    out = torch.empty(shape, device=dev)
    def perop():
        rb._lib.check(lib.roi3d_car3d_grad_image(vp(g.data_ptr()), vp(tb.data_ptr()), vp(ti.data_ptr()), len(boxes), c, c, c,
                                                 shape[0], shape[1], shape[2], shape[3], C, 0, vp(out.data_ptr()), stream()))
    def fused():
        rb._lib.check(lib.roi3d_pyramid_roi_align_grad(vp(g.data_ptr()), gm_ptrs, lshapes, B, C, vp(boxes_br.data_ptr()), R, ishape, c, c, c, stream()))
    for n in (8, 6, 4, 3, 2, 1):
        rb.set_option("car_fill_ctas_per_sm", n)
        print("crop %2d fill CTAs/SM %d: per-op P2 call %.4f ms   fused 4-level call %.4f ms" % (c, n, timeit(perop), timeit(fused)), flush=True)
rb.set_option("car_fill_ctas_per_sm", 0)
