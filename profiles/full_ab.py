"""A/B: FULL-channel template and equal y-tiles in the plane kernels (car_experiment bits: 1 = forward without FULL,
2 = backward without FULL, 4 = backward with greedy y-tiles), per-op P2 calls and the fused pyramid calls, cfg2."""
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
torch.manual_seed(0)
fms = [torch.randn(s, device=dev) for s in shapes]
gms = [torch.empty(s, device=dev) for s in shapes]
fm_ptrs = (vp * 4)(*[t.data_ptr() for t in fms]); gm_ptrs = (vp * 4)(*[t.data_ptr() for t in gms])
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
    crops = torch.empty_like(g)
    def f_op(): rb._lib.check(lib.roi3d_car3d_fwd(vp(fms[0].data_ptr()), shape[0], shape[1], shape[2], shape[3], C, vp(tb.data_ptr()), vp(ti.data_ptr()), len(boxes), c, c, c, 0, 0.0, vp(crops.data_ptr()), stream()))
    def b_op(): rb._lib.check(lib.roi3d_car3d_grad_image(vp(g.data_ptr()), vp(tb.data_ptr()), vp(ti.data_ptr()), len(boxes), c, c, c, shape[0], shape[1], shape[2], shape[3], C, 0, vp(gms[0].data_ptr()), stream()))
    def f_fu(): rb._lib.check(lib.roi3d_pyramid_roi_align_fwd(fm_ptrs, lshapes, B, C, vp(boxes_br.data_ptr()), R, ishape, c, c, c, vp(crops.data_ptr()), stream()))
    def b_fu(): rb._lib.check(lib.roi3d_pyramid_roi_align_grad(vp(g.data_ptr()), gm_ptrs, lshapes, B, C, vp(boxes_br.data_ptr()), R, ishape, c, c, c, stream()))
    for ex in (0, 1, 2, 4, 6, 7):
        rb.set_option("car_experiment", ex)
        print("crop %2d experiment %d: fwd per-op %.4f fused %.4f | bwd per-op %.4f fused %.4f ms" % (c, ex, timeit(f_op), timeit(f_fu), timeit(b_op), timeit(b_fu)), flush=True)
rb.set_option("car_experiment", 0)
