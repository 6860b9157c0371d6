"""Output-stationary grad-image kernel (car_bwd_variant 3) vs the RED scatter kernel (2) at the BASELINE shapes:
cfg2 P2 (256 ROIs, 7^3 / 14^3), cfg4 P2 (1000 ROIs, 14^3), over tile depth and channel groups per thread.
The C-ABI call is timed with CUDA events into a preallocated output (no allocator inside the timed call)."""
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
cases = [("cfg2", 2, 128, (7, 14))] + ([("cfg4", 1, 1000, (14,))] if "--cfg4" in sys.argv else [])
for name, B, R, crops in cases:
    boxes, bidx, _ = roi3d_synth.pyramid_rois(R, B, vol, seed=2002)[2]
    shape = roi3d_synth.level_shape(vol, 2, batch=B)
    tb, ti = torch.from_numpy(boxes).to(dev), torch.from_numpy(bidx).to(dev)
    out = torch.empty(shape, device=dev)
    for c in crops:
        g = torch.randn((len(boxes), c, c, c, shape[4]), device=dev)
        bb = roi3d_synth.car_algorithmic_bytes(boxes, shape, (c, c, c), True)
        comp = (int(np.prod(shape)) + g.numel()) * 4
        def call():
            Bq, H, W, D, C = shape
            rb._lib.check(lib.roi3d_car3d_grad_image(vp(g.data_ptr()), vp(tb.data_ptr()), vp(ti.data_ptr()), len(boxes), c, c, c,
                                                     Bq, H, W, D, C, 0, vp(out.data_ptr()), vp(torch.cuda.current_stream().cuda_stream)))
        rb.set_option("car_bwd_variant", 2); rb.set_option("car_lanes_v", 0)
        t2 = timeit(call); ref = out.clone()
        print("%s crop %2d scatter          %.4f ms  alg %5.0f GB/s  compulsory %5.0f GB/s" % (name, c, t2, bb / t2 / 1e6, comp / t2 / 1e6))
        rb.set_option("car_bwd_variant", 3)
        best = None
        for tz in (8, 16, 32):
            rb.set_option("car_os_tile_depth", tz)
            try:
                t3 = timeit(call)
            except Exception as e:
                print("  tz%2d failed: %s" % (tz, e)); continue
            err = float((out - ref).abs().max() / ref.abs().max())
            print("%s crop %2d os tz%2d  %.4f ms  alg %5.0f GB/s  compulsory %5.0f GB/s  maxerr %.1e" %
                  (name, c, tz, t3, bb / t3 / 1e6, comp / t3 / 1e6, err))
            if best is None or t3 < best[0]: best = (t3, tz)
        print("%s crop %2d BEST os %.4f ms tz%d (scatter %.4f)" % ((name, c) + best + (t2,)))
        for k in ("car_os_tile_depth",): rb.set_option(k, 0)
        del g
