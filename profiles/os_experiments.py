"""Where does the output-stationary kernel's time go?  cfg2 P2 14^3: ring depth x stage size x tile depth, and the two
half-kernels (debug 1: consumers skip the arithmetic = pure staging; debug 2: producer skips the copies = pure compute)."""
import ctypes, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import roi3d_b200 as rb, roi3d_synth
dev = torch.device('cuda', 0)
lib = rb._lib.load(); vp = ctypes.c_void_p
vol = (128, 128, 128)
def timeit(fn, reps=10):
    for _ in range(2): fn()
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
c = int(sys.argv[1]) if len(sys.argv) > 1 else 14
g = torch.randn((len(boxes), c, c, c, shape[4]), device=dev)
def call():
    Bq, H, W, D, C = shape
    rb._lib.check(lib.roi3d_car3d_grad_image(vp(g.data_ptr()), vp(tb.data_ptr()), vp(ti.data_ptr()), len(boxes), c, c, c,
                                             Bq, H, W, D, C, 0, vp(out.data_ptr()), vp(torch.cuda.current_stream().cuda_stream)))
rb.set_option("car_bwd_variant", 3)
for tz, ns, sbk in ((16, 2, 8), (16, 4, 8), (16, 2, 16), (32, 2, 8), (32, 4, 8), (32, 2, 16), (8, 2, 8), (16, 4, 4), (64, 2, 8)):
    for dbg in (0, 1, 2):
        rb.set_option("car_os_tile_depth", tz); rb.set_option("car_os_ring_stages", ns); rb.set_option("car_os_stage_kib", sbk)
        rb.set_option("car_os_debug", dbg)
        try:
            t = timeit(call)
            print("tz%2d ns%2d sb%2dK debug%d  %.4f ms" % (tz, ns, sbk, dbg, t))
        except Exception as e:
            print("tz%2d ns%2d sb%2dK debug%d failed %s" % (tz, ns, sbk, dbg, e))
