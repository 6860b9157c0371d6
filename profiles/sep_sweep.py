"""Row-walk separable kernels (variant 4) vs the plane-staged kernels (variant 2) at cfg2 P2: bit-equality of the forward,
tolerance of the backward, and the timing over the tuning knobs (rows per tile, ring slots, CTAs/SM target, V)."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import roi3d_b200 as rb, roi3d_synth
dev = torch.device('cuda', 0)
what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
vol, B = (128, 128, 128), 2
routed = roi3d_synth.pyramid_rois(128, B, vol, seed=2002)
boxes, bidx, _ = routed[2]
shape = roi3d_synth.level_shape(vol, 2, batch=B)
torch.manual_seed(0)
image = torch.randn(shape, device=dev)
tb, ti = torch.from_numpy(boxes).to(dev), torch.from_numpy(bidx).to(dev)
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)
def opts(**kw):
    for k, v in kw.items(): rb.set_option(k, v)
crops = (14, 7) if len(sys.argv) < 3 else tuple(int(v) for v in sys.argv[2].split(","))
for c in crops:
    if what == "fwd":
        opts(car_fwd_variant=2, car_lanes_v=0, car_ctas_per_sm_target=0)
        ref = rb.crop_and_resize_3d(image, tb, ti, (c, c, c))
        t2 = timeit(lambda: rb.crop_and_resize_3d(image, tb, ti, (c, c, c)))
        print("crop %2d plane-staged (variant 2): %.4f ms" % (c, t2), flush=True)
        opts(car_fwd_variant=4)
        for V in (1, 2):
            for rc, ns in (((8, 9), (12, 13), (16, 17), (8, 17)) if V == 1 else ((4, 5), (5, 6), (6, 7), (8, 9))):
                for tgt in (16, 32):
                    opts(car_lanes_v=V, car_sep_rows=rc, car_sep_ring=ns, car_ctas_per_sm_target=tgt)
                    try:
                        out = rb.crop_and_resize_3d(image, tb, ti, (c, c, c))
                        ok = torch.equal(out, ref)
                        t4 = timeit(lambda: rb.crop_and_resize_3d(image, tb, ti, (c, c, c)))
                        print("crop %2d row-walk V=%d rows/tile %2d ring %2d target %2d: %.4f ms  bit-equal %s" % (c, V, rc, ns, tgt, t4, ok), flush=True)
                    except Exception as e:
                        print("crop %2d row-walk V=%d rows/tile %2d ring %2d target %2d: %s" % (c, V, rc, ns, tgt, e), flush=True)
    else:
        g = torch.randn((len(boxes), c, c, c, shape[4]), device=dev)
        opts(car_bwd_variant=2, car_lanes_v=0, car_ctas_per_sm_target=0)
        ref = rb.crop_and_resize_3d_grad_image(g, tb, ti, shape)
        t2 = timeit(lambda: rb.crop_and_resize_3d_grad_image(g, tb, ti, shape))
        print("crop %2d plane-staged scatter (variant 2): %.4f ms" % (c, t2), flush=True)
        opts(car_bwd_variant=4)
        for V in (1, 2):
            for rc, ns in ((4, 9), (8, 9), (8, 17), (14, 15), (14, 29), (2, 5)):
                for tgt in (8, 16, 32):
                    opts(car_lanes_v=V, car_sep_rows=rc, car_sep_ring=ns, car_ctas_per_sm_target=tgt)
                    try:
                        out = rb.crop_and_resize_3d_grad_image(g, tb, ti, shape)
                        err = float((out - ref).abs().max() / ref.abs().max())
                        t4 = timeit(lambda: rb.crop_and_resize_3d_grad_image(g, tb, ti, shape))
                        print("crop %2d row-walk V=%d rows/tile %2d ring %2d target %2d: %.4f ms  max err / max %.2e" % (c, V, rc, ns, tgt, t4, err), flush=True)
                    except Exception as e:
                        print("crop %2d row-walk V=%d rows/tile %2d ring %2d target %2d: %s" % (c, V, rc, ns, tgt, e), flush=True)
opts(car_fwd_variant=0, car_bwd_variant=0, car_lanes_v=0, car_sep_rows=0, car_sep_ring=0, car_ctas_per_sm_target=0)
