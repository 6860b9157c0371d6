#!/usr/bin/env python
"""BASELINE.json configs[4] -- microbench sweep on one B200: CropAndResize3D fwd / grad-image over
ROIs 64..8192 x pool {7,14,28} x C {64,128,256} (P2 of a 128^3 volume, this fork's (2,2,1) strides and the isotropic
variant), NMS3D over 1k..100k boxes.  Device-resident inputs, C-ABI calls, CUDA events, median of `reps`.
Prints a markdown table (stdout) and one JSON line per measurement (stderr).  Memory is bounded: combinations whose
crops tensor would exceed --max-gb are skipped (and said so)."""
import argparse
import ctypes
import json
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import roi3d_b200 as rb   # noqa: E402
import roi3d_synth        # noqa: E402


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--max-gb", type=float, default=24.0)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--nms-only", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = rb._lib.load()
    vp = ctypes.c_void_p
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
        if os.path.exists("MEASURED_PEAKS.json") else 6650.0
    vol = (128, 128, 128)
    print("## CropAndResize3D sweep (P2 of 128^3, B=1; GB/s = algorithmic bytes / time; %% of %.0f GB/s measured HBM)\n" % peak)
    print("| layout | C | pool | ROIs | fwd ms | fwd GB/s | fwd % | bwd ms | bwd GB/s | bwd % |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    rois_list = [64, 512, 4096] if args.quick else [64, 128, 256, 512, 1024, 2048, 4096, 8192]
    for iso in (() if args.nms_only else (False, True)):
        for C in (64, 128, 256):
            shape = roi3d_synth.level_shape(vol, 2, batch=1, channels=C, isotropic=iso)
            torch.manual_seed(C)
            image = torch.randn(shape, device=dev)
            gimg = torch.empty(shape, device=dev)
            for p in (7, 14, 28):
                for n in rois_list:
                    out_gb = n * p ** 3 * C * 4 / 1e9
                    if 2 * out_gb > args.max_gb:
                        print("| %s | %d | %d | %d | skipped: crops + grads would need %.0f GB | | | | | |" %
                              ("iso" if iso else "(2,2,1)", C, p, n, 2 * out_gb))
                        continue
                    boxes = roi3d_synth.rois(n, vol, seed=2000 + n)
                    bidx = np.zeros(n, np.int32)
                    tb, ti = torch.from_numpy(boxes).to(dev), torch.from_numpy(bidx).to(dev)
                    crops = torch.empty((n, p, p, p, C), device=dev)
                    grads = torch.randn((n, p, p, p, C), device=dev)
                    B, H, W, D, _ = shape
                    s = lambda: vp(torch.cuda.current_stream().cuda_stream)   # noqa: E731
                    ws = torch.empty(int(lib.roi3d_car3d_workspace_bytes(n)), dtype=torch.uint8, device=dev)   # caller-owned scratch (ROI order)
                    f = lambda: rb._lib.check(lib.roi3d_car3d_fwd_ws(vp(image.data_ptr()), B, H, W, D, C, vp(tb.data_ptr()), vp(ti.data_ptr()), n, p, p, p, 0, 0.0, vp(crops.data_ptr()), vp(ws.data_ptr()), ws.numel(), s()))   # noqa: E731
                    g = lambda: rb._lib.check(lib.roi3d_car3d_grad_image_ws(vp(grads.data_ptr()), vp(tb.data_ptr()), vp(ti.data_ptr()), n, p, p, p, B, H, W, D, C, 0, vp(gimg.data_ptr()), vp(ws.data_ptr()), ws.numel(), s()))   # noqa: E731
                    tf_, tb_ = timeit(f, args.reps), timeit(g, args.reps)
                    fb = roi3d_synth.car_algorithmic_bytes(boxes, shape, (p, p, p), False)
                    bb = roi3d_synth.car_algorithmic_bytes(boxes, shape, (p, p, p), True)
                    row = {"op": "car3d", "iso": iso, "C": C, "pool": p, "rois": n, "fwd_ms": tf_, "bwd_ms": tb_,
                           "fwd_gbs": fb / tf_ / 1e6, "bwd_gbs": bb / tb_ / 1e6}
                    print(json.dumps(row), file=sys.stderr)
                    print("| %s | %d | %d | %d | %.4f | %.0f | %.0f | %.4f | %.0f | %.0f |" %
                          ("iso" if iso else "(2,2,1)", C, p, n, tf_, row["fwd_gbs"], 100 * row["fwd_gbs"] / peak,
                           tb_, row["bwd_gbs"], 100 * row["bwd_gbs"] / peak))
                    del crops, grads
            del image, gimg
            torch.cuda.empty_cache()
    print("\n## NMS3D sweep (clustered boxes, thr 0.7, max_output_size = n/6 rounded, device-resident)\n")
    print("| boxes | max_out | kept | ms | boxes/s | pairs/s |")
    print("|---|---|---|---|---|---|")
    for n in ([1000, 6000, 20000] if args.quick else [1000, 2000, 6000, 20000, 50000, 100000]):
        mo = max(n // 6, 1) if n != 20000 else 2000
        nb, ns = roi3d_synth.nms_boxes(n, (256, 256, 256) if n > 6000 else vol)
        d_nb, d_ns = torch.from_numpy(nb).to(dev), torch.from_numpy(ns).to(dev)
        wsb = lib.roi3d_nms3d_workspace_bytes(n)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        keep = torch.empty(mo, dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        f = lambda: rb._lib.check(lib.roi3d_nms3d(vp(d_nb.data_ptr()), vp(d_ns.data_ptr()), n, mo, 0.7, vp(keep.data_ptr()), vp(cnt.data_ptr()), vp(ws.data_ptr()), wsb, vp(torch.cuda.current_stream().cuda_stream)))   # noqa: E731
        t = timeit(f, args.reps)
        row = {"op": "nms3d", "boxes": n, "max_out": mo, "kept": int(cnt.item()), "ms": t}
        print(json.dumps(row), file=sys.stderr)
        print("| %d | %d | %d | %.4f | %.3g | %.3g |" % (n, mo, row["kept"], t, n / t * 1e3, n * (n - 1) / 2 / t * 1e3))
        del ws
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
