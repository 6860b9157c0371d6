#!/usr/bin/env python
"""bench.py -- headline benchmark of the ROI hot path (BASELINE.json metric:
"3D ROIAlign fwd+bwd ROIs/s & HBM GB/s; NMS3D ms @6k boxes; 1/2/4/8 B200").

Workload (config.workload = "cfg2", BASELINE.json configs[1]): training_head_e2e step, batch 2 of
128^3 volumes, TRAIN_ROIS_PER_IMAGE=128 -> 256 ROIs routed over P2..P5 (C=256, this fork's (2,2,1)
strides), CropAndResize3D forward + grad-image for the classifier (7^3) and mask (14^3) heads:
per step up to 8 forward ops and 8 grad-image ops.  One "ROI" of the metric is one ROI taken through
all four (7^3 fwd, 14^3 fwd, 7^3 bwd, 14^3 bwd).

  value   ROIs/s with every input already resident in HBM (C-ABI calls on device pointers).
  e2e     ROIs/s through the public host-buffer API (pinned numpy/CPU tensors in, host results out):
          H2D of feature maps / boxes / grads and D2H of crops / grad images inside the timed region.
  N > 1   one process per GPU (torchrun), every rank runs its own batch (images never share ROIs, no
          collective on the data path): weak scaling, value = N * ROIs / max-over-ranks time.

`--impl reference` times the reference's CPU implementation of the same step (the oracle port of the
wheel's kernels, all host cores) on the same config.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import roi3d_synth  # noqa: E402

VOLUME = (128, 128, 128)
BATCH = 2
ROIS_PER_IMAGE = 128
CROPS = ((7, 7, 7), (14, 14, 14))
METRIC = "3D ROIAlign fwd+bwd ROIs/s"
UNIT = "ROIs/s"
WORKLOAD = ("cfg2: batch 2 x 128^3, 128 ROIs/image, P2-P5 C=256, CropAndResize3D fwd + grad-image "
            "for 7^3 and 14^3 heads (16 op calls/step)")


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------
WORKLOAD_NAME = "cfg2"


def run_cfg4(torch, rb, lib, dev, world, peak, ptr, stream, barrier):
    """BASELINE configs[3] on this rank: one 128^3 image, 1000 ROIs x 14^3 x 256 ch, CropAndResize3D forward + grad-image on
    the level that receives the ROIs (P2), C-ABI calls replayed from a CUDA graph; whole-job ROIs/s over all ranks."""
    import ctypes
    vp = ctypes.c_void_p
    crop, R = (14, 14, 14), 1000
    routed = roi3d_synth.pyramid_rois(R, 1, VOLUME, seed=2002)
    shape = roi3d_synth.level_shape(VOLUME, 2, batch=1)
    boxes, bidx, _ = routed[2]
    n = len(boxes)
    image = torch.from_numpy(roi3d_synth.feature_map(VOLUME, 2, batch=1)).to(dev)
    tb, ti = torch.from_numpy(boxes).to(dev), torch.from_numpy(bidx).to(dev)
    torch.manual_seed(44)
    grads = torch.randn((n,) + crop + (shape[4],), device=dev)
    crops = torch.empty_like(grads)
    gimg = torch.empty(shape, device=dev)
    B, H, W, D, C = shape
    fb = roi3d_synth.car_algorithmic_bytes(boxes, shape, crop, backward=False)
    bb = roi3d_synth.car_algorithmic_bytes(boxes, shape, crop, backward=True)
    car_ws = torch.empty(int(lib.roi3d_car3d_workspace_bytes(n)), dtype=torch.uint8, device=dev)   # caller-owned scratch (ROI order)

    def fwd():
        rb._lib.check(lib.roi3d_car3d_fwd_ws(ptr(image), B, H, W, D, C, ptr(tb), ptr(ti), n, 14, 14, 14, 0, 0.0, ptr(crops),
                                             ptr(car_ws), car_ws.numel(), stream()))

    def bwd():
        rb._lib.check(lib.roi3d_car3d_grad_image_ws(ptr(grads), ptr(tb), ptr(ti), n, 14, 14, 14, B, H, W, D, C, 0, ptr(gimg),
                                                    ptr(car_ws), car_ws.numel(), stream()))

    def ev_time(fn, reps):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        barrier()
        return rb.sharding.max_over_ranks(a.elapsed_time(b)) / reps
    t_f, t_b = ev_time(fwd, 10), ev_time(bwd, 10)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fwd()
        bwd()
    t_s = ev_time(gr.replay, 10)
    del gr, grads, crops, gimg, image
    torch.cuda.empty_cache()
    return {"workload": "cfg4: one 128^3 image per GPU, %d ROIs (all on P2) x 14^3 x 256 ch, fwd + grad-image" % n,
            "ms_per_step": round(t_s, 4), "value": round(world * n / (t_s * 1e-3), 1), "unit": UNIT, "n_gpus": world,
            "fwd_ms": round(t_f, 4), "fwd_frac": round(fb / (t_f * 1e-3) / 1e9 / peak, 4),
            "bwd_ms": round(t_b, 4), "bwd_frac": round(bb / (t_b * 1e-3) / 1e9 / peak, 4)}


def run_mixed_levels(torch, rb, lib, dev, images, ptr, stream):
    """PyramidROIAlign where the ROIs spread over the levels (sides 32-128 px -> P2/P3/P4): the drop-in path (host
    routing, one op call per level, crops gathered back into ROI order = what core/models.py:663-683 executes) against
    the fused entry points.  Both legs are CUDA-graph replays of forward + backward for the 7^3 and 14^3 pools."""
    vp = ctypes.c_void_p
    R = ROIS_PER_IMAGE
    car_ws = torch.empty(int(lib.roi3d_car3d_workspace_bytes(BATCH * R)), dtype=torch.uint8, device=dev)
    boxes_br = np.stack([roi3d_synth.rois(R, VOLUME, 7000 + b, side_px=(32.0, 128.0)) for b in range(BATCH)])
    flat = boxes_br.reshape(-1, 6)
    bidx_all = np.repeat(np.arange(BATCH, dtype=np.int32), R)
    lv = roi3d_synth.roi_levels(flat, VOLUME)
    fms = [images[l] for l in roi3d_synth.LEVELS]
    C = fms[0].shape[4]
    d_boxes_br = torch.from_numpy(boxes_br).to(dev)
    per_level = []
    for li, level in enumerate(roi3d_synth.LEVELS):
        sel = np.nonzero(lv == level)[0]
        per_level.append({"n": len(sel), "sel": torch.from_numpy(sel).to(dev), "fm": fms[li],
                          "boxes": torch.from_numpy(np.ascontiguousarray(flat[sel])).to(dev),
                          "bidx": torch.from_numpy(np.ascontiguousarray(bidx_all[sel])).to(dev),
                          "gimg": {c: torch.empty_like(fms[li]) for c in CROPS}})
    order = torch.argsort(torch.cat([pl["sel"] for pl in per_level]))       # level-sorted rows -> ROI order
    crops_lv = {c: torch.empty((BATCH * R,) + c + (C,), device=dev) for c in CROPS}      # level-sorted crops
    pooled = {c: torch.empty((BATCH * R,) + c + (C,), device=dev) for c in CROPS}        # ROI order (the layer's output)
    grads = {c: torch.randn((BATCH * R,) + c + (C,), device=dev) for c in CROPS}         # ROI order (the layer's input gradient)
    grads_lv = {c: torch.empty_like(grads[c]) for c in CROPS}
    inv = torch.cat([pl["sel"] for pl in per_level])

    def per_op_step(glue):
        for c in CROPS:
            off = 0
            for pl in per_level:
                B_, H, W, D, _ = pl["fm"].shape
                out = crops_lv[c][off:off + pl["n"]]
                rb._lib.check(lib.roi3d_car3d_fwd_ws(ptr(pl["fm"]), B_, H, W, D, C, ptr(pl["boxes"]), ptr(pl["bidx"]), pl["n"],
                                                     c[0], c[1], c[2], 0, 0.0, vp(out.data_ptr()), ptr(car_ws), car_ws.numel(), stream()))
                off += pl["n"]
            if glue:
                torch.index_select(crops_lv[c], 0, order, out=pooled[c])        # tf.gather(pooled, ix), core/models.py:675-683
        for c in CROPS:
            if glue:
                torch.index_select(grads[c], 0, inv, out=grads_lv[c])           # its gradient: rows back into level order
            src = grads_lv[c] if glue else grads[c]
            off = 0
            for pl in per_level:
                B_, H, W, D, _ = pl["fm"].shape
                gin = src[off:off + pl["n"]]
                rb._lib.check(lib.roi3d_car3d_grad_image_ws(vp(gin.data_ptr()), ptr(pl["boxes"]), ptr(pl["bidx"]), pl["n"], c[0], c[1], c[2],
                                                            B_, H, W, D, C, 0, ptr(pl["gimg"][c]), ptr(car_ws), car_ws.numel(), stream()))
                off += pl["n"]

    lshapes = (ctypes.c_int * 12)(*[int(d) for fm in fms for d in fm.shape[1:4]])
    ishape = (ctypes.c_float * 3)(*[float(v) for v in VOLUME])
    fm_ptrs = (vp * 4)(*[fm.data_ptr() for fm in fms])
    gm_ptrs = {c: (vp * 4)(*[pl["gimg"][c].data_ptr() for pl in per_level]) for c in CROPS}

    def fused_step():
        for c in CROPS:
            rb._lib.check(lib.roi3d_pyramid_roi_align_fwd_ws(fm_ptrs, lshapes, BATCH, C, ptr(d_boxes_br), R, ishape, c[0], c[1], c[2],
                                                             ptr(pooled[c]), ptr(car_ws), car_ws.numel(), stream()))
        for c in CROPS:
            rb._lib.check(lib.roi3d_pyramid_roi_align_grad_ws(ptr(grads[c]), gm_ptrs[c], lshapes, BATCH, C, ptr(d_boxes_br), R, ishape,
                                                              c[0], c[1], c[2], ptr(car_ws), car_ws.numel(), stream()))

    def graph_ms(fn, reps=20):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
        for _ in range(3):
            gr.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            gr.replay()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    res = {"rois_per_level": {"P%d" % l: pl["n"] for l, pl in zip(roi3d_synth.LEVELS, per_level)},
           "per_op_ms": round(graph_ms(lambda: per_op_step(False)), 4),
           "per_op_with_reorder_ms": round(graph_ms(lambda: per_op_step(True)), 4),
           "fused_ms": round(graph_ms(fused_step), 4),
           "note": "ROI sides 32-128 px; per_op = one C-ABI call per level and pool, crops left in level order; "
                   "per_op_with_reorder adds the layer's tf.gather back into ROI order and its gradient (torch.index_select "
                   "stands in for the TF glue); fused = roi3d_pyramid_roi_align_fwd/grad (routing and order restore inside)"}
    return res


def set_workload(name):
    """cfg2 (default, BASELINE configs[1]) or cfg4 (configs[3]: mask-head stress, 1000 ROIs x 14^3 x 256 ch, one image per GPU)."""
    global BATCH, ROIS_PER_IMAGE, CROPS, WORKLOAD, WORKLOAD_NAME
    WORKLOAD_NAME = name
    if name == "cfg4":
        BATCH, ROIS_PER_IMAGE, CROPS = 1, 1000, ((14, 14, 14),)
        WORKLOAD = ("cfg4: mask-head stress, one 128^3 image per GPU, 1000 ROIs x 14^3 x 256 ch over P2-P5, "
                    "CropAndResize3D fwd + grad-image (8 op calls/step)")
    elif name != "cfg2":
        raise SystemExit("unknown workload %r" % name)


def make_workload(seed=2002, with_data=True):
    routed = roi3d_synth.pyramid_rois(ROIS_PER_IMAGE, BATCH, VOLUME, seed=seed)
    ops = []
    for level in roi3d_synth.LEVELS:
        boxes, bidx, _ = routed[level]
        shape = roi3d_synth.level_shape(VOLUME, level, batch=BATCH)
        image = roi3d_synth.feature_map(VOLUME, level, batch=BATCH) if with_data else None
        for crop in CROPS:
            n = len(boxes)
            ops.append({
                "level": level, "crop": crop, "n": n, "shape": shape, "image": image,
                "boxes": boxes, "bidx": bidx,
                "grads": roi3d_synth.grads_like((n,) + crop + (shape[4],), 4000 + level * 10 + crop[0])
                if with_data else None,
                "fwd_bytes": roi3d_synth.car_algorithmic_bytes(boxes, shape, crop, backward=False) if n else 0,
                "bwd_bytes": roi3d_synth.car_algorithmic_bytes(boxes, shape, crop, backward=True),
            })
    return ops


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, enabled=True):
        self.rows, self.proc, self.index, self.enabled = [], None, index, enabled

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import roi3d_b200 as rb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = rb._lib.load()
    vp = ctypes.c_void_p
    stream = lambda: vp(torch.cuda.current_stream().cuda_stream)   # noqa: E731
    if os.environ.get("BENCH_CAR_EXPERIMENT"):                     # A/B switches of profiles/*.py (e.g. 16: ROIs in the order given)
        rb.set_option("car_experiment", int(os.environ["BENCH_CAR_EXPERIMENT"]))

    ops = make_workload(seed=2002)         # weak scaling: every rank runs the SAME batch, so per-GPU work is fixed exactly
    seed_is_default = True                 # (the ncu traffic figures were captured on this ROI set)
    total_rois = BATCH * ROIS_PER_IMAGE
    # ---- device-resident state ---------------------------------------------------------------
    images = {}
    for op in ops:
        lv = op["level"]
        if lv not in images:
            images[lv] = torch.from_numpy(op["image"]).to(dev)
        op["d_image"] = images[lv]
        op["d_boxes"] = torch.from_numpy(op["boxes"]).to(dev)
        op["d_bidx"] = torch.from_numpy(op["bidx"]).to(dev)
        op["d_grads"] = torch.from_numpy(op["grads"]).to(dev)
        op["d_crops"] = torch.empty_like(op["d_grads"])
        op["d_gimg"] = torch.empty(op["shape"], dtype=torch.float32, device=dev)

    def ptr(t):
        return vp(t.data_ptr() if t.numel() else 0)

    # caller-owned scratch of the crop-and-resize calls (the ROI processing order), one per op node (nodes may run on
    # different streams with --op-streams)
    for op in ops:
        op["d_ws"] = torch.empty(int(lib.roi3d_car3d_workspace_bytes(max(op["n"], 1))), dtype=torch.uint8, device=dev)

    def fwd(op):
        B, H, W, D, C = op["shape"]
        c = op["crop"]
        rb._lib.check(lib.roi3d_car3d_fwd_ws(ptr(op["d_image"]), B, H, W, D, C, ptr(op["d_boxes"]), ptr(op["d_bidx"]),
                                             op["n"], c[0], c[1], c[2], 0, 0.0, ptr(op["d_crops"]), ptr(op["d_ws"]), op["d_ws"].numel(), stream()))

    def bwd(op):
        B, H, W, D, C = op["shape"]
        c = op["crop"]
        rb._lib.check(lib.roi3d_car3d_grad_image_ws(ptr(op["d_grads"]), ptr(op["d_boxes"]), ptr(op["d_bidx"]), op["n"],
                                                    c[0], c[1], c[2], B, H, W, D, C, 0, ptr(op["d_gimg"]), ptr(op["d_ws"]), op["d_ws"].numel(), stream()))

    # The 8 forward nodes of a step are independent of each other, and so are the 8 grad-image nodes (each writes its
    # own tensor; TF's executor schedules such nodes concurrently, SURVEY.md 8b "Threading / streams").  With
    # --op-streams N > 1 the non-empty ops of each phase are spread over N streams that fork from and join into the
    # main stream; the backward phase starts only after every forward has finished, as in training.
    side = [torch.cuda.Stream(dev) for _ in range(max(args.op_streams, 1) - 1)]

    def phase(fn):
        if not side:
            for op in ops:
                fn(op)
            return
        cur = torch.cuda.current_stream()
        lanes = [cur] + side
        for s_ in side:
            s_.wait_stream(cur)
        busy = sorted((op for op in ops if op["n"]), key=lambda o: -o["fwd_bytes"])
        idle = [op for op in ops if not op["n"]]
        for i, op in enumerate(busy + idle):
            with torch.cuda.stream(lanes[i % len(lanes)]):
                fn(op)
        for s_ in side:
            cur.wait_stream(s_)

    def step():
        phase(fwd)
        phase(bwd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The step is launch-heavy (16 C-ABI calls, 12 of them on empty levels): capture it once in a CUDA graph and time
    # replays.  Every call is stream-ordered and allocation-free, so the capture is exact; --no-graph times eager calls.
    graph = None
    rb.reset_kernel_launches()
    if not args.no_graph:
        for _ in range(2):
            step()                                      # cudaFuncSetAttribute etc. happen outside the capture
        torch.cuda.synchronize()
        rb.reset_kernel_launches()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
    launches_per_step = rb.kernel_launches() if graph is not None else None
    run_step = graph.replay if graph is not None else step

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local, enabled=(rank == 0)) as clocks:      # one nvidia-smi poller per job, not per rank
        # the sampler runs from the warm-up (same kernels, same load) through the timed region so that a
        # millisecond-scale region still yields several nvidia-smi samples under load
        t_w = time.perf_counter()
        nw = 0
        while nw < max(args.warmup, 3) or time.perf_counter() - t_w < 0.6:
            run_step()
            nw += 1
            if nw % 50 == 0:
                torch.cuda.synchronize()
        barrier()
        if graph is None:
            rb.reset_kernel_launches()
        barrier()
        e0.record()
        for _ in range(args.steps):
            run_step()
        e1.record()
        barrier()
    launches = launches_per_step * args.steps if graph is not None else rb.kernel_launches()
    ms_by_rank = [round(v / args.steps, 4) for v in rb.sharding.gather_over_ranks(e0.elapsed_time(e1))]
    ms_total = rb.sharding.max_over_ranks(e0.elapsed_time(e1))     # device time, slowest rank
    ms_step = ms_total / args.steps
    value = world * total_rois / (ms_step * 1e-3)

    # ---- per-op timing (CUDA events on the launching stream) -> roofline of the dominant kernel ----
    peak, peak_src = load_peaks()
    per_op = []
    reps = max(args.steps, 30)

    def pct(v, q):
        v = sorted(v)
        return v[min(len(v) - 1, int(q * len(v)))]

    for kind, fn in (("fwd", fwd), ("bwd", bwd)):
        for op in ops:
            evs = []
            for _ in range(3):
                fn(op)
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn(op)
                b.record()
                evs.append((a, b))
            torch.cuda.synchronize()
            t = [a.elapsed_time(b) for a, b in evs]
            ms = statistics.mean(t)
            nbytes = op[kind + "_bytes"]
            per_op.append({"op": "car3d_" + ("fwd" if kind == "fwd" else "grad_image"), "level": op["level"],
                           "crop": op["crop"][0], "n": op["n"], "ms": round(ms, 4), "ms_p10": round(pct(t, 0.1), 4),
                           "ms_p50": round(pct(t, 0.5), 4), "ms_p90": round(pct(t, 0.9), 4), "reps": reps, "alg_bytes": nbytes,
                           "gbs": round(nbytes / (ms * 1e-3) / 1e9, 1) if ms > 0 else None})
    dom = max(per_op, key=lambda r: r["ms"])
    dom_name = "%s P%d crop %d^3 n=%d" % (dom["op"], dom["level"], dom["crop"], dom["n"])
    try:                                   # DRAM bytes per launch measured once with `ncu --set full` (profiles/README.md)
        with open(os.path.join(ROOT, "profiles", "traffic_r2.json")) as f:
            traffic = json.load(f).get(dom_name) if rank == 0 and seed_is_default else None
    except Exception:  # noqa: BLE001
        traffic = None
    roofline = {"bound": "hbm", "kernel": dom_name,
                "achieved": dom["gbs"], "peak": peak, "unit": "GB/s", "frac": round(dom["gbs"] / peak, 4),
                "traffic": traffic, "peak_source": peak_src,
                "ms": dom["ms"], "ms_p10": dom["ms_p10"], "ms_p50": dom["ms_p50"], "ms_p90": dom["ms_p90"], "reps": reps,
                "note": "achieved = algorithmic bytes (SURVEY.md 8d) / mean CUDA-event duration of the C-ABI call"
                        + (" (zero-fill kernel + scatter kernel); the 8d formula charges the footprint of every ROI as a DRAM "
                           "read-modify-write, but overlapping footprints and the tail of the zero-fill are served by the 126 MB "
                           "L2, so frac can exceed 1: achieved_dram / frac_dram divide the DRAM bytes ncu measured for this call "
                           "(traffic) by the same time" if dom["op"].endswith("grad_image") else "")}
    if traffic:
        roofline["achieved_dram"] = round(traffic / (dom["ms"] * 1e-3) / 1e9, 1)
        roofline["frac_dram"] = round(roofline["achieved_dram"] / peak, 4)
    sum_bytes = sum(r["alg_bytes"] for r in per_op)
    step_gbs = sum_bytes / (ms_step * 1e-3) / 1e9

    # ---- the same step through the fused PyramidROIAlign3D entry points: 2 + 2 launches ----
    boxes_br = np.stack([roi3d_synth.rois(ROIS_PER_IMAGE, VOLUME, (2002 + rank) * 131 + b) for b in range(BATCH)])
    d_boxes_br = torch.from_numpy(boxes_br).to(dev)
    fms = [images[lv] for lv in roi3d_synth.LEVELS]
    C = fms[0].shape[4]
    lshapes = (ctypes.c_int * 12)(*[int(d) for fm in fms for d in fm.shape[1:4]])
    ishape = (ctypes.c_float * 3)(*[float(v) for v in VOLUME])
    fm_ptrs = (vp * 4)(*[fm.data_ptr() for fm in fms])
    pooled = {c: torch.empty((BATCH, ROIS_PER_IMAGE) + c + (C,), device=dev) for c in CROPS}
    pgrads = {c: torch.randn((BATCH, ROIS_PER_IMAGE) + c + (C,), device=dev) for c in CROPS}
    gms = {c: [torch.empty_like(fm) for fm in fms] for c in CROPS}
    gm_ptrs = {c: (vp * 4)(*[g.data_ptr() for g in gms[c]]) for c in CROPS}

    pyr_ws = torch.empty(int(lib.roi3d_car3d_workspace_bytes(BATCH * ROIS_PER_IMAGE)), dtype=torch.uint8, device=dev)

    def fused_step():
        for c in CROPS:
            rb._lib.check(lib.roi3d_pyramid_roi_align_fwd_ws(fm_ptrs, lshapes, BATCH, C, ptr(d_boxes_br), ROIS_PER_IMAGE, ishape,
                                                             c[0], c[1], c[2], ptr(pooled[c]), ptr(pyr_ws), pyr_ws.numel(), stream()))
        for c in CROPS:
            rb._lib.check(lib.roi3d_pyramid_roi_align_grad_ws(ptr(pgrads[c]), gm_ptrs[c], lshapes, BATCH, C, ptr(d_boxes_br),
                                                              ROIS_PER_IMAGE, ishape, c[0], c[1], c[2], ptr(pyr_ws), pyr_ws.numel(), stream()))

    for _ in range(3):
        fused_step()
    torch.cuda.synchronize()
    rb.reset_kernel_launches()
    run_fused = fused_step
    if not args.no_graph:                               # same launch mode as the per-op step it is compared with
        fgraph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(fgraph):
            fused_step()
        run_fused = fgraph.replay
    fused_launches = rb.kernel_launches()
    for _ in range(5):
        run_fused()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        run_fused()
    f1.record()
    barrier()
    fused_ms = rb.sharding.max_over_ranks(f0.elapsed_time(f1)) / args.steps
    fused = {"ms_per_step": round(fused_ms, 4), "value": round(world * total_rois / (fused_ms * 1e-3), 1), "unit": UNIT,
             "launches_per_step": int(fused_launches) if not args.no_graph else None,
             "vs_per_op_step": round(fused_ms / ms_step, 4),
             "note": "roi3d_pyramid_roi_align_fwd/grad: routing + 4 levels + order restore in one launch per pool shape, one "
                     "zero-fill kernel for the four grad maps; same launch mode (CUDA graph replay) as the per-op step"}
    if not args.no_graph:
        del fgraph
    if WORKLOAD_NAME == "cfg2" and not args.no_graph:
        try:
            fused["mixed_levels"] = run_mixed_levels(torch, rb, lib, dev, images, ptr, stream)
        except torch.cuda.OutOfMemoryError:
            fused["mixed_levels"] = {"skipped": "out of memory"}
        torch.cuda.empty_cache()
    # float16 output (the target files' payload) written by the crop kernel vs float32 crop + separate conversion
    f16 = {}
    for c in CROPS:
        def _f16(c=c):
            return rb.pyramid_roi_align_3d(d_boxes_br, VOLUME, fms, c, out_dtype=torch.float16)

        def _f32_then_pack(c=c):
            with torch.no_grad():
                return rb.pack_f16(rb.pyramid_roi_align_3d(d_boxes_br, VOLUME, fms, c))
        ts = []
        for fn in (_f16, _f32_then_pack):
            ms = []
            for it in range(12):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                if it >= 4:
                    ms.append(a.elapsed_time(b))
            ts.append(round(statistics.median(ms), 4))
        f16["crop%d" % c[0]] = {"fused_f16_ms": ts[0], "f32_then_pack_ms": ts[1]}
    fused["f16_output"] = f16
    del pooled, pgrads, gms

    # ---- NMS3D: cfg1 (6000 -> 1000 @0.7), cfg3 (20000 -> 2000 @0.7), and the dense-cluster regime a trained RPN produces ----
    def time_nms(n, max_out, thr, vol, reps=50, **kw):
        bx, sc = roi3d_synth.nms_boxes(n, vol, **kw)
        d_b, d_s = torch.from_numpy(bx).to(dev), torch.from_numpy(sc).to(dev)
        nws = lib.roi3d_nms3d_workspace_bytes(n)
        wsp = torch.empty(nws, dtype=torch.uint8, device=dev)
        kp = torch.empty(max_out, dtype=torch.int32, device=dev)
        ct = torch.zeros(1, dtype=torch.int32, device=dev)

        def call():
            rb._lib.check(lib.roi3d_nms3d(ptr(d_b), ptr(d_s), n, max_out, thr, ptr(kp), ptr(ct), ptr(wsp), nws, stream()))

        def loop(fn):
            out = []
            for it in range(reps + 10):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                if it >= 10:
                    out.append(a.elapsed_time(b))
            return sorted(out)
        eager = loop(call)
        gr = torch.cuda.CUDAGraph()                     # the call is stream-ordered and allocation-free: capturable as is
        with torch.cuda.graph(gr):
            call()
        replay = loop(gr.replay)
        del gr
        med = eager[len(eager) // 2]
        return {"boxes": n, "max_out": max_out, "iou_threshold": thr, "kept": int(ct.item()), "ms": round(med, 4),
                "ms_p10": round(eager[len(eager) // 10], 4), "ms_p90": round(eager[len(eager) * 9 // 10], 4),
                "ms_graph_replay": round(replay[len(replay) // 2], 4), "boxes_per_s": round(n / (med * 1e-3)),
                "pairs_per_s": round(n * (n - 1) / 2 / (med * 1e-3))}, (bx, sc, kp, ct)

    nms, (nb, ns, keep, cnt) = time_nms(6000, 1000, 0.7, VOLUME)
    nms20, _ = time_nms(20000, 2000, 0.7, (256, 256, 256), reps=30)
    nms_dense, _ = time_nms(6000, 1000, 0.7, VOLUME, reps=30, cluster=64, jitter=0.05)
    nms20_dense, _ = time_nms(20000, 2000, 0.7, (256, 256, 256), reps=20, cluster=64, jitter=0.05)
    d_nb, d_ns = torch.from_numpy(nb).to(dev), torch.from_numpy(ns).to(dev)
    for _ in range(3):                                  # warm the workspace / pinned-buffer caches
        rb.non_max_suppression_3d(nb, ns, 1000, 0.7)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        kept = rb.non_max_suppression_3d(nb, ns, 1000, 0.7)
    nms_e2e_ms = (time.perf_counter() - t0) / 50 * 1e3
    # batched: the 8 images of BASELINE cfg4 in one set of launches (ProposalLayer's batch_slice loop, core/models.py:487)
    nbb = torch.cat([d_nb] * 8)
    nsb = torch.cat([d_ns] * 8)
    offs = torch.arange(9, dtype=torch.int32, device=dev) * 6000
    wsb8 = lib.roi3d_nms3d_batched_workspace_bytes(6000, 8)
    ws8 = torch.empty(wsb8, dtype=torch.uint8, device=dev)
    keep8 = torch.empty(8 * 1000, dtype=torch.int32, device=dev)
    cnt8 = torch.zeros(8, dtype=torch.int32, device=dev)
    b_ms = []
    for it in range(30):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rb._lib.check(lib.roi3d_nms3d_batched(ptr(nbb), ptr(nsb), ptr(offs), 8, 6000, 1000, 0.7, ptr(keep8), ptr(cnt8),
                                              ptr(ws8), wsb8, stream()))
        b.record()
        torch.cuda.synchronize()
        if it >= 5:
            b_ms.append(a.elapsed_time(b))
    del ws8
    nms.update({"batched_8x6000_ms": round(statistics.median(b_ms), 4), "e2e_host_buffers_ms": round(nms_e2e_ms, 4),
                "e2e_kept": int(len(kept)), "cfg3_20000_to_2000": nms20, "dense_clusters_6000": nms_dense,
                "dense_clusters_20000": nms20_dense})

    # ---- ProposalLayer on the device (cfg1 front half): top-k 6000 of N anchors -> decode -> NMS3D -> pad ----
    n_anchor = 393216                                   # 32*32*128*3 anchors of the finest RPN level at 128^3
    an = torch.from_numpy(roi3d_synth.nms_boxes(n_anchor, VOLUME, seed=77)[0]).to(dev)
    torch.manual_seed(78)
    dlt = torch.randn((n_anchor, 6), device=dev) * 0.5
    scr = torch.rand(n_anchor, device=dev)
    pl_ms = []
    for it in range(25):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        props, pcount = rb.proposal_layer(scr, dlt, an, (0.1, 0.1, 0.1, 0.2, 0.2, 0.2), VOLUME[2], 6000, 1000, 0.7)
        b.record()
        torch.cuda.synchronize()
        if it >= 5:
            pl_ms.append(a.elapsed_time(b))
    proposal = {"anchors": n_anchor, "pre_nms_limit": 6000, "proposal_count": 1000, "kept": int(pcount.item()),
                "ms": round(statistics.median(pl_ms), 4), "note": "top-k radix select + decode + NMS3D + gather/pad, no host sync"}
    del an, dlt, scr

    # ---- rows f3 / f4: DetectionLayer, mask targets, target-file payloads -------------------
    def timed(fn, reps=25, warm=5):
        out = []
        for it in range(reps + warm):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            if it >= warm:
                out.append(a.elapsed_time(b))
        return statistics.median(out)

    # cfg3 detection half: 2000 refined ROIs per image, 2 classes, DETECTION_MAX_INSTANCES keeps; batch 8 in one call
    rng = np.random.default_rng(4242)
    det_B, det_R, det_M = 8, 2000, 200
    d_rois = torch.from_numpy(np.stack([roi3d_synth.nms_boxes(det_R, VOLUME, seed=500 + i)[0] for i in range(det_B)])).to(dev)
    d_probs = torch.softmax(torch.from_numpy(rng.standard_normal((det_B, det_R, 2)).astype(np.float32) * 2), -1).to(dev)
    d_dl = torch.from_numpy(rng.standard_normal((det_B, det_R, 2, 6)).astype(np.float32)).to(dev)
    det_ms = timed(lambda: rb.refine_detections(d_rois, d_probs, d_dl, VOLUME, 0.5, 0.3, None, det_M))
    det1_ms = timed(lambda: rb.refine_detections(d_rois[0], d_probs[0], d_dl[0], VOLUME, 0.5, 0.3, None, det_M))
    _, det_cnt = rb.refine_detections(d_rois, d_probs, d_dl, VOLUME, 0.5, 0.3, None, det_M, return_counts=True)
    detection = {"images": det_B, "rois_per_image": det_R, "max_instances": det_M, "ms_batch8": round(det_ms, 4),
                 "ms_single_image": round(det1_ms, 4), "kept": [int(v) for v in det_cnt.cpu()],
                 "note": "decode + filters folded into NMS3D candidates + gather/normalise/pad; 5 launches, no host sync"}
    del d_rois, d_probs, d_dl
    # mask targets: 128 positive ROIs x 28^3 from uint8 ground-truth masks of a 128^3 volume, rounded + bit-packed
    mt_G, mt_n = 32, 128
    gmask = (torch.rand((mt_G,) + tuple(VOLUME), device=dev) > 0.7).to(torch.uint8)
    mt_boxes = torch.from_numpy(roi3d_synth.rois(mt_n, VOLUME, seed=9)).to(dev)
    mt_assign = torch.randint(0, mt_G, (mt_n,), device=dev, dtype=torch.int32)
    mt_ms = timed(lambda: rb.mask_targets(gmask, mt_boxes, mt_assign, (28, 28, 28), packed=True))
    mask_t = {"rois": mt_n, "mask_shape": [28, 28, 28], "gt_masks": mt_G, "ms": round(mt_ms, 4),
              "out_gbs": round(mt_n * 28 ** 3 * (4 + 0.125) / mt_ms / 1e6, 1)}
    del gmask
    # target-file payloads: rois_aligned of 1000 ROIs x 7^3 x 256 -> float16; a 1000 x 28^3 mask -> bits
    ra = torch.randn((1000, 7, 7, 7, 256), device=dev)
    tm = (torch.rand((1000, 28, 28, 28), device=dev) > 0.5).float()
    f16_ms = timed(lambda: rb.pack_f16(ra))
    h16 = rb.pack_f16(ra)
    uf16_ms = timed(lambda: rb.unpack_f16(h16))
    pb_ms = timed(lambda: rb.pack_bits(tm))
    bits_t, _ = rb.pack_bits(tm)
    ub_ms = timed(lambda: rb.unpack_bits(bits_t, tm.shape))
    wire = {"pack_f16_gbs": round(ra.numel() * 6 / f16_ms / 1e6, 1), "unpack_f16_gbs": round(ra.numel() * 6 / uf16_ms / 1e6, 1),
            "pack_bits_gbs": round(tm.numel() * 4.125 / pb_ms / 1e6, 1), "unpack_bits_gbs": round(tm.numel() * 4.125 / ub_ms / 1e6, 1),
            "note": "algorithmic bytes (read + write) / event time, through the Python mirror (includes its allocation)"}
    del ra, tm, h16, bits_t

    # ---- end to end through the public API with host buffers ---------------------------------------
    # Per step: every level's feature map is uploaded once (pinned host -> device), boxes / box indices / grads
    # go in as host buffers with each call, every crop and every grad image comes back to pinned host memory.
    # `rb.deferred()` lets the 16 independent op calls overlap their copies with each other's kernels.
    h_images = {lv: torch.from_numpy(op["image"]).pin_memory() for lv, op in ((o["level"], o) for o in ops)}
    h = [{"boxes": torch.from_numpy(op["boxes"]), "bidx": torch.from_numpy(op["bidx"]),
          "grads": torch.from_numpy(op["grads"]).pin_memory()} for op in ops]
    h2d = sum(t.numel() * 4 for t in h_images.values()) + sum(x["grads"].numel() * 4 for x in h) + \
        2 * sum(x["boxes"].numel() * 4 + x["bidx"].numel() * 4 for x in h)
    # crops of every op + grad images of the ops that have boxes (an empty op's zero-filled grad image is produced
    # directly in host memory by the host-buffer API: nothing crosses PCIe for it)
    d2h = sum(x["grads"].numel() * 4 for x in h) + sum(int(np.prod(op["shape"])) * 4 for op in ops if op["n"] > 0)

    # The 16 op nodes of a step are independent; like an executor that runs ready nodes first, the step issues
    #  1. the grad-image nodes of the smallest pool first: they need only their own 88 MB of grads, so the download engine
    #     (which carries more bytes than the upload engine and bounds the step) gets its first 268 MB result after ~2 ms
    #     and moves it while the P2 map goes up; then that pool's forward nodes (the map is uploaded right before its
    #     first use), then the larger pool's forward and grad-image nodes.  A two-engine model of the step with this
    #     box's copy rates (55 / 56 GB/s alone, 47 GB/s each way when both are busy) gives 31.3 ms for "forward first"
    #     and 28.7 ms for this order,
    #  2. the forward nodes of the empty levels (their maps are still uploaded: they are inputs of the call),
    #  3. the grad-image nodes of the EMPTY levels: their result is the op's zero-fill, which the host-buffer API writes
    #     straight into host memory -- on this thread, while the copies queued above are in flight.
    order_b0 = [i for i, op in enumerate(ops) if op["n"] == 0]
    order_b1 = [i for i, op in enumerate(ops) if op["n"] > 0]

    def e2e_step():
        with rb.deferred():
            outs = [None] * (2 * len(ops))
            d_img = {}

            def dmap(lv):
                if lv not in d_img:
                    d_img[lv] = rb.upload(h_images[lv])       # on the pipeline's upload stream: uploads keep their issue order
                return d_img[lv]
            def fwd_nodes(crop):
                for i in order_b1:
                    if ops[i]["crop"] == crop:      # mixed call: device-resident map, host boxes -> host result
                        outs[i] = rb.crop_and_resize_3d(dmap(ops[i]["level"]), h[i]["boxes"], h[i]["bidx"], crop)

            def bwd_nodes(crop):
                for i in order_b1:
                    if ops[i]["crop"] == crop:
                        outs[len(ops) + i] = rb.crop_and_resize_3d_grad_image(h[i]["grads"], h[i]["boxes"], h[i]["bidx"], ops[i]["shape"])
            for j, crop in enumerate(sorted(CROPS)):
                if j == 0 and len(CROPS) > 1 and os.environ.get("BENCH_E2E_ORDER", "ready-first") != "forward-first":   # smallest input first (see above)
                    bwd_nodes(crop)
                    fwd_nodes(crop)
                else:
                    fwd_nodes(crop)
                    bwd_nodes(crop)
            for i in order_b0:
                outs[i] = rb.crop_and_resize_3d(dmap(ops[i]["level"]), h[i]["boxes"], h[i]["bidx"], ops[i]["crop"])
            for i in order_b0:                          # host-side zero-fill: runs while the copies above are in flight
                outs[len(ops) + i] = rb.crop_and_resize_3d_grad_image(h[i]["grads"], h[i]["boxes"], h[i]["bidx"], ops[i]["shape"])
        return outs

    # Copy policy: full duplex (each result goes down as soon as its kernel has finished).  The alternative, every rank
    # uploading its whole step before any download starts (BENCH_E2E_PIPE=uploads-first), was measured on the 8-GPU box
    # because its probe shows more one-directional than duplex host bandwidth (233 / 128 vs 82 + 82 GB/s): 174.0 vs
    # 169.8 ms/step -- no gain, the shared host is the limit either way.
    pipe_mode = os.environ.get("BENCH_E2E_PIPE", "duplex")
    rb.host_pipeline(uploads_first=(pipe_mode == "uploads-first"))
    e2e_steps = max(2, min(args.steps, 5))
    if args.no_e2e:
        e2e_steps, e2e_s = 1, float("inf")
    else:
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_s = time.perf_counter() - t0
    rb.host_pipeline(uploads_first=False)
    e2e_by_rank = rb.sharding.gather_over_ranks(e2e_s)
    e2e_s = rb.sharding.max_over_ranks(e2e_s)
    e2e_value = world * total_rois * e2e_steps / e2e_s

    # ---- BASELINE configs[3] (cfg4: mask-head stress, one image per GPU, 1000 ROIs x 14^3 x 256 ch) on every rank ----
    cfg4 = None
    if WORKLOAD_NAME == "cfg2" and not args.no_cfg4:
        try:
            cfg4 = run_cfg4(torch, rb, lib, dev, world, peak, ptr, stream, barrier)
        except torch.cuda.OutOfMemoryError:             # bounded: never take the box down for an extra key
            cfg4 = {"skipped": "out of memory"}
    # the other half of BASELINE's metric and the rows the driver record keeps only inside `roofline`
    roofline["secondary"] = {
        "nms3d_ms_6k": nms["ms"], "nms3d_ms_6k_graph_replay": nms["ms_graph_replay"], "nms3d_boxes_per_s_6k": nms["boxes_per_s"],
        "nms3d_pairs_per_s_6k": nms["pairs_per_s"], "nms3d_kept_6k": nms["kept"],
        "nms3d_ms_20k": nms20["ms"], "nms3d_boxes_per_s_20k": nms20["boxes_per_s"], "nms3d_pairs_per_s_20k": nms20["pairs_per_s"],
        "nms3d_ms_6k_dense_clusters": nms_dense["ms"], "nms3d_ms_20k_dense_clusters": nms20_dense["ms"],
        "nms3d_batched_8x6000_ms": nms["batched_8x6000_ms"],
        "nms3d_bound": "SM fp32 ALU / L2 (IoU bitmask) + latency (greedy scan); ncu counters: profiles/r2_nms_ncu_summary.txt",
        "pyramid_fused_ms_per_step": fused["ms_per_step"], "per_op_ms_per_step": round(ms_step, 4),
        "cfg4": cfg4,
        "per_op": [{k: r[k] for k in ("op", "level", "crop", "n", "ms", "ms_p10", "ms_p90", "gbs")} |
                   {"frac": round(r["gbs"] / peak, 4) if r["gbs"] else None} for r in per_op if r["n"]],
    }

    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "rois_per_step_per_gpu": total_rois, "l2": "inputs larger than L2 (feature maps + grads are %d MB per step; no explicit flush)" %
                   ((sum(t.numel() for t in images.values()) + sum(op["d_grads"].numel() for op in ops)) * 4 // 2 ** 20), "sharding": "one batch per GPU, no collective",
                   "launch": (("CUDA graph replay of the %d C-ABI calls" % (2 * len(ops))) if graph is not None else "eager C-ABI calls") +
                             ("; independent op nodes of each phase on %d streams (forward phase joins before the backward phase)" % args.op_streams
                              if args.op_streams > 1 else "; one stream")},
        "ms_per_step_by_rank": ms_by_rank,
        "hbm_gbs_step": round(step_gbs, 1), "hbm_frac_step": round(step_gbs / peak, 4),
        "roofline": roofline,
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "ms_per_step": round(e2e_s / e2e_steps * 1e3, 3),
                "ms_per_step_by_rank": [round(v / e2e_steps * 1e3, 2) for v in e2e_by_rank], "copy_policy": pipe_mode,
                "pcie_gbs_by_rank": [{"h2d": round(h2d * e2e_steps / v / 1e9, 1), "d2h": round(d2h * e2e_steps / v / 1e9, 1)}
                                     for v in e2e_by_rank] if not args.no_e2e else None,
                "note": "host buffers in / host results out; both copy directions run concurrently, so each rate is bytes of that "
                        "direction / whole step time (single-GPU ceilings of this pool: 55 GB/s one way, 47 GB/s each way "
                        "when both are busy, profiles/pcie_probe.py)"},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "nms3d": nms,
        "pyramid_fused": fused,
        "proposal_layer": proposal,
        "detection_layer": detection,
        "mask_targets": mask_t,
        "target_file_payloads": wire,
        "ops": per_op,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # The CPU baseline (oracle port / the reference's own binaries) runs in a CHILD process: this process -- the
        # product arm -- never maps anything under oracle/.
        del ops, images
        torch.cuda.empty_cache()
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-child", "all", "--workload", WORKLOAD_NAME],
                             stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
        try:
            line.update(json.loads(res.stdout.strip().splitlines()[-1]))
        except Exception:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                    "sample": "CPU baseline child failed: " + (res.stderr or res.stdout)[-300:]}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------
# CPU baselines (test infrastructure, never the product):
#   kind "reference": the reference's OWN compiled kernels (oracle/_ref, run by oracle/refrun without
#                     TensorFlow).  They are single-threaded; the independent op nodes of a step run
#                     concurrently on a thread pool, which is all the parallelism TF's inter-op pool
#                     could give them.
#   kind "port":      the C oracle (bit-identical to the above) with OpenMP over boxes / channels on all
#                     host cores -- a generous upper bound for a CPU implementation of the same algorithm.
# ------------------------------------------------------------------------------------------
def _cpu_engines():
    import oracle
    try:
        from oracle import refrun
        ref = refrun.load()
    except Exception:  # noqa: BLE001
        ref = None
    return oracle, ref


def cpu_step(ops, kind, threads, sample_rois):
    """One step on the host over the first `sample_rois` ROIs (spread over the ops like the full step).
    Returns (seconds for the sampled step, ROIs done, seconds of the zero-fill-only part)."""
    from concurrent.futures import ThreadPoolExecutor
    oracle, ref = _cpu_engines()
    frac = min(1.0, sample_rois / float(BATCH * ROIS_PER_IMAGE))
    sub = []
    done = 0
    for op in ops:
        n = max(int(round(op["n"] * frac)), 1) if op["n"] else 0
        sub.append((op, op["boxes"][:n], op["bidx"][:n], op["grads"][:n]))
        if op["crop"] == CROPS[0]:
            done += n
    if kind == "reference":
        fwd = [lambda o=o, b=b, i=i: ref.crop_and_resize_3d(o["image"], b, i, o["crop"]) for o, b, i, g in sub if len(b)]
        bwd = [lambda o=o, b=b, i=i, g=g: ref.crop_and_resize_3d_grad_image(g, b, i, o["shape"]) for o, b, i, g in sub]
        fill = [lambda o=o, b=b, i=i, g=g: ref.crop_and_resize_3d_grad_image(g[:0], b[:0], i[:0], o["shape"]) for o, b, i, g in sub]
        with ThreadPoolExecutor(max(1, threads)) as ex:
            t0 = time.perf_counter()
            list(ex.map(lambda f: f(), fwd))
            list(ex.map(lambda f: f(), bwd))
            t1 = time.perf_counter()
            list(ex.map(lambda f: f(), fill))
            t2 = time.perf_counter()
    else:
        t0 = time.perf_counter()
        for o, b, i, g in sub:
            if len(b):
                oracle.crop_and_resize_3d(o["image"], b, i, o["crop"], threads=threads)
        for o, b, i, g in sub:
            oracle.crop_and_resize_3d_grad_image(g, b, i, o["shape"], threads=threads)
        t1 = time.perf_counter()
        for o, b, i, g in sub:
            oracle.crop_and_resize_3d_grad_image(g[:0], b[:0], i[:0], o["shape"], threads=threads)
        t2 = time.perf_counter()
    return t1 - t0, done, t2 - t1


def cpu_baseline(ops, kind, threads, sample_rois, repeats=1):
    total = BATCH * ROIS_PER_IMAGE
    runs = []
    while len(runs) < max(repeats, 1) and sum(r[0] + r[2] for r in runs) < 15.0:      # bounded: <= ~15 s of CPU work
        runs.append(cpu_step(ops, kind, threads, sample_rois))
    t_all, done, t_fill = (statistics.mean(r[0] for r in runs), runs[0][1], statistics.mean(r[2] for r in runs))
    if kind == "reference":
        threads = min(threads, len(ops))               # one single-threaded kernel per op node: at most len(ops) run at once
    # per-ROI work scales with the ROI count; the zero-fill of the 8 grad images is paid once per step
    t_step = min(t_fill, t_all) + max(t_all - t_fill, 0.0) * total / max(done, 1)
    how = ("the reference's own compiled kernels (oracle/_ref via oracle/refrun), op nodes of the step run "
           "concurrently on %d threads" % threads) if kind == "reference" else \
          ("C oracle port, OpenMP over boxes (fwd) / channels (bwd) on %d threads" % threads)
    return {"value": round(total / t_step, 3), "unit": UNIT, "cores": threads, "kind": kind,
            "sample": "%d of %d ROIs through all 16 ops on full-size feature maps, %d pass(es) (%.1f s of CPU work); step "
                      "time = zero-fill + per-ROI time x %d; %s" % (done, total, len(runs), t_all * len(runs), total, how),
            "ms_per_step": round(t_step * 1e3, 1)}


def run_cpu_baseline_child():
    """Child process of the GPU arm: times the reference's CPU implementation of the step and prints one JSON object."""
    _, ref = _cpu_engines()
    if ref is not None:
        from oracle import refrun
        refrun.harden_process()
    ops = make_workload(seed=2002)
    nthr = host_threads()
    full = BATCH * ROIS_PER_IMAGE
    out = {}
    if ref is not None:
        out["cpu_baseline"] = cpu_baseline(ops, "reference", min(nthr, 16), full, repeats=5)
        out["cpu_baseline_reference_1thread"] = cpu_baseline(ops, "reference", 1, full, repeats=2)
        out["cpu_baseline_port_allcores"] = cpu_baseline(ops, "port", nthr, full, repeats=5)
    else:
        out["cpu_baseline"] = cpu_baseline(ops, "port", nthr, full, repeats=10)
        out["cpu_baseline_port_1thread"] = cpu_baseline(ops, "port", 1, full, repeats=2)
    print(json.dumps(out))


def host_threads():
    """Host cores available to this process (torchrun exports OMP_NUM_THREADS=1, which must not shrink the baseline)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def run_reference(args):
    """The reference's CPU implementation of the step on the box's host cores: its own compiled kernels
    (oracle/_ref) when available, else the bit-identical C port; same config / metric / unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not args.reference_child:
        # The measurement runs in a locked-down child (no file writes, no network: it executes the reference wheel's
        # machine code); this process only relays the child's JSON line, so the line reaches stdout whether stdout is
        # a pipe, a terminal or a file (RLIMIT_FSIZE = 0 in the child would turn a redirected print into EFBIG).
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--reference-child", "--gpus", str(args.gpus),
               "--steps", str(args.steps), "--warmup", str(args.warmup), "--workload", args.workload]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
        if res.returncode != 0 or not lines:
            sys.stderr.write(res.stderr[-2000:])
            raise SystemExit("reference arm failed (child exit code %d)" % res.returncode)
        print(lines[-1])
        return
    oracle, ref = _cpu_engines()
    kind = "reference" if ref is not None else "port"
    if ref is not None:                                 # this process only ever runs the reference: lock it down
        from oracle import refrun
        refrun.harden_process()
    ops = make_workload(seed=2002)
    threads = min(host_threads(), 16) if kind == "reference" else host_threads()
    total = BATCH * ROIS_PER_IMAGE
    sample = total                                  # every step is the whole cfg2 step (~2 s on the host)
    for _ in range(min(args.warmup, 1)):
        cpu_step(ops, kind, threads, 8)
    t_sum, vals, last = 0.0, [], None
    for _ in range(max(args.steps, 1)):
        last = cpu_baseline(ops, kind, threads, sample)
        vals.append(last["ms_per_step"] * 1e-3)
        t_sum += vals[-1] * sample / total
        if t_sum > 120:
            break
    t_step = statistics.mean(vals)
    value = round(total / t_step, 3)
    last["value"] = value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
        "warmup": min(args.warmup, 1), "ms_per_step": round(t_step * 1e3, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rois_per_step_per_gpu": total},
        "cpu_baseline": last,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4"])
    ap.add_argument("--op-streams", type=int, default=1,
                    help="streams the independent op nodes of each phase are spread over (1 = one stream, serial)")
    ap.add_argument("--no-graph", action="store_true", help="time eager C-ABI calls instead of CUDA-graph replays")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="kernel iteration only: skip the host-buffer leg")
    ap.add_argument("--no-cfg4", action="store_true", help="skip the extra BASELINE configs[3] measurement")
    ap.add_argument("--cpu-baseline-child", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--reference-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    set_workload(args.workload)
    if args.cpu_baseline_child:
        run_cpu_baseline_child()
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
