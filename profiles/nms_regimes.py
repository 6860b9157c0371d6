"""NMS3D timing across scan regimes (how deep the greedy scan must go before max_out boxes are kept)."""
import os, sys, statistics
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import roi3d_b200 as rb, roi3d_synth
dev = torch.device("cuda", 0)
lib = rb._lib.load()
if os.environ.get("NMS_EXPERIMENT"):
    rb.set_option("car_experiment", int(os.environ["NMS_EXPERIMENT"]))   # 8: tail masks without the kept-row skip
cases = [
    ("cfg1 6000->1000 @0.7 (SURVEY 8d boxes)", 6000, 1000, 0.7, dict()),
    ("6000->1000 @0.3", 6000, 1000, 0.3, dict()),
    ("6000->1000 @0.7 dense clusters of 64", 6000, 1000, 0.7, dict(cluster=64, jitter=0.05)),
    ("6000->6000 @0.7 (full scan)", 6000, 6000, 0.7, dict()),
    ("6000->6000 @0.5 dense clusters of 64", 6000, 6000, 0.5, dict(cluster=64, jitter=0.05)),
    ("cfg3 20000->2000 @0.7", 20000, 2000, 0.7, dict()),
    ("20000->2000 @0.7 dense clusters of 64", 20000, 2000, 0.7, dict(cluster=64, jitter=0.05)),
    ("1000->1000 @0.5", 1000, 1000, 0.5, dict()),
]
for name, n, mo, thr, kw in cases:
    vol = (256, 256, 256) if n > 6000 else (128, 128, 128)
    b, s = roi3d_synth.nms_boxes(n, vol, **kw)
    db, ds = torch.from_numpy(b).to(dev), torch.from_numpy(s).to(dev)
    wsb = lib.roi3d_nms3d_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    keep = torch.empty(mo, dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    ms = []
    for it in range(40):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rb._lib.check(lib.roi3d_nms3d(db.data_ptr(), ds.data_ptr(), n, mo, thr, keep.data_ptr(), cnt.data_ptr(),
                                      ws.data_ptr(), wsb, torch.cuda.current_stream().cuda_stream))
        e.record()
        torch.cuda.synchronize()
        if it >= 10:
            ms.append(a.elapsed_time(e))
    # the same call captured once in a CUDA graph (it is stream-ordered and allocation-free) and replayed
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        rb._lib.check(lib.roi3d_nms3d(db.data_ptr(), ds.data_ptr(), n, mo, thr, keep.data_ptr(), cnt.data_ptr(),
                                      ws.data_ptr(), wsb, torch.cuda.current_stream().cuda_stream))
    gms = []
    for it in range(40):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        e.record()
        torch.cuda.synchronize()
        if it >= 10:
            gms.append(a.elapsed_time(e))
    import time
    t0 = time.perf_counter()
    for _ in range(200):
        rb._lib.check(lib.roi3d_nms3d(db.data_ptr(), ds.data_ptr(), n, mo, thr, keep.data_ptr(), cnt.data_ptr(),
                                      ws.data_ptr(), wsb, torch.cuda.current_stream().cuda_stream))
    host_us = (time.perf_counter() - t0) / 200 * 1e6
    torch.cuda.synchronize()
    k = int(cnt.item())
    order = np.argsort(-s, kind="stable")
    rank = np.empty(n, np.int64); rank[order] = np.arange(n)
    kept = keep[:k].cpu().numpy()
    depth = int(rank[kept[-1]]) + 1 if k else 0
    if k < mo:
        depth = n
    print("%-44s kept %5d  scan depth %6d  eager %.4f ms  graph %.4f ms  host enqueue %.1f us" %
          (name, k, depth, statistics.median(ms), statistics.median(gms), host_us), flush=True)
