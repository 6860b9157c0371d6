"""Stress run (not collected by pytest): many random NMS3D problems back to back on one stream, compared with the CPU
oracle -- shakes out ordering bugs between the chained kernels (head/tail phases, programmatic dependent launch).
usage: python tests/stress_nms.py [problems]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle          # noqa: E402
import roi3d_b200 as rb  # noqa: E402
import roi3d_synth     # noqa: E402

dev = torch.device("cuda", 0)
rng = np.random.default_rng(12345)
problems = int(sys.argv[1]) if len(sys.argv) > 1 else 300
bad = 0
pending = []
for it in range(problems):
    n = int(rng.choice([1, 7, 33, 500, 513, 1025, 2049, 3000, 6000, 9000, 13000]))
    max_out = int(rng.choice([1, 10, 100, 300, 1000, 2000, n]))
    thr = float(rng.choice([0.0, 0.3, 0.5, 0.7, 1.0]))
    kw = dict(cluster=int(rng.choice([1, 8, 64])), jitter=float(rng.choice([0.02, 0.15])))
    boxes, scores = roi3d_synth.nms_boxes(n, (128, 128, 128), seed=1000 + it, **kw)
    if it % 7 == 0:
        scores[rng.integers(0, n, max(1, n // 50))] = np.nan
    # enqueue several problems before looking at any result, so that launches of consecutive calls overlap
    lib = rb._lib.load()
    db, ds = torch.from_numpy(boxes).to(dev), torch.from_numpy(scores).to(dev)
    wsb = lib.roi3d_nms3d_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    keep = torch.empty(max(max_out, 1), dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    rb._lib.check(lib.roi3d_nms3d(db.data_ptr(), ds.data_ptr(), n, max_out, thr, keep.data_ptr(), cnt.data_ptr(), ws.data_ptr(), wsb,
                                  torch.cuda.current_stream().cuda_stream))
    pending.append((it, boxes, scores, max_out, thr, keep, cnt, db, ds, ws))
    if len(pending) == 8 or it == problems - 1:
        torch.cuda.synchronize()
        for (i, b, s, mo, t, k, c, *_rest) in pending:
            ref = oracle.non_max_suppression_3d(b, s, mo, t)
            got = k[: int(c.item())].cpu().numpy()
            if not np.array_equal(got, ref):
                bad += 1
                print("MISMATCH problem", i, len(b), mo, t, len(got), len(ref), flush=True)
        pending = []
print("stress_nms: %d problems, %d mismatches" % (problems, bad))
sys.exit(1 if bad else 0)
