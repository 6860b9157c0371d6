"""Profiling driver: one cfg2 step restricted to the level that receives ROIs (P2) + NMS3D @6k.
Run plain first, then under ncu (see profiles/README.md).  Prints nothing that is a bench value."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import roi3d_b200 as rb   # noqa: E402
import roi3d_synth        # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
vol, B = (128, 128, 128), 2
routed = roi3d_synth.pyramid_rois(128, B, vol, seed=2002)
boxes, bidx, _ = routed[2]
shape = roi3d_synth.level_shape(vol, 2, batch=B)
torch.manual_seed(0)
image = torch.randn(shape, device=dev)
tb, ti = torch.from_numpy(boxes).to(dev), torch.from_numpy(bidx).to(dev)
nb, ns = roi3d_synth.nms_boxes(6000, vol)
dnb, dns = torch.from_numpy(nb).to(dev), torch.from_numpy(ns).to(dev)
grads = {c: torch.randn((len(boxes), c, c, c, shape[4]), device=dev) for c in (7, 14)}
for _ in range(iters):
    for c in (7, 14):
        out = rb.crop_and_resize_3d(image, tb, ti, (c, c, c))
        gi = rb.crop_and_resize_3d_grad_image(grads[c], tb, ti, shape)
    keep = rb.non_max_suppression_3d(dnb, dns, 1000, 0.7)
torch.cuda.synchronize()
print("ok", tuple(out.shape), tuple(gi.shape), len(keep))
