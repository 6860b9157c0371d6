"""Issue order of the cfg2 step's op nodes through the host-buffer API (pinned host in, pinned host out): which order keeps
both PCIe directions busy.  Nodes: F7 / F14 (need the P2 map, 268 MB up; give 88 / 702 MB down), B7 / B14 (need 88 / 702 MB
of grads up; give 268 MB down each); the empty levels' nodes follow in every order."""
import itertools, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import roi3d_b200 as rb, roi3d_synth
import bench
dev = torch.device("cuda", 0)
ops = bench.make_workload(seed=2002)
h_images = {o["level"]: torch.from_numpy(o["image"]).pin_memory() for o in ops}
h = [{"boxes": torch.from_numpy(o["boxes"]), "bidx": torch.from_numpy(o["bidx"]), "grads": torch.from_numpy(o["grads"]).pin_memory()} for o in ops]
busy = [i for i, o in enumerate(ops) if o["n"] > 0]
idle = [i for i, o in enumerate(ops) if o["n"] == 0]
node = {}
for i in busy:
    c = ops[i]["crop"][0]
    node["F%d" % c] = ("f", i); node["B%d" % c] = ("b", i)

def step(order, upload_stream):
    with rb.deferred():
        d_img = {}
        def dmap(lv):
            if lv not in d_img:
                d_img[lv] = rb.upload(h_images[lv]) if upload_stream else h_images[lv].to(dev, non_blocking=True)
            return d_img[lv]
        outs = []
        for name in order:
            kind, i = node[name]
            if kind == "f":
                outs.append(rb.crop_and_resize_3d(dmap(ops[i]["level"]), h[i]["boxes"], h[i]["bidx"], ops[i]["crop"]))
            else:
                outs.append(rb.crop_and_resize_3d_grad_image(h[i]["grads"], h[i]["boxes"], h[i]["bidx"], ops[i]["shape"]))
        for i in idle:
            outs.append(rb.crop_and_resize_3d(dmap(ops[i]["level"]), h[i]["boxes"], h[i]["bidx"], ops[i]["crop"]))
        for i in idle:
            outs.append(rb.crop_and_resize_3d_grad_image(h[i]["grads"], h[i]["boxes"], h[i]["bidx"], ops[i]["shape"]))
    return outs

res = []
for upload_stream in (False, True):
    for order in itertools.permutations(["F7", "B7", "F14", "B14"]):
        step(order, upload_stream); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            step(order, upload_stream)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 3 * 1e3
        res.append((ms, upload_stream, order))
        print("%-22s maps via %-13s %.2f ms/step" % (" ".join(order), "upload stream" if upload_stream else "current stream", ms), flush=True)
print("best:", sorted(res)[:3])
