import sys, statistics
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import roi3d_b200 as rb, roi3d_synth
dev=torch.device('cuda',0)
vol=(128,128,128)
shape=roi3d_synth.level_shape(vol,2,batch=1,channels=256)
n=128
boxes=roi3d_synth.rois(n,vol,seed=5); tb=torch.from_numpy(boxes).to(dev); ti=torch.zeros(n,dtype=torch.int32,device=dev)
def timeit(fn,reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ev=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a,b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a,b in ev)
for p in (14,20,28):
    g=torch.randn((n,p,p,p,256),device=dev)
    bb=roi3d_synth.car_algorithmic_bytes(boxes,shape,(p,p,p),True)
    for V in (1,2):
        for tgt in (8,16,32):
            rb.set_option("car_lanes_v",V); rb.set_option("car_ctas_per_sm_target",tgt)
            t=timeit(lambda: rb.crop_and_resize_3d_grad_image(g,tb,ti,shape))
            print("pool %d V%d tgt %2d: bwd %.4f ms %5.0f GB/s"%(p,V,tgt,t,bb/t/1e6),flush=True)
    del g
