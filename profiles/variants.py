import sys, os, ctypes, statistics
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import roi3d_b200 as rb, roi3d_synth
dev=torch.device('cuda',0)
vol,B=(128,128,128),2
routed=roi3d_synth.pyramid_rois(128,B,vol,seed=2002)
boxes,bidx,_=routed[2]
shape=roi3d_synth.level_shape(vol,2,batch=B)
torch.manual_seed(0)
image=torch.randn(shape,device=dev)
tb,ti=torch.from_numpy(boxes).to(dev),torch.from_numpy(bidx).to(dev)
def timeit(fn,reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a,b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a,b in ev)
for c in (7,14):
    g=torch.randn((len(boxes),c,c,c,shape[4]),device=dev)
    fb=roi3d_synth.car_algorithmic_bytes(boxes,shape,(c,c,c),False); bb=roi3d_synth.car_algorithmic_bytes(boxes,shape,(c,c,c),True)
    for name,opts in (('direct',dict(car_fwd_variant=1,car_bwd_variant=1,car_lanes_v=0)),('plane V1',dict(car_fwd_variant=2,car_bwd_variant=2,car_lanes_v=1)),('plane V2',dict(car_fwd_variant=2,car_bwd_variant=2,car_lanes_v=2))):
        for k,v in opts.items(): rb.set_option(k,v)
        tf=timeit(lambda: rb.crop_and_resize_3d(image,tb,ti,(c,c,c)))
        tbw=timeit(lambda: rb.crop_and_resize_3d_grad_image(g,tb,ti,shape))
        print('crop %2d %-9s fwd %.4f ms %6.0f GB/s | bwd %.4f ms %6.0f GB/s'%(c,name,tf,fb/tf/1e6,tbw,bb/tbw/1e6))
