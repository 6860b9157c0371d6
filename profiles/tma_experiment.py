import sys, statistics
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import roi3d_b200 as rb, roi3d_synth, oracle
dev=torch.device('cuda',0)
# parity first (small random cases)
rng=np.random.default_rng(7)
bad=0
for it in range(60):
    C=int(rng.choice([4,8,12,32,36,64,96,128,256])); B=int(rng.integers(1,3))
    H,W,D=(int(v) for v in rng.integers(1,24,3)); crop=tuple(int(v) for v in rng.integers(1,18,3)); n=int(rng.integers(1,10))
    sp=float(rng.choice([0.0,0.4,1.5]))
    image=rng.standard_normal((B,H,W,D,C),dtype=np.float32)
    boxes=(rng.random((n,6))*(1+sp)-sp/2).astype(np.float32); bidx=rng.integers(0,B,n).astype(np.int32)
    ref=oracle.crop_and_resize_3d(image,boxes,bidx,crop,"trilinear",-2.0)
    rb.set_option("car_fwd_variant",3)
    out=rb.crop_and_resize_3d(torch.from_numpy(image).to(dev),torch.from_numpy(boxes).to(dev),torch.from_numpy(bidx).to(dev),crop,extrapolation_value=-2.0).cpu().numpy()
    if not np.array_equal(out,ref):
        bad+=1; print('MISMATCH',it,C,(H,W,D),crop,n,np.abs(out-ref).max())
print('parity mismatches',bad,flush=True)
vol,B=(128,128,128),2
routed=roi3d_synth.pyramid_rois(128,B,vol,seed=2002)
def timeit(fn,reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a,b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a,b in ev)
for lvl in (2,3):
    boxes,bidx,_=routed[lvl]
    if len(boxes)==0: continue
    shape=roi3d_synth.level_shape(vol,lvl,batch=B)
    image=torch.randn(shape,device=dev)
    tb,ti=torch.from_numpy(boxes).to(dev),torch.from_numpy(bidx).to(dev)
    for c in (7,14,28):
        fb=roi3d_synth.car_algorithmic_bytes(boxes,shape,(c,c,c),False)
        for name,opts in (('planeV1',dict(car_fwd_variant=2,car_lanes_v=1)),('planeV2',dict(car_fwd_variant=2,car_lanes_v=2)),('tma',dict(car_fwd_variant=3,car_lanes_v=0))):
            for k,v in opts.items(): rb.set_option(k,v)
            for tgt in (8,16,24):
                rb.set_option("car_ctas_per_sm_target",tgt)
                tf=timeit(lambda: rb.crop_and_resize_3d(image,tb,ti,(c,c,c)))
                print('lvl %d n %d crop %2d %-8s tgt %2d fwd %.4f ms %6.0f GB/s'%(lvl,len(boxes),c,name,tgt,tf,fb/tf/1e6),flush=True)
