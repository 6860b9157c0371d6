import sys, statistics
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import roi3d_b200 as rb, roi3d_synth
dev=torch.device('cuda',0); lib=rb._lib.load()
def t(n,mo,thr,var):
    rb.set_option("nms_sort_variant",var)
    vol=(256,256,256) if n>6000 else (128,128,128)
    b,s=roi3d_synth.nms_boxes(n,vol)
    db,ds=torch.from_numpy(b).to(dev),torch.from_numpy(s).to(dev)
    wsb=lib.roi3d_nms3d_workspace_bytes(n); ws=torch.empty(wsb,dtype=torch.uint8,device=dev)
    keep=torch.empty(mo,dtype=torch.int32,device=dev); cnt=torch.zeros(1,dtype=torch.int32,device=dev)
    ms=[]
    for it in range(30):
        a,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record()
        rb._lib.check(lib.roi3d_nms3d(db.data_ptr(),ds.data_ptr(),n,mo,thr,keep.data_ptr(),cnt.data_ptr(),ws.data_ptr(),wsb,torch.cuda.current_stream().cuda_stream))
        e.record(); torch.cuda.synchronize()
        if it>=8: ms.append(a.elapsed_time(e))
    return statistics.median(ms), int(cnt.item())
for n in (1000,2000,4000,6000,8000,12000,20000,50000,100000):
    mo=max(1,round(n/6))
    r1=t(n,mo,0.7,1) if n<=50000 else (float('nan'),0)
    r2=t(n,mo,0.7,2)
    print("n %6d max_out %5d  rank-sort %.4f ms  bucketed %.4f ms  kept %d/%d"%(n,mo,r1[0],r2[0],r1[1],r2[1]),flush=True)
