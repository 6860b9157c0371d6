import torch, time
dev=torch.device('cuda',0)
n=1<<28  # 1 GiB floats? 256M floats = 1 GiB
h1=torch.empty(n,dtype=torch.float32).pin_memory(); h2=torch.empty(n,dtype=torch.float32).pin_memory()
d1=torch.empty(n,dtype=torch.float32,device=dev); d2=torch.empty(n,dtype=torch.float32,device=dev)
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
def t(fn,reps=3):
    fn(); torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter()-t0)/reps
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1,non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
def both(): h2d(); d2h()
gb=n*4/1e9
print("H2D %.1f GB/s  D2H %.1f GB/s  both: %.1f GB/s each direction"%(gb/t(h2d),gb/t(d2h),gb/t(both)))
