import torch, time
n=839619
h=torch.empty(n,dtype=torch.float64).pin_memory(); h2=torch.empty(n,dtype=torch.float64).pin_memory()
d=torch.empty(n,dtype=torch.float64,device='cuda'); d2=torch.empty(n,dtype=torch.float64,device='cuda')
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
def t(fn,reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps*1e6
print("H2D 6.7MB us", t(lambda: d.copy_(h,non_blocking=True)), "GB/s", n*8/t(lambda: d.copy_(h,non_blocking=True))/1e3)
print("D2H 6.7MB us", t(lambda: h2.copy_(d2,non_blocking=True)))
def both():
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
print("H2D || D2H us", t(both))
def sync_each():
    d.copy_(h,non_blocking=True); torch.cuda.current_stream().synchronize()
print("H2D + sync us", t(sync_each))
