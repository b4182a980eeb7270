import torch, time
n=315*1000*1000
h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device='cuda')
h2=torch.empty(31*1000*1000,dtype=torch.uint8).pin_memory(); d2=torch.empty(31*1000*1000,dtype=torch.uint8,device='cuda')
def t(f,reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps
a=t(lambda: d.copy_(h,non_blocking=True)); print("H2D 315MB one stream: %.2f ms %.1f GB/s"%(a*1e3,n/a/1e9))
s1,s2=torch.cuda.Stream(),torch.cuda.Stream(); half=n//2
def two():
    with torch.cuda.stream(s1): d[:half].copy_(h[:half],non_blocking=True)
    with torch.cuda.stream(s2): d[half:].copy_(h[half:],non_blocking=True)
a=t(two); print("H2D 315MB two streams: %.2f ms %.1f GB/s"%(a*1e3,n/a/1e9))
a=t(lambda: h2.copy_(d2,non_blocking=True)); print("D2H 31MB: %.2f ms %.1f GB/s"%(a*1e3,31e6/a/1e9))
def both():
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
a=t(both); print("H2D 315MB + D2H 31MB concurrent: %.2f ms"%(a*1e3))
