import torch, time
n=65536*18
for mb in (1,):
    h=torch.empty(n, dtype=torch.float64).pin_memory(); h2=torch.empty(n, dtype=torch.float64).pin_memory()
    d=torch.empty(n, dtype=torch.float64, device='cuda'); d2=torch.empty(n, dtype=torch.float64, device='cuda')
    s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
    def t(fn, reps=20):
        torch.cuda.synchronize(); t0=time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps
    by=n*8
    a=t(lambda: d.copy_(h, non_blocking=True)); print('H2D %.3f ms %.1f GB/s'%(a*1e3, by/a/1e9))
    a=t(lambda: h2.copy_(d2, non_blocking=True)); print('D2H %.3f ms %.1f GB/s'%(a*1e3, by/a/1e9))
    def both():
        with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    a=t(both); print('both %.3f ms %.1f GB/s each'%(a*1e3, by/a/1e9))
    # chunked 4
    c=n//4
    def chunked():
        for k in range(4):
            with torch.cuda.stream(s1): d[k*c:(k+1)*c].copy_(h[k*c:(k+1)*c], non_blocking=True)
    a=t(chunked); print('H2D 4 chunks %.3f ms %.1f GB/s'%(a*1e3, by/a/1e9))
