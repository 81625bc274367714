import torch, time
n=1592524800
h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device='cuda')
h2=torch.empty(n*2//3,dtype=torch.uint8).pin_memory(); d2=torch.empty(n*2//3,dtype=torch.uint8,device='cuda')
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
def t(f,rep=3):
    torch.cuda.synchronize(); best=1e9
    for _ in range(rep):
        t0=time.perf_counter(); f(); torch.cuda.synchronize(); best=min(best,time.perf_counter()-t0)
    return best
a=t(lambda: d.copy_(h,non_blocking=True)); print("H2D 1.59GB: %.1f ms %.1f GB/s"%(a*1e3,n/a/1e9))
b=t(lambda: h2.copy_(d2,non_blocking=True)); print("D2H 1.06GB: %.1f ms %.1f GB/s"%(b*1e3,n*2/3/b/1e9))
def both():
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
c=t(both); print("both concurrently: %.1f ms"%(c*1e3))
# chunked 32-image copies
ch=199065600
def chunked():
    for i in range(8): d[i*ch:(i+1)*ch].copy_(h[i*ch:(i+1)*ch],non_blocking=True)
e=t(chunked); print("H2D in 8 chunks: %.1f ms"%(e*1e3))
import os; print("cpus",os.cpu_count())
