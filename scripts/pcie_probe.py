import torch, time
def bw(fn, nbytes, n=5):
    fn(); torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return nbytes*n/(time.perf_counter()-t)/1e9
N=128; R,C=3508,2480
h=torch.empty((N,R,C),dtype=torch.uint8).pin_memory(); d=torch.empty((N,R,C),dtype=torch.uint8,device='cuda')
print('H2D contiguous GB/s', bw(lambda: d.copy_(h,non_blocking=True), h.numel()))
print('D2H contiguous GB/s', bw(lambda: h.copy_(d,non_blocking=True), h.numel()))
ho=torch.empty((N,R-1,C-1),dtype=torch.uint8).pin_memory(); do=torch.empty((N,R-1,C),dtype=torch.uint8,device='cuda')
print('D2H 2D (odd width) GB/s', bw(lambda: ho.copy_(do[:,:,:C-1],non_blocking=True), ho.numel()))
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
h2=torch.empty((N,R,C),dtype=torch.uint8).pin_memory(); d2=torch.empty((N,R,C),dtype=torch.uint8,device='cuda')
print('bidirectional total GB/s', bw(both, 2*h.numel()))
