import torch, time
dev=torch.device('cuda',0)
n=1<<30
h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device=dev)
h2=torch.empty(n//2,dtype=torch.uint8).pin_memory(); d2=torch.empty(n//2,dtype=torch.uint8,device=dev)
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
def t(fn,reps=5):
    fn(); torch.cuda.synchronize(); a=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-a)/reps
print('H2D 1GiB GB/s', n/1e9/t(lambda: d.copy_(h,non_blocking=True)))
print('D2H 1GiB GB/s', n/1e9/t(lambda: h.copy_(d,non_blocking=True)))
def both():
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
x=t(both); print('duplex: H2D 1GiB + D2H 0.5GiB in', x*1e3,'ms -> H2D GB/s', n/1e9/x)
# 2D copy like the trimmed tiles
src=torch.empty((4096,4096*3),dtype=torch.uint8).pin_memory(); dst=torch.empty((4096,4096*3),dtype=torch.uint8,device=dev)
print('tile 50MB contiguous GB/s', src.numel()/1e9/t(lambda: dst.copy_(src,non_blocking=True),20))
print('tile 2D 930px wide GB/s', 4096*930*3/1e9/t(lambda: dst[:, :930*3].copy_(src[:, :930*3],non_blocking=True),20))
