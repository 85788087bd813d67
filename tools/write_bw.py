import torch
x=torch.empty(4*1024**3,dtype=torch.uint8,device='cuda')
y=torch.empty(4*1024**3,dtype=torch.uint8,device='cuda')
def t(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e)/n
ms=t(lambda: x.zero_()); print("memset 4GiB: %.3f ms %.0f GB/s"%(ms, x.numel()/ms/1e6))
xf=x.view(torch.float32)
ms=t(lambda: xf.fill_(1.5)); print("fill 4GiB: %.3f ms %.0f GB/s"%(ms, x.numel()/ms/1e6))
ms=t(lambda: y.copy_(x)); print("copy 4GiB: %.3f ms %.0f GB/s (r+w)"%(ms, 2*x.numel()/ms/1e6))
ms=t(lambda: xf.sum()); print("read(sum) 4GiB: %.3f ms %.0f GB/s"%(ms, x.numel()/ms/1e6))
