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
# write-only with incompressible data: the source is a 32 MiB random block that stays in L2 and is
# broadcast over the 4 GiB destination, so DRAM sees (almost) only writes
xs = torch.randint(0, 255, (32 * 1024**2,), dtype=torch.uint8, device='cuda').view(torch.float32)
yv = y.view(torch.float32).view(128, -1)
ms=t(lambda: yv.copy_(xs.expand(128, -1))); print("write-only random data 4GiB: %.3f ms %.0f GB/s"%(ms, y.numel()/ms/1e6))
xr = torch.randint(0, 255, (4*1024**3,), dtype=torch.uint8, device='cuda')
ms=t(lambda: y.copy_(xr)); print("copy random 4GiB: %.3f ms %.0f GB/s (r+w)"%(ms, 2*x.numel()/ms/1e6))
import ctypes, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from upnerf_b200 import _lib as L
lib = L.lib()
def fillp():
    L.check(lib.upnerf_fill_pattern(ctypes.c_void_p(y.data_ptr()), ctypes.c_int64(y.numel()), ctypes.c_uint32(7), L.stream_ptr()), "fill")
ms=t(fillp); print("write-only hash pattern (st.global.cs v4) 4GiB: %.3f ms %.0f GB/s"%(ms, y.numel()/ms/1e6))
for ld in (256, 64, 320):
    rows = (y.numel() // 2 // ld) // 128 * 128
    def probe():
        L.check(lib.upnerf_tma_store_probe(ctypes.c_void_p(y.data_ptr()), ctypes.c_int64(rows), ctypes.c_int64(ld), ctypes.c_int(4), L.stream_ptr()), "probe")
    ms=t(probe); print("TMA store probe 128x64 bf16 boxes, ld=%d: %.3f ms %.0f GB/s"%(ld, ms, rows*ld*2/ms/1e6))
