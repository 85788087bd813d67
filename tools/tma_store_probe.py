"""TMA-store probe: rate of 128 x 64 bf16 box stores (the fused trunk's activation boxes) per SM as a
function of the number of bulk stores in flight.  depth 1 = issue, wait for the shared-memory read, repeat:
16 KB / (cycles per store) exposes the latency a kernel pays when a box must be read out before it
can be overwritten."""
import ctypes
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from upnerf_b200 import _lib as L

lib = L.lib()


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


y = torch.empty(2 * 1024**3, dtype=torch.uint8, device="cuda")
for ld in (256, 64):
    rows = (y.numel() // 2 // ld) // 128 * 128
    for depth in (1, 2, 3, 4):
        def probe():
            L.check(lib.upnerf_tma_store_probe(ctypes.c_void_p(y.data_ptr()), ctypes.c_int64(rows), ctypes.c_int64(ld),
                                               ctypes.c_int(depth), L.stream_ptr()), "probe")
        ms = t(probe)
        gbs = rows * ld * 2 / ms / 1e6
        bpc = gbs * 1e9 / 148 / 1.965e9
        print("ld=%d depth=%d: %.3f ms %.0f GB/s = %.1f B/clk/SM = %.0f cycles per 16 KB store" % (ld, depth, ms, gbs, bpc, 16384 / bpc))

# L2-resident footprint, many passes inside one launch: the TMA engine's own store rate (no HBM limit)
for mb in (16, 32):
    rows = (mb * 1024 * 1024 // 2 // 256) // 128 * 128
    reps = 63
    def probe2():
        L.check(lib.upnerf_tma_store_probe(ctypes.c_void_p(y.data_ptr()), ctypes.c_int64(rows), ctypes.c_int64(256),
                                           ctypes.c_int(4 + (reps << 8)), L.stream_ptr()), "probe")
    ms = t(probe2)
    gbs = rows * 256 * 2 * (reps + 1) / ms / 1e6
    print("L2-resident %d MB x %d passes, ld=256 depth=4: %.3f ms %.0f GB/s = %.1f B/clk/SM" % (mb, reps + 1, ms, gbs, gbs * 1e9 / 148 / 1.965e9))
