#!/usr/bin/env python
"""CTN_UMMA_DBG=128 python tools/umma_timeline.py pw1|pw2 : per-slab timeline (ns) of CTA 0: producer / MMA / epilogue."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dnn-based_source_separation_b200"))
from ctn_b200 import _native as N
dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "pw1"
M, K, epi = {"pw1": (512, 128, 2), "pw2": (256, 512, 0)}[which]
B, pitch, frames = 32, 4096, 3999
A = torch.randn(B, K, pitch, device=dev); W = torch.randn(M, K, device=dev) / K ** 0.5
D = torch.empty(B, M, pitch, device=dev); bias = torch.randn(M, device=dev); slope = torch.tensor([0.25], device=dev)
stats = torch.zeros(B, 2, dtype=torch.float64, device=dev); ws = torch.empty(8 * M * K * 4 + (1 << 20), dtype=torch.uint8, device=dev)
for _ in range(3):
    N.ctn_debug_pointwise(A.data_ptr(), W.data_ptr(), D.data_ptr(), B, M, K, frames, pitch, bias.data_ptr(), slope.data_ptr(), stats.data_ptr(), epi, N.MATH_TF32X3, None, ws.data_ptr(), ws.numel(), N.stream_ptr(dev))
torch.cuda.synchronize()
buf = (C.c_ulonglong * (3 * 4096))()
N.check(N.ctn_debug_timeline(buf, 3 * 4096))
k_slabs = (K + 31) // 32
t0 = buf[1]
nsl = min(3 * k_slabs + 2, 40)
print(f"{which}: k_slabs={k_slabs}; times in ns relative to the producer's first stage grant")
print("slab | prod: wait_empty_start  granted  arrived(full) | mma: wait_full_start  full_seen  issued+committed")
for q in range(nsl):
    p = [buf[q * 4 + i] - t0 for i in range(3)]
    m = [buf[4096 + q * 4 + i] - t0 for i in range(3)]
    print(f"{q:4d} | {p[0]:8d} {p[1]:8d} {p[2]:8d} | {m[0]:8d} {m[1]:8d} {m[2]:8d}")
print("item | epi: wait_tfull_start  tfull_seen  done")
for it in range(6):
    e = [buf[8192 + it * 4 + i] - t0 for i in range(3)]
    print(f"{it:4d} | {e[0]:8d} {e[1]:8d} {e[2]:8d}")
last = max(i for i in range(1024) if buf[4096 + i * 4 + 2])
print("slabs recorded", last + 1, "total ns", buf[4096 + last * 4 + 2] - t0, "=> ns/slab", (buf[4096 + last * 4 + 2] - t0) / (last + 1))
