#!/usr/bin/env python
"""Times one pointwise contraction (ctn_debug_pointwise) for the paper shapes under the CTN_UMMA_* debug knobs.
Usage (under gpurun): CTN_UMMA_DBG=1 python tools/umma_time.py   -> prints ms per call per shape."""
import ctypes as C
import os
import sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dnn-based_source_separation_b200"))
from ctn_b200 import _native as N
dev = torch.device("cuda", 0)
B, pitch, frames = 32, 4096, 3999
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("CTN_UMMA"))
import sys as _s
SHAPES = (("pw1", 512, 128, 2), ("pw2", 256, 512, 0), ("head", 128, 512, 0), ("mask", 1024, 128, 0))
if len(_s.argv) > 1: SHAPES = tuple(x for x in SHAPES if x[0] in _s.argv[1:])
for (name, M, K, epi) in SHAPES:
    A = torch.randn(B, K, pitch, device=dev)
    W = torch.randn(M, K, device=dev) / K ** 0.5
    D = torch.empty(B, M, pitch, device=dev)
    bias = torch.randn(M, device=dev); slope = torch.tensor([0.25], device=dev)
    stats = torch.zeros(B, 2, dtype=torch.float64, device=dev)
    ws = torch.empty(8 * M * K * 4 + (1 << 20), dtype=torch.uint8, device=dev)
    for math, mname in ((N.MATH_TF32X3, "tf32x3"), (N.MATH_TF32, "tf32")):
        def call():
            return N.ctn_debug_pointwise(A.data_ptr(), W.data_ptr(), D.data_ptr(), B, M, K, frames, pitch, bias.data_ptr(), slope.data_ptr(),
                                         stats.data_ptr(), epi, math, None, ws.data_ptr(), ws.numel(), N.stream_ptr(dev))
        for _ in range(3):
            rc = call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * M * K * frames * B
        print(f"[{tag}] {name} {mname}: rc={rc} {ms:.3f} ms/call (incl. wimg build)  {fl/ms/1e9:.1f} TFLOP/s")
