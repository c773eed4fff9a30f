#!/usr/bin/env python
"""GPU probe for the tcgen05 pointwise kernel: compares ctn_debug_pointwise(tf32x3 / tf32) against the FFMA kernel and,
for structured operands, prints which (k, t) element each output actually received -- used to pin the UMMA descriptor
encodings on hardware.  Run under gpurun; wraps nothing dangerous (each case is bounded)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dnn-based_source_separation_b200"))
from ctn_b200 import _native as N

dev = torch.device("cuda", 0)


def run(A, W, math, epi=0, bias=None, slope=None, dbg=None, frames=None):
    B, K, pitch = A.shape
    M = W.shape[0]
    frames = frames or pitch
    D = torch.full((B, M, pitch), float("nan"), device=dev)
    ws = torch.empty(8 * M * max(K, 32) * 4 + (1 << 20), dtype=torch.uint8, device=dev)
    stats = torch.zeros(B, 2, dtype=torch.float64, device=dev)
    dbg_arr = (C.c_uint32 * 4)(*dbg) if dbg else None
    rc = N.ctn_debug_pointwise(A.data_ptr(), W.data_ptr(), D.data_ptr(), B, M, K, frames, pitch, N.ptr(bias), N.ptr(slope),
                               stats.data_ptr(), epi, math, dbg_arr, ws.data_ptr(), ws.numel(), N.stream_ptr(dev))
    torch.cuda.synchronize()
    return rc, D, stats


def structured(variant_name, dbg):
    # A[k][t] = t + 128*k (exact in tf32 for < 2048 -> K=16), W = identity(16x16 padded): D[n][t] should equal A[n][t]
    K, pitch = 16, 128
    A = (torch.arange(pitch, device=dev).float()[None, :] + 128.0 * torch.arange(K, device=dev).float()[:, None])[None].contiguous()
    W = torch.eye(16, device=dev)
    rc, D, _ = run(A, W, N.MATH_TF32, dbg=dbg)
    ok = rc == 0 and torch.equal(D, A)
    print(f"[structured {variant_name}] rc={rc} exact={ok}")
    if rc == 0 and not ok:
        d = D[0].cpu()
        for n in (0, 1, 7, 8, 15):
            row = d[n]
            ks = (row // 128).int().tolist()
            ts = (row % 128).int().tolist()
            print(f"   n={n}: t=0..7 -> got (k,t)=", list(zip(ks[:8], ts[:8])), " t=32..35 ->", list(zip(ks[32:36], ts[32:36])))
    return ok


def randomized(math, name, shapes):
    worst = 0.0
    for (B, M, K, pitch, frames) in shapes:
        g = torch.Generator(device="cpu").manual_seed(M * 7 + K)
        A = torch.randn(B, K, pitch, generator=g).to(dev)
        W = (torch.randn(M, K, generator=g) / K ** 0.5).to(dev)
        bias = torch.randn(M, generator=g).to(dev)
        slope = torch.tensor([0.25], device=dev)
        for epi in (0, 2):
            rc0, D0, s0 = run(A, W, N.MATH_FP32, epi, bias, slope, frames=frames)
            rc1, D1, s1 = run(A, W, math, epi, bias, slope, frames=frames)
            err = float((D0 - D1).abs().max()) if rc0 == 0 and rc1 == 0 else float("nan")
            ref = A.double().transpose(1, 2) @ W.double().t()
            serr = float((s0 - s1).abs().max() / (s0.abs().max() + 1e-30))
            print(f"[{name}] B={B} M={M} K={K} pitch={pitch} frames={frames} epi={epi}: rc=({rc0},{rc1}) max|simt-umma|={err:.3e} "
                  f"stats rel diff={serr:.2e} nan={int(torch.isnan(D1).sum())}")
            worst = max(worst, err if err == err else 1e9)
    return worst


if __name__ == "__main__":
    print("has_tcgen05", N.ctn_has_tcgen05())
    ok = structured("default", None)
    shapes = [(1, 16, 32, 128, 128), (2, 128, 128, 256, 200), (2, 512, 128, 512, 512), (3, 256, 512, 384, 383), (1, 24, 12, 256, 157),
              (2, 1024, 128, 256, 256), (1, 10, 24, 128, 100)]
    w3 = randomized(N.MATH_TF32X3, "tf32x3", shapes)
    w1 = randomized(N.MATH_TF32, "tf32", shapes[:3])
    print(f"WORST tf32x3={w3:.3e} tf32={w1:.3e}")
