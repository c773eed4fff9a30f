"""Times ctn_bilstm_proj_fwd alone at the cfg4 shapes (CUDA events); CTN_LSTM_DBG / CTN_LSTM_STAGES are read by the library."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dnn-based_source_separation_b200"))
import torch
from ctn_b200 import _native as N

def run(NSEQ, T, Fi=64, H=128, Fo=64, reps=5):
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(0)
    k = 1.0 / H ** 0.5
    shapes = [(4 * H, Fi), (4 * H, H), (4 * H,), (4 * H,)] * 2
    w = [((torch.rand(s, generator=g) * 2 - 1) * k).to(dev) for s in shapes]
    fc = ((torch.rand(Fo, 2 * H, generator=g) * 2 - 1) / (2 * H) ** 0.5).to(dev)
    z = torch.randn(NSEQ, T, Fi, generator=g).to(dev)
    ptrs = (N._fp * 8)(*[t.data_ptr() for t in w])
    nws = N.ctn_bilstm_workspace_bytes(Fi, H, Fo)
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    P = torch.empty(2, NSEQ, T, Fo, device=dev)
    st = N.stream_ptr(dev)
    def call():
        N.check(N.ctn_bilstm_proj_fwd(z.data_ptr(), NSEQ, T, Fi, H, ptrs, fc.data_ptr(), Fo, P.data_ptr(), None, None, ws.data_ptr(), nws, st), "lstm")
    for _ in range(2):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms

def timeline():
    import ctypes as C
    buf = (C.c_ulonglong * 160)()
    N.check(N.ctn_debug_lstm_timeline(buf, 160), "tl")
    v = list(buf)
    t0 = v[0]
    names = {0: ["wait-acc/x", "x-issued", "h-wait", "h-issued"], 1: ["accfull", "ld+free", "math", "h-stored"]}
    for st in range(4):
        for c in range(5):
            row = []
            for role in range(2):
                row.append(" ".join(f"{(v[(((st * 2 + role) * 5 + c) * 4 + k)] - t0):7d}" if v[(((st * 2 + role) * 5 + c) * 4 + k)] else "      -" for k in range(4)))
            print(f"step {100 + st} chunk {c}: issuer [{row[0]}]  epilogue [{row[1]}]")


if __name__ == "__main__":
    if int(os.environ.get("CTN_LSTM_DBG", "0")) & 16:
        run(4112, 250, reps=1)
        timeline()
        sys.exit(0)
    for NSEQ, T in [(4112, 250), (4000, 257), (128 * 74, 250)]:
        ms = run(NSEQ, T)
        print(f"dbg={os.environ.get('CTN_LSTM_DBG', '0')} stages={os.environ.get('CTN_LSTM_STAGES', '-')} NSEQ={NSEQ} T={T}: {ms:.3f} ms/call = {ms * 1e3 / T:.2f} us/step", flush=True)
