import torch, time
dev = torch.device("cuda", 0)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for mb in (268, 1024):
    n = mb * 1024 * 1024 // 4
    a = torch.empty(n, device=dev); b = torch.empty(n, device=dev)
    t = timeit(lambda: a.fill_(1.0)); print(f"fill  {mb} MB: {t:.3f} ms  write {mb/1024/t*1000:.0f} GiB/s = {mb*1.048576/t:.0f} GB/s")
    t = timeit(lambda: b.copy_(a)); print(f"copy  {mb} MB: {t:.3f} ms  r+w {2*mb*1.048576/t:.0f} GB/s")
    t = timeit(lambda: a.sum()); print(f"sum   {mb} MB: {t:.3f} ms  read {mb*1.048576/t:.0f} GB/s")
    t = timeit(lambda: torch.add(a, 1.0, out=b)); print(f"add   {mb} MB: {t:.3f} ms  r+w {2*mb*1.048576/t:.0f} GB/s")
