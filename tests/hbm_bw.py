import torch
dev = torch.device("cuda")
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device=dev)   # 4 GB
b = torch.empty(n, dtype=torch.float32, device=dev)
def t(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: a.zero_());            print(f"write-only (memset 4 GB): {4.295/ms*1e3:.0f} GB/s")
ms = t(lambda: a.fill_(1.5));         print(f"write-only (fill kernel 4 GB): {4.295/ms*1e3:.0f} GB/s")
ms = t(lambda: torch.sum(a));         print(f"read-only (sum 4 GB): {4.295/ms*1e3:.0f} GB/s")
ms = t(lambda: b.copy_(a));           print(f"copy (4 GB read + 4 GB write): {8.59/ms*1e3:.0f} GB/s")
