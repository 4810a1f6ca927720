import torch, time
n = 1 << 28   # 1 GiB of float32
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: a.fill_(1.0)); print("write-only  %.1f GB/s" % (n * 4 / ms / 1e6))
ms = t(lambda: a.sum()); print("read-only   %.1f GB/s" % (n * 4 / ms / 1e6))
ms = t(lambda: b.copy_(a)); print("copy (r+w)  %.1f GB/s" % (2 * n * 4 / ms / 1e6))
ms = t(lambda: torch.cuda.memset if False else a.zero_()); print("memset      %.1f GB/s" % (n * 4 / ms / 1e6))
