"""Development aid: pure-write and copy bandwidth on this GPU (context for the write-dominated hess kernel)."""
import torch
n = 90_000_000
a = torch.empty(n, dtype=torch.float64, device="cuda")
b = torch.empty(n, dtype=torch.float64, device="cuda")
def t(f, k=50):
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
ms = t(lambda: a.zero_()); print(f"memset 720 MB: {ms:.4f} ms -> {720 / ms:.0f} GB/s write-only")
ms = t(lambda: a.fill_(1.5)); print(f"fill   720 MB: {ms:.4f} ms -> {720 / ms:.0f} GB/s write-only")
ms = t(lambda: b.copy_(a)); print(f"copy 720+720 MB: {ms:.4f} ms -> {1440 / ms:.0f} GB/s read+write")
c = torch.empty(2 ** 30, dtype=torch.bfloat16, device="cuda"); d = torch.empty_like(c)
ms = t(lambda: d.copy_(c), 10); print(f"copy 2+2 GiB (the MEASURED_PEAKS recipe): {ms:.4f} ms -> {2 * c.numel() * 2 / ms / 1e6:.0f} GB/s")
