"""Development probe: pure-write, pure-read and copy bandwidth of the device (torch ops, CUDA events)."""
import torch
n = 2 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda")
b = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: a.zero_()); print(f"memset 2 GiB: {ms:.3f} ms  {n/ms/1e6:.0f} GB/s written")
ms = t(lambda: a.fill_(7)); print(f"fill   2 GiB: {ms:.3f} ms  {n/ms/1e6:.0f} GB/s written")
ms = t(lambda: b.copy_(a)); print(f"copy   2 GiB: {ms:.3f} ms  {2*n/ms/1e6:.0f} GB/s read+written")
a32 = a.view(torch.int32)
ms = t(lambda: a32.sum()); print(f"read   2 GiB: {ms:.3f} ms  {n/ms/1e6:.0f} GB/s read")
