"""Pinned device->host copy rate of this box for the bench's result size (the floor of the e2e leg)."""
import torch

n = 2048 * 256 * 256 * 4
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
for _ in range(3):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    h.copy_(d, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"D2H {n / 1e6:.0f} MB pinned: {ms:.2f} ms = {n / ms / 1e6:.1f} GB/s")
