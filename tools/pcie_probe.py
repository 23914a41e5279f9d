#!/usr/bin/env python
"""Host<->device copy ceilings of the box (pinned memory): H2D alone, D2H alone, both at once — the bound of the
end-to-end merge (680 MB in, 340 MB out for VLMo-base)."""
import time

import torch

dev = torch.device("cuda", 0)
h_in = torch.empty(680 << 20, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(340 << 20, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(680 << 20, dtype=torch.uint8, device=dev)
d_out = torch.empty(340 << 20, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def run(h2d, d2h, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best


a, b, c = run(True, False), run(False, True), run(True, True)
print(f"H2D 680 MiB: {a * 1e3:.2f} ms = {(680 << 20) / a * 1e-9:.1f} GB/s")
print(f"D2H 340 MiB: {b * 1e3:.2f} ms = {(340 << 20) / b * 1e-9:.1f} GB/s")
print(f"both at once: {c * 1e3:.2f} ms (sum of the two alone: {(a + b) * 1e3:.2f} ms) = {(1020 << 20) / c * 1e-9:.1f} GB/s aggregate")
