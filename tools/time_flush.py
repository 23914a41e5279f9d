"""Development helper: host cost vs GPU time of one grouped launch (48 text-tower Grams)."""
import sys
import time

import torch

sys.path.insert(0, ".")
import vl_merging_b200 as vlm  # noqa: E402

cache = vlm.GramCache(defer_bytes=128 << 20, max_pending_bytes=8 << 30)
xs = [torch.randn(2560, 768, device="cuda") for _ in range(36)] + [torch.randn(2560, 3072, device="cuda") for _ in range(12)]
for it in range(3):
    for i, x in enumerate(xs):
        cache.accumulate(f"g{i}", x)
    cache.flush()
torch.cuda.synchronize()
host, gpu = [], []
for it in range(10):
    for i, x in enumerate(xs):
        cache.accumulate(f"g{i}", x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    cache.flush()
    b.record()
    host.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    gpu.append(a.elapsed_time(b))
print(f"flush of 48 problems: host {1e3 * sum(host) / len(host):.3f} ms, event-to-event {sum(gpu) / len(gpu):.3f} ms")
t0 = time.perf_counter()
for it in range(200):
    cache.accumulate("single", xs[0][:64])
cache.flush()
torch.cuda.synchronize()
x = xs[36]
cache2 = vlm.GramCache()
torch.cuda.synchronize()
t0 = time.perf_counter()
for it in range(200):
    cache2.accumulate("g", x)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"immediate accumulate: host {(t1 - t0) / 200 * 1e6:.1f} us per call (python + ctypes + 2 tensor maps + launch)")
