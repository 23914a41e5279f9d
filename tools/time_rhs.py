"""Development helper: times vlm_regmean_rhs / vlm_gram_scale_accum / vlm_spd_solve_right per layer shape."""
import sys
import time

import torch

sys.path.insert(0, ".")
import vl_merging_b200 as vlm  # noqa: E402
from vl_merging_b200 import _lib  # noqa: E402

L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
for out_f, in_f in ((2304, 768), (768, 768), (3072, 768), (768, 3072)):
    w = torch.randn(out_f, in_f, device="cuda")
    x = torch.randn(in_f * 2, in_f, device="cuda")
    g = (x.T @ x).contiguous()
    acc = torch.empty(out_f, in_f, dtype=torch.float64, device="cuda")
    s = torch.empty(in_f, in_f, dtype=torch.float64, device="cuda")
    for name, fn in (
        ("rhs", lambda: L.vlm_regmean_rhs(w.data_ptr(), out_f, in_f, in_f, g.data_ptr(), 0, in_f, 0.9, acc.data_ptr(), in_f, 0, st)),
        ("scale", lambda: L.vlm_gram_scale_accum(g.data_ptr(), 0, in_f, in_f, 0.9, s.data_ptr(), in_f, 0, st)),
    ):
        for _ in range(3):
            _lib.check(fn())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            _lib.check(fn())
        torch.cuda.synchronize()
        print(f"{name} out={out_f} in={in_f}: {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms")
    _lib.check(L.vlm_gram_scale_accum(g.data_ptr(), 0, in_f, in_f, 0.9, s.data_ptr(), in_f, 0, st))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _lib.check(L.vlm_spd_solve_right(s.data_ptr(), in_f, in_f, acc.data_ptr(), out_f, in_f, st))
    torch.cuda.synchronize()
    print(f"solve out={out_f} in={in_f}: {(time.perf_counter() - t0) * 1e3:.3f} ms")
