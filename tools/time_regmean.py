#!/usr/bin/env python
"""RegMean of a VLMo-base / large checkpoint from device Grams: wall time vs the number of solve streams."""
import sys
import time

import torch

sys.path.insert(0, ".")
import vl_merging_b200 as vlm  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "base"
cfg = vlm.vlmo_config(name, attn_impl="sdpa")
with torch.device("cuda"):
    model = vlm.VLMo(cfg)
vlm.init_synthetic_(model.eval(), seed=1)
cache = vlm.GramCache(defer_bytes=128 << 20)
cache.register(model)
bs = 64 if name == "base" else 32
with torch.no_grad():
    for seed in range(2 if name == "base" else 5):   # >= 4096 text rows for the large model
        model(vlm.synthetic_batch(bs, cfg, seed=seed, device="cuda"))
cache.remove_hooks()
sd = {k: v.detach() for k, v in model.state_dict().items()}
mcfg = dict(vlffn_start_layer_index=cfg["vlffn_start_layer_index"], scaling_for_non_diag=0.9,
            loss_names={"irtr": 1, "vqa": 0, "nlvr2": 0})
L = cfg["num_layers"]
for n in (1, 8, 8, 12, 12, 16, 16, 24, 24, 8):
    torch.cuda.synchronize()
    t = time.perf_counter()
    vlm.regmean(sd, mcfg, num_layers=L, gram_matrices=cache, solve_streams=n)
    torch.cuda.synchronize()
    print(f"{name}: solve_streams={n}: {time.perf_counter() - t:.4f} s")
