#!/usr/bin/env python
"""Host-side timeline of one end-to-end merge_weights call (pinned host checkpoint in, host tensors out)."""
import sys
import time

import torch

sys.path.insert(0, ".")
import vl_merging_b200 as vlm  # noqa: E402

cfg = vlm.vlmo_config("base")
with torch.device("cuda"):
    model = vlm.VLMo(cfg)
vlm.init_synthetic_(model, seed=1)
sd = {k: v.detach() for k, v in model.state_dict().items()}
host_sd = {k: (v.cpu().pin_memory() if "transformer.blocks" in k else v.cpu()) for k, v in sd.items()}
mcfg = dict(vlffn_start_layer_index=10, only_activate_used_experts=True, merge_ratio=0.5, loss_names={"irtr": 1, "vqa": 0, "nlvr2": 0})
vlm.merge_weights(host_sd, mcfg)
for rep in range(2):
    torch.cuda.synchronize()
    stats = {"trace": []}
    t0 = time.perf_counter()
    out = vlm.merge_weights(host_sd, mcfg, stats=stats)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"run {rep}: total {1e3 * (t1 - t0):.2f} ms")
    for what, t in stats["trace"]:
        print(f"   {1e3 * (t - t0):7.2f} ms  {what}")
    del out
