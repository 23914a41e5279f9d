"""oracle.py — TEST INFRASTRUCTURE.  CPU (numpy) restatement of the reference's merge hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker.  The product (vl-merging_b200/) never does.

Pinned against the real reference: tests/golden/*.npz were produced by executing the UNMODIFIED
reference functions through oracle/ref_harness.py (script: oracle/make_golden.py), and
tests/test_oracle.py checks every function below against them (bit-exact for the fp32 merges).
The reference itself has no tests or golden vectors for this path.

Each function cites the reference lines it follows; paths are relative to /root/reference/.
State dicts are {name: np.ndarray}.
"""
from collections import defaultdict

import numpy as np

# key templates: src/vilt/modules/vilt_module.py:376-384 (= :543-551 = :650-658).
# (expert key pattern, merged key pattern, sub-names formatted into the last slot)
LAYER_ORDERS = (
    ("transformer.blocks.{i}.attn.{m}.qkv.weight", "transformer.blocks.{i}.attn.qkv.weight", (None,)),
    ("transformer.blocks.{i}.attn.{m}.proj.{n}", "transformer.blocks.{i}.attn.proj.{n}", ("weight", "bias")),
    ("transformer.blocks.{i}.attn.{m}.{n}", "transformer.blocks.{i}.attn.{n}", ("q_bias", "v_bias")),
    ("transformer.blocks.{i}.mlp.{m}.fc1.{n}", "transformer.blocks.{i}.mlp.fc1.{n}", ("weight", "bias")),
    ("transformer.blocks.{i}.mlp.{m}.fc2.{n}", "transformer.blocks.{i}.mlp.fc2.{n}", ("weight", "bias")),
    ("transformer.blocks.{i}.norm1.{m}.{n}", "transformer.blocks.{i}.norm1.{n}", ("weight", "bias")),
    ("transformer.blocks.{i}.norm2.{m}.{n}", "transformer.blocks.{i}.norm2.{n}", ("weight", "bias")),
)


def passthrough(state_dict):
    """vilt_module.py:370-373 / :537-540 / :644-647: everything outside the blocks, and gamma_*."""
    return {k: v for k, v in state_dict.items() if "transformer.blocks." not in k or "gamma" in k}


def modalities_interp(i, cfg):
    """Expert selection of merge_weights / sum_task_vectors: vilt_module.py:555-567 / :666-678."""
    if i < cfg["vlffn_start_layer_index"]:
        return ["v", "l"]
    if cfg["only_activate_used_experts"]:
        if cfg["loss_names"]["irtr"] > 0:
            return ["v", "l"]
        if cfg["loss_names"]["vqa"] > 0:
            return ["vl"]
        if cfg["loss_names"]["nlvr2"] > 0:
            return ["vl"]
        return None  # the reference then fails on len(None), :569
    return ["v", "l", "vl"]


def modalities_regmean(i, cfg):
    """Expert selection of regmean: vilt_module.py:397-404 (no only_activate_used_experts switch)."""
    if i < cfg["vlffn_start_layer_index"]:
        return ["v", "l"]
    if cfg["loss_names"]["irtr"] > 0:
        return ["v", "l"]
    if cfg["loss_names"]["vqa"] > 0:
        return ["vl"]
    return ["v", "l", "vl"]


def _targets(i):
    """(expert key with {m} left open, merged key, is_linear_weight) in the reference's insertion order."""
    for src, dst, subs in LAYER_ORDERS:
        for n in subs:
            s = src.replace("{i}", str(i))
            d = dst.replace("{i}", str(i))
            if n is not None:
                s = s.replace("{n}", n)
                d = d.replace("{n}", n)
            yield s, d


def merge_weights(state_dict, cfg, num_layers=12):
    """Interpolation merge.  vilt_module.py:533-638; `range(12)` is :553."""
    new = passthrough(state_dict)
    alpha = cfg["merge_ratio"]
    for i in range(num_layers):
        mods = modalities_interp(i, cfg)
        if len(mods) == 1:  # :569-572
            ratios = {mods[0]: 1}
        elif len(mods) == 3:  # :574-579
            ratios = {"v": (2 / 3) * alpha, "l": (2 / 3) * (1 - alpha), "vl": 1 / 3}
        else:  # :580-584
            ratios = {"v": alpha, "l": 1 - alpha}
        for src, dst in _targets(i):
            acc = 0
            for m in mods:  # :592-599 (same loop at :609-616, :626-633)
                name = src.replace("{m}", m)
                if name in state_dict:
                    acc = acc + np.float32(ratios[m]) * state_dict[name]
                else:
                    acc = state_dict[dst]
                    break
            new[dst] = acc
    return new


def sum_task_vectors(state_dict, central, cfg, num_layers=12):
    """Modality arithmetic.  vilt_module.py:640-746.  `later_weight = central_weight[name]` (:700)
    aliases the central tensor and `later_weight += ...` (:706) updates it in place, so every
    modality is subtracted from the UPDATED centre: theta <- theta + lam*(theta_m - theta)."""
    new = passthrough(state_dict)
    lam = cfg["sum_lambda"]
    for i in range(num_layers):
        mods = modalities_interp(i, cfg)
        ratios = {mods[0]: 1} if len(mods) == 1 else {m: lam for m in mods}  # :680-694
        for src, dst in _targets(i):
            acc = central[dst].copy()
            for m in mods:  # :702-709
                name = src.replace("{m}", m)
                if name in state_dict:
                    acc = acc + np.float32(ratios[m]) * (state_dict[name] - acc)
                else:
                    acc = state_dict[dst]
                    break
            new[dst] = acc
    return new


def scale_g(g, scaling):
    """vilt_module.py:388-392: scaling*G + (1 - scaling)*diag(G), fp64."""
    g = np.asarray(g, dtype=np.float64)
    return scaling * g + (1 - scaling) * np.diag(np.diag(g))


def regmean(state_dict, grams, cfg, num_layers=12):
    """RegMean.  vilt_module.py:366-531.  Linear weights W (out,in): (sum_m W_m Ghat_m)(sum_m Ghat_m)^-1
    in fp64 with an explicit inverse (:432-434, :483-484); everything else: mean over the experts
    present (:436-457, :486-529).  A modality whose Gram is missing is skipped (:419-420, :470-471)."""
    new = passthrough(state_dict)
    scaling = cfg["scaling_for_non_diag"]
    for i in range(num_layers):
        mods = modalities_regmean(i, cfg)
        for src, dst in _targets(i):
            is_linear = src.endswith("qkv.weight") or (
                src.endswith(".weight") and (".proj." in src or ".fc1." in src or ".fc2." in src))
            if is_linear:
                summed, acc = 0, 0
                for m in mods:
                    name = src.replace("{m}", m)
                    gram_name = name.replace(".qkv.weight", "") if name.endswith("qkv.weight") else name[: -len(".weight")]
                    if name in state_dict:
                        if gram_name not in grams:
                            continue
                        g = scale_g(grams[gram_name], scaling)
                        summed = summed + g
                        acc = acc + state_dict[name].astype(np.float64) @ g
                    else:
                        acc = state_dict[dst]
                        break
                if isinstance(summed, int):  # :429-430 — no Gram seen at all
                    new[dst] = acc
                else:
                    new[dst] = acc @ np.linalg.inv(summed)
            else:
                acc, count = 0, 0
                for m in mods:
                    name = src.replace("{m}", m)
                    if name in state_dict:
                        acc = acc + state_dict[name]
                        count += 1
                    else:
                        acc = state_dict[dst]
                        break
                new[dst] = acc if count == 0 else acc / np.float32(count)
    return new


# ---- Gram caching ---------------------------------------------------------------------------------

# src/cache_gram_matrices.py:264-276
ALL_KEYS_MOE = (
    "mlp.fc1", "mlp.fc1",
    "mlp.v.fc1", "mlp.l.fc1", "mlp.vl.fc1", "mlp.v.fc2", "mlp.l.fc2", "mlp.vl.fc2",
    "attn",
    "attn.v", "attn.l", "attn.vl",
    "attn.proj",
    "attn.v.proj", "attn.l.proj", "attn.vl.proj",
)
ALL_KEYS_UFO = ("mlp.fc1", "mlp.fc2", "attn.proj", "norm1", "norm2")


def is_hooked(name, use_moe=True):
    """Module selection rule, src/cache_gram_matrices.py:278-279."""
    keys = ALL_KEYS_MOE if use_moe else ALL_KEYS_UFO
    return any(name.endswith(k) for k in keys) and ".bias" not in name


def hook_gram_input(store, name, x):
    """src/cache_gram_matrices.py:246-254: X = input.reshape(-1, D) in fp64; store[name] += X^T X.
    All rows count (padded text positions included)."""
    x = np.asarray(x)
    flat = x.reshape(-1, x.shape[-1]).astype(np.float64)
    store[name] = store[name] + flat.T @ flat


def new_gram_store():
    """middle_representations, src/cache_gram_matrices.py:236."""
    return defaultdict(float)


def reference_hook_torch(store):
    """The reference hook restated with the reference's own torch calls (src/cache_gram_matrices.py:246-254):
    fp64 cast, torch.matmul, accumulate on the host.  Used by bench.py's CPU-baseline / --impl reference
    legs, which time the reference's CPU implementation of the path on the box's host cores."""
    import torch

    def hook_gram_input(module, input, output):
        if isinstance(input, tuple):
            input = input[0]
        flat = input.reshape(-1, input.shape[-1]).to(torch.float64)
        store[module.module_name] += torch.matmul(flat.T, flat).detach().cpu()

    return hook_gram_input


def irtr_recall(img_feats, txt_feats, iids, tiids):
    """src/vilt/modules/objectives.py:684-710 in numpy: scores = img @ txt.T; recall@{1,5,10} both directions.
    Returns (scores, (ir_r1, ir_r5, ir_r10, tr_r1, tr_r5, tr_r10))."""
    scores = np.asarray(img_feats) @ np.asarray(txt_feats).T
    iids, tiids = np.asarray(iids), np.asarray(tiids)

    def top10(mat):
        """Indices of the 10 largest entries of every row, best first (ties: lower index first, like a stable
        descending sort) — a partition first, so the 5,000 x 25,000 COCO-sized matrix does not need six full sorts."""
        k = min(10, mat.shape[1])
        if mat.shape[1] > 64:
            cand = np.argpartition(-mat, k - 1, axis=1)[:, :k]
        else:
            cand = np.broadcast_to(np.arange(mat.shape[1]), mat.shape).copy()
        vals = np.take_along_axis(mat, cand, axis=1)
        order = np.lexsort((cand, -vals), axis=1)
        return np.take_along_axis(cand, order, axis=1)[:, :k]

    by_image, by_caption = top10(scores), top10(scores.T)
    res = {}
    for k in (1, 5, 10):
        top = by_image[:, :k]                                              # per image: best captions (:688-690)
        res[f"tr_r{k}"] = (iids[:, None] == tiids[top]).max(axis=1).astype(np.float32).mean()
        top = by_caption[:, :k].T                                          # per caption: best images (:699-701)
        res[f"ir_r{k}"] = (tiids[None, :] == iids[top]).max(axis=0).astype(np.float32).mean()
    return scores, (res["ir_r1"], res["ir_r5"], res["ir_r10"], res["tr_r1"], res["tr_r5"], res["tr_r10"])
