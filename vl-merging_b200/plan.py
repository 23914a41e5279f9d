"""plan.py — host-side logic of the three merge methods: which expert tensors feed which merged tensor,
with which coefficients.  Pure Python (no torch, no CUDA); the result is a list of MergeOp that
merge.py turns into ONE vlm_merge_plan launch (+ the RegMean linear problems).

Reference semantics reproduced (src/vilt/modules/vilt_module.py):
  * key templates `layer_orders`                       :376-384 = :543-551 = :650-658
  * expert selection per layer                         :397-404 (regmean), :555-567 / :666-678
  * ratios                                             :569-584 (interpolation), :680-694 (arithmetic)
  * a missing expert key => the already-shared tensor
    state_dict[merged_key] is passed through           :597-599 and the same `else: ...; break` elsewhere
  * a missing Gram => that modality is skipped          :419-420, :470-471
  * everything outside the blocks and gamma_* copied    :370-373 / :537-540 / :644-647
  * hard-coded 12 layers                               :395, :553, :665  (num_layers=12 by default here;
                                                       ViT-L callers pass 24)
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

WSUM, SEQ_LERP, MEAN = 0, 1, 2

# (expert key, merged key, sub-names) in the reference's order
_TEMPLATES = (
    ("attn.{m}.qkv.weight", "attn.qkv.weight", (None,)),
    ("attn.{m}.proj.{n}", "attn.proj.{n}", ("weight", "bias")),
    ("attn.{m}.{n}", "attn.{n}", ("q_bias", "v_bias")),
    ("mlp.{m}.fc1.{n}", "mlp.fc1.{n}", ("weight", "bias")),
    ("mlp.{m}.fc2.{n}", "mlp.fc2.{n}", ("weight", "bias")),
    ("norm1.{m}.{n}", "norm1.{n}", ("weight", "bias")),
    ("norm2.{m}.{n}", "norm2.{n}", ("weight", "bias")),
)
_LINEAR_SUFFIXES = ("attn.qkv.weight", "attn.proj.weight", "mlp.fc1.weight", "mlp.fc2.weight")


@dataclass
class MergeOp:
    dst: str                                   # merged (ufo) key
    mode: int = WSUM                           # WSUM / SEQ_LERP / MEAN
    srcs: List[str] = field(default_factory=list)   # expert keys, in accumulation order
    coefs: List[float] = field(default_factory=list)
    central: bool = False                      # SEQ_LERP: source 0 is central_weight[dst]
    passthrough: Optional[str] = None          # set => new[dst] = state_dict[passthrough], nothing computed
    # RegMean linear problem: (expert weight key, Gram key) per contributing modality
    regmean: Optional[List[Tuple[str, str]]] = None


def is_passthrough_key(k):
    return "transformer.blocks." not in k or "gamma" in k


def layer_targets(i):
    """(expert key with {m} open, merged key) for layer i, in the reference's insertion order."""
    p = f"transformer.blocks.{i}."
    for src, dst, subs in _TEMPLATES:
        for n in subs:
            yield (p + src).replace("{n}", n or ""), (p + dst).replace("{n}", n or "")


def modalities_interp(i, cfg):
    if i < cfg["vlffn_start_layer_index"]:
        return ["v", "l"]
    if cfg["only_activate_used_experts"]:
        for task, mods in (("irtr", ["v", "l"]), ("vqa", ["vl"]), ("nlvr2", ["vl"])):
            if cfg["loss_names"][task] > 0:
                return mods
        # the reference dies here with TypeError: object of type 'NoneType' has no len() (:569)
        raise TypeError("only_activate_used_experts=True needs one of loss_names irtr / vqa / nlvr2 > 0")
    return ["v", "l", "vl"]


def modalities_regmean(i, cfg):
    if i < cfg["vlffn_start_layer_index"]:
        return ["v", "l"]
    if cfg["loss_names"]["irtr"] > 0:
        return ["v", "l"]
    if cfg["loss_names"]["vqa"] > 0:
        return ["vl"]
    return ["v", "l", "vl"]


def _present_or_passthrough(src, dst, mods, keys):
    """Expert keys present for every modality, or (None, dst) when the reference would `break` to the
    already-shared tensor (raises KeyError like the reference when that one is missing too)."""
    names = [src.replace("{m}", m) for m in mods]
    if all(n in keys for n in names):
        return names, None
    if dst not in keys:
        raise KeyError(dst)
    return None, dst


def plan_merge_weights(keys, cfg, num_layers=12):
    alpha = cfg["merge_ratio"]
    ops = []
    for i in range(num_layers):
        mods = modalities_interp(i, cfg)
        if len(mods) == 1:
            ratios = {mods[0]: 1.0}
        elif len(mods) == 3:
            ratios = {"v": (2 / 3) * alpha, "l": (2 / 3) * (1 - alpha), "vl": 1 / 3}
        else:
            ratios = {"v": alpha, "l": 1 - alpha}
        for src, dst in layer_targets(i):
            names, through = _present_or_passthrough(src, dst, mods, keys)
            if through:
                ops.append(MergeOp(dst, passthrough=through))
            else:
                ops.append(MergeOp(dst, WSUM, names, [ratios[m] for m in mods]))
    return ops


def plan_sum_task_vectors(keys, central_keys, cfg, num_layers=12):
    lam = cfg["sum_lambda"]
    ops = []
    for i in range(num_layers):
        mods = modalities_interp(i, cfg)
        coef = 1.0 if len(mods) == 1 else lam
        for src, dst in layer_targets(i):
            if dst not in central_keys:
                raise KeyError(dst)  # central_weight[later_name], :700
            names, through = _present_or_passthrough(src, dst, mods, keys)
            if through:
                ops.append(MergeOp(dst, passthrough=through))
            else:
                ops.append(MergeOp(dst, SEQ_LERP, names, [coef] * len(mods), central=True))
    return ops


def plan_regmean(keys, gram_keys, cfg, num_layers=12):
    ops = []
    for i in range(num_layers):
        mods = modalities_regmean(i, cfg)
        for src, dst in layer_targets(i):
            names, through = _present_or_passthrough(src, dst, mods, keys)
            if through:
                ops.append(MergeOp(dst, passthrough=through))
            elif dst.endswith(_LINEAR_SUFFIXES):
                pairs = []
                for name in names:
                    gram = name[: -len(".qkv.weight")] if name.endswith(".qkv.weight") else name[: -len(".weight")]
                    if gram in gram_keys:
                        pairs.append((name, gram))
                if not pairs:
                    # the reference silently stores the integer 0 here (:429-430 with later_weight = 0),
                    # which then breaks load_state_dict; fail at the cause instead
                    raise KeyError(f"no Gram matrix for any expert of {dst}")
                ops.append(MergeOp(dst, regmean=pairs))
            else:
                ops.append(MergeOp(dst, MEAN, names, [1.0] * len(names)))
    return ops
