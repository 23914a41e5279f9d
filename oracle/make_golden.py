"""make_golden.py — TEST INFRASTRUCTURE.  Generates tests/golden/*.npz by executing the UNMODIFIED
reference (through oracle/ref_harness.py) on seeded synthetic inputs.  Run in the build container,
where /root/reference exists:

    python oracle/make_golden.py            # all fixtures
    python oracle/make_golden.py merge      # only tests/golden/merge_small.npz

The fixtures carry inputs AND reference outputs, so the tests that consume them need neither the
reference nor this script (the GPU box has no /root/reference).
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")

LOSS_ZERO = {k: 0 for k in ("itm", "ifm", "mlm", "vqa", "nlvr2", "irtr", "mim", "image_only_mim",
                            "text_only_mlm", "img_cls", "mnc", "mld")}


def synthetic_all_moe_state_dict(hidden, mlp, num_layers, vlffn_start, gen):
    """A state_dict with the key layout of a VLMo all_moe checkpoint at toy
    width: the reference merge methods only look at key names and tensor values."""
    sd = {}

    def rnd(*shape):
        return torch.randn(*shape, generator=gen, dtype=torch.float32)

    sd["logit_scale"] = rnd(())
    sd["relative_position_bias_table"] = rnd(11, 2 * num_layers)
    sd["text_embeddings.word_embeddings.weight"] = rnd(13, hidden)
    sd["transformer.cls_token"] = rnd(1, 1, hidden)
    sd["transformer.norm.weight"] = rnd(hidden)
    sd["transformer.norm.bias"] = rnd(hidden)
    for i in range(num_layers):
        p = f"transformer.blocks.{i}."
        sd[p + "gamma_1"] = rnd(hidden)
        sd[p + "gamma_2"] = rnd(hidden)
        for m in (["v", "l"] if i < vlffn_start else ["v", "l", "vl"]):
            sd[p + f"attn.{m}.q_bias"] = rnd(hidden)
            sd[p + f"attn.{m}.v_bias"] = rnd(hidden)
            sd[p + f"attn.{m}.qkv.weight"] = rnd(3 * hidden, hidden)
            sd[p + f"attn.{m}.proj.weight"] = rnd(hidden, hidden)
            sd[p + f"attn.{m}.proj.bias"] = rnd(hidden)
            sd[p + f"norm1.{m}.weight"] = rnd(hidden)
            sd[p + f"norm1.{m}.bias"] = rnd(hidden)
            sd[p + f"mlp.{m}.fc1.weight"] = rnd(mlp, hidden)
            sd[p + f"mlp.{m}.fc1.bias"] = rnd(mlp)
            sd[p + f"mlp.{m}.fc2.weight"] = rnd(hidden, mlp)
            sd[p + f"mlp.{m}.fc2.bias"] = rnd(hidden)
            sd[p + f"norm2.{m}.weight"] = rnd(hidden)
            sd[p + f"norm2.{m}.bias"] = rnd(hidden)
    return sd


def synthetic_ufo_block_state_dict(hidden, mlp, num_layers, gen):
    sd = {}

    def rnd(*shape):
        return torch.randn(*shape, generator=gen, dtype=torch.float32)

    for i in range(num_layers):
        p = f"transformer.blocks.{i}."
        sd[p + "attn.q_bias"] = rnd(hidden)
        sd[p + "attn.v_bias"] = rnd(hidden)
        sd[p + "attn.qkv.weight"] = rnd(3 * hidden, hidden)
        sd[p + "attn.proj.weight"] = rnd(hidden, hidden)
        sd[p + "attn.proj.bias"] = rnd(hidden)
        sd[p + "norm1.weight"] = rnd(hidden)
        sd[p + "norm1.bias"] = rnd(hidden)
        sd[p + "mlp.fc1.weight"] = rnd(mlp, hidden)
        sd[p + "mlp.fc1.bias"] = rnd(mlp)
        sd[p + "mlp.fc2.weight"] = rnd(hidden, mlp)
        sd[p + "mlp.fc2.bias"] = rnd(hidden)
        sd[p + "norm2.weight"] = rnd(hidden)
        sd[p + "norm2.bias"] = rnd(hidden)
    return sd


def synthetic_grams(hidden, mlp, num_layers, vlffn_start, gen, with_vl=True):
    """fp64 SPD Gram matrices keyed like the reference's Gram file."""
    from collections import defaultdict

    grams = defaultdict(float)
    for i in range(num_layers):
        mods = ["v", "l"] + (["vl"] if (with_vl and i >= vlffn_start) else [])
        for m in mods:
            for suffix, d in ((f"attn.{m}", hidden), (f"attn.{m}.proj", hidden), (f"mlp.{m}.fc1", hidden),
                              (f"mlp.{m}.fc2", mlp)):
                x = torch.randn(3 * d, d, generator=gen, dtype=torch.float64) + 0.3
                grams[f"transformer.blocks.{i}.{suffix}"] = x.T @ x
    return grams


def _np(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def make_merge_golden(path=None):
    path = path or os.path.join(GOLDEN, "merge_small.npz")
    hidden, mlp, L, vl0 = 8, 32, 12, 10
    gen = torch.Generator().manual_seed(20261017)
    sd = synthetic_all_moe_state_dict(hidden, mlp, L, vl0, gen)
    central = synthetic_ufo_block_state_dict(hidden, mlp, L, gen)
    grams = synthetic_grams(hidden, mlp, L, vl0, gen)

    # a checkpoint in which layer 11 was already shared: no expert keys, ufo keys instead
    # (exercises the `else: later_weight = state_dict[later_name]; break` branches)
    sd_shared11 = {k: v for k, v in sd.items() if not (k.startswith("transformer.blocks.11.") and "gamma" not in k)}
    for k, v in synthetic_ufo_block_state_dict(hidden, mlp, L, gen).items():
        if k.startswith("transformer.blocks.11."):
            sd_shared11[k] = v
    # Gram file lacking the language Grams of layer 3 (regmean then skips that modality, :419-420)
    grams_missing = {k: v for k, v in grams.items() if not (k.startswith("transformer.blocks.3.") and ".l" in k)}

    def cfg(**kw):
        c = dict(vlffn_start_layer_index=vl0, only_activate_used_experts=False, merge_ratio=0.5, sum_lambda=1,
                 scaling_for_non_diag=1, central_weight=None, gram_matrices=None,
                 loss_names=dict(LOSS_ZERO, irtr=1.0))
        c.update(kw)
        return c

    tmp = tempfile.mkdtemp()
    central_path = os.path.join(tmp, "central.pth")
    grams_path = os.path.join(tmp, "grams.pth")
    grams_missing_path = os.path.join(tmp, "grams_missing.pth")
    torch.save(grams, grams_path)
    torch.save(grams_missing, grams_missing_path)

    variants = {
        # name: (method, input sd id, cfg overrides)
        "interp_a0.5": ("merge_weights", "sd", dict(merge_ratio=0.5)),
        "interp_a0.3_used_irtr": ("merge_weights", "sd", dict(merge_ratio=0.3, only_activate_used_experts=True)),
        "interp_used_vqa": ("merge_weights", "sd", dict(merge_ratio=0.7, only_activate_used_experts=True,
                                                        loss_names=dict(LOSS_ZERO, vqa=1))),
        "interp_shared11": ("merge_weights", "sd_shared11", dict(merge_ratio=0.25)),
        "arith_l0.75": ("sum_task_vectors", "sd", dict(sum_lambda=0.75)),
        "arith_l1_used_irtr": ("sum_task_vectors", "sd", dict(sum_lambda=1, only_activate_used_experts=True)),
        "arith_used_nlvr2": ("sum_task_vectors", "sd", dict(sum_lambda=0.4, only_activate_used_experts=True,
                                                            loss_names=dict(LOSS_ZERO, nlvr2=1))),
        "arith_shared11": ("sum_task_vectors", "sd_shared11", dict(sum_lambda=0.6)),
        "regmean_s1.0": ("regmean", "sd", dict(scaling_for_non_diag=1.0)),
        "regmean_s0.9": ("regmean", "sd", dict(scaling_for_non_diag=0.9)),
        "regmean_vqa": ("regmean", "sd", dict(scaling_for_non_diag=0.95, loss_names=dict(LOSS_ZERO, vqa=1))),
        "regmean_all3": ("regmean", "sd", dict(scaling_for_non_diag=0.8, loss_names=dict(LOSS_ZERO, mlm=1))),
        "regmean_missing_gram": ("regmean", "sd", dict(scaling_for_non_diag=0.9, gram_matrices="missing")),
        "regmean_shared11": ("regmean", "sd_shared11", dict(scaling_for_non_diag=0.9)),
    }
    inputs = {"sd": sd, "sd_shared11": sd_shared11}
    out = {}
    meta = {"hidden": hidden, "mlp": mlp, "num_layers": L, "variants": {}}
    for k, v in _np(sd).items():
        out[f"in/sd/{k}"] = v
    for k, v in _np(sd_shared11).items():
        if k.startswith("transformer.blocks.11."):
            out[f"in/sd_shared11/{k}"] = v  # the rest equals sd
    for k, v in _np(central).items():
        out[f"central/{k}"] = v
    for k, v in _np(grams).items():
        out[f"gram/{k}"] = v
    for vname, (method, sd_id, over) in variants.items():
        c = cfg(**over)
        src = {k: v.clone() for k, v in inputs[sd_id].items()}
        if method == "merge_weights":
            res = rh.ref_merge_weights(src, c)
        elif method == "sum_task_vectors":
            # the reference mutates the loaded central dict: give it a fresh file each time
            torch.save({"state_dict": {k: v.clone() for k, v in central.items()}}, central_path)
            c["central_weight"] = central_path
            res = rh.ref_sum_task_vectors(src, c)
        else:
            c["gram_matrices"] = grams_missing_path if over.get("gram_matrices") == "missing" else grams_path
            res = rh.ref_regmean(src, c)
        # inputs must not have been mutated by the reference
        for k, v in inputs[sd_id].items():
            assert torch.equal(v, src[k]), (vname, k)
        c_meta = dict(c)
        c_meta["central_weight"] = None
        c_meta["gram_matrices"] = "missing" if over.get("gram_matrices") == "missing" else None
        meta["variants"][vname] = {"method": method, "input": sd_id, "cfg": c_meta,
                                   "keys": list(res.keys())}
        for k, v in res.items():
            if "transformer.blocks." in k and "gamma" not in k:
                out[f"out/{vname}/{k}"] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB, {len(variants)} variants")


if __name__ == "__main__":
    what = sys.argv[1:] or ["merge", "model"]
    if "merge" in what:
        make_merge_golden()
    if "model" in what:
        try:
            from make_golden_model import make_model_golden
        except ImportError:
            make_model_golden = None
        if make_model_golden:
            make_model_golden()
