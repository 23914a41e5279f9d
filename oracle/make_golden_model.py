"""make_golden_model.py — TEST INFRASTRUCTURE.  Golden vectors for the calibration forward and the
merged-model forward, produced by the UNMODIFIED reference model at the vit_tiny width
(vision_transformer.py:1260-1266; 12 layers, hidden 192, 224 px -> 197 image tokens, 40 text tokens):

  1. weights: vl_merging_b200.init_synthetic_ (an integer hash, exact on every machine) on our VLMo
     mirror, loaded into the reference ViLTransformerSS (all_moe);
  2. reference hooks (registration rule + hook_gram_input restated verbatim in ref_harness) around the
     reference's infer_image_ft / infer_text_ft on a seeded synthetic batch -> Gram summaries;
  3. reference merge_weights / sum_task_vectors / regmean on that state_dict -> load into a reference
     ufo model -> cls_feats and the image x text similarity matrix (objectives.py:684).

Writes tests/golden/model_tiny.npz (inputs are regenerated from seeds by the tests).
"""
import json
import os
import sys
import tempfile
from collections import defaultdict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_harness as rh  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
TASK = "task_finetune_irtr_coco_square_randaug_base_image384"
TINY = dict(vit="vit_tiny_patch16_224", hidden_size=192, num_heads=3, image_size=224, load_path="",
            random_initialization=True, per_gpu_batchsize=2)
CALIB_BATCHES = [(4, 11, False), (3, 12, True)]   # (batch size, seed, ragged text)
EVAL_IMAGES, EVAL_TEXTS = (5, 21), (7, 22)


def gram_summary(g):
    g = g.double()
    return {"diag": g.diag().numpy(), "rowsum": g.sum(1).numpy(), "fro": float(g.norm()), "trace": float(g.trace())}


def make_model_golden(path=None):
    import vl_merging_b200 as vlm

    path = path or os.path.join(GOLDEN, "model_tiny.npz")
    cfg = vlm.vlmo_config("tiny")
    mine = vlm.init_synthetic_(vlm.VLMo(cfg).eval(), seed=1)
    ref_cfg = rh.make_config([TASK, "all_moe"], **TINY)
    ref = rh.build_model(ref_cfg)
    missing, unexpected = ref.load_state_dict(mine.state_dict(), strict=False)
    assert not unexpected, unexpected

    out, meta = {}, {"calib_batches": CALIB_BATCHES, "eval": [EVAL_IMAGES, EVAL_TEXTS]}

    # ---- 2. calibration with the reference's hooks ------------------------------------------------
    store = defaultdict(float)
    handles = rh.ref_register_gram_hooks(ref, store, use_moe=True)
    meta["hooked_modules"] = [m.module_name for m in ref.modules() if hasattr(m, "module_name")]
    with torch.no_grad():
        for bs, seed, pad in CALIB_BATCHES:
            batch = vlm.synthetic_batch(bs, cfg, seed=seed, pad=pad)
            img = ref.infer_image_ft(batch)["cls_feats"]
            txt = ref.infer_text_ft(batch)["cls_feats"]
    for h in handles:
        h.remove()
    out["calib/last_img_cls"] = img.numpy()
    out["calib/last_txt_cls"] = txt.numpy()
    meta["gram_keys"] = list(store.keys())
    for k, g in store.items():
        s = gram_summary(g)
        out[f"gram/{k}/diag"], out[f"gram/{k}/rowsum"] = s["diag"], s["rowsum"]
        out[f"gram/{k}/fro_trace"] = np.array([s["fro"], s["trace"]])
    for k in ("transformer.blocks.0.attn.v",):
        out[f"gram_full/{k}"] = store[k].numpy()

    # ---- 3. merges + merged-model forward -----------------------------------------------------------
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    central_model = vlm.init_synthetic_(vlm.VLMo(vlm.vlmo_config("tiny", use_moe=False)).eval(), seed=2)
    tmp = tempfile.mkdtemp()
    central_path, gram_path = os.path.join(tmp, "central.pth"), os.path.join(tmp, "grams.pth")
    torch.save(store, gram_path)
    base = dict(ref_cfg)
    variants = {
        "interp": ("merge_weights", dict(merge_ratio=0.5)),
        "arith": ("sum_task_vectors", dict(sum_lambda=0.75)),
        "regmean": ("regmean", dict(scaling_for_non_diag=0.9)),
    }
    meta["variants"] = {k: {"method": m, "cfg": c} for k, (m, c) in variants.items()}
    ufo_cfg = rh.make_config([TASK, "ufo"], **TINY)
    img_batch = vlm.synthetic_batch(EVAL_IMAGES[0], cfg, seed=EVAL_IMAGES[1])
    txt_batch = vlm.synthetic_batch(EVAL_TEXTS[0], cfg, seed=EVAL_TEXTS[1], pad=True)
    for vname, (method, over) in variants.items():
        c = dict(base, **over)
        src = {k: v.clone() for k, v in sd.items()}
        if method == "merge_weights":
            merged = rh.ref_merge_weights(src, c)
        elif method == "sum_task_vectors":
            torch.save({"state_dict": central_model.state_dict()}, central_path)
            c["central_weight"] = central_path
            merged = rh.ref_sum_task_vectors(src, c)
        else:
            c["gram_matrices"] = gram_path
            merged = rh.ref_regmean(src, c)
        ufo = rh.build_model(ufo_cfg)
        missing, unexpected = ufo.load_state_dict(merged, strict=False)  # vilt_module.py:293
        assert not [m for m in missing if "transformer.blocks" in m], missing
        with torch.no_grad():
            i_cls = ufo.infer_image_ft(img_batch)["cls_feats"]
            t_cls = ufo.infer_text_ft(txt_batch)["cls_feats"]
        out[f"merged/{vname}/img_cls"] = i_cls.numpy()
        out[f"merged/{vname}/txt_cls"] = t_cls.numpy()
        out[f"merged/{vname}/scores"] = (i_cls @ t_cls.t()).numpy()
        # a few merged tensors, to localise a failure (full state_dicts are too large for a fixture)
        for k in ("transformer.blocks.0.attn.qkv.weight", "transformer.blocks.11.mlp.fc2.weight",
                  "transformer.blocks.5.norm1.bias"):
            out[f"merged/{vname}/tensor/{k}"] = merged[k].numpy()[:8]  # first rows only: keeps the fixture small
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB; {len(store)} Grams")


if __name__ == "__main__":
    make_model_golden()
