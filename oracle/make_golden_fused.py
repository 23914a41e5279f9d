"""make_golden_fused.py — TEST INFRASTRUCTURE.  Golden vectors for Gram caching on the FUSED vision-language
route, produced by the UNMODIFIED reference model at the vit_tiny width:

  1. weights: vl_merging_b200.init_synthetic_ loaded into the reference ViLTransformerSS (all_moe);
  2. the reference's hooks (ref_harness: registration rule + hook_gram_input verbatim) around the reference's
     `infer` (vilt_module.py:1071-1156, every block with type_id 2): the `vl` experts of the deep layers see all
     40 + 197 tokens, the `l` / `v` experts of the shallow layers see ROW SLICES of the joint sequence
     (vision_transformer.py:619-637, :667-677) -> Gram summaries + a few full Grams;
  3. reference `regmean` with those Grams for a VQA-style task (deep layers: the `vl` expert alone,
     vilt_module.py:401-402) and an NLVR2-style one (deep layers: v, l and vl, :403-404) -> loaded into a
     reference ufo model -> `infer` features.

Writes tests/golden/fused_tiny.npz (inputs are regenerated from seeds by the tests).
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
from make_golden_model import GOLDEN, TASK, TINY, gram_summary  # noqa: E402

CALIB_BATCHES = [(4, 31, False), (3, 32, True)]   # (batch size, seed, ragged text)
EVAL_BATCH = (5, 41, True)
FULL_GRAMS = ("transformer.blocks.0.attn.l", "transformer.blocks.3.mlp.v.fc1", "transformer.blocks.11.attn.vl",
              "transformer.blocks.10.mlp.vl.fc1")


def loss_names(**on):
    base = {k: 0 for k in ("itm", "itc", "mlm", "textmlm", "vqa", "nlvr2", "irtr")}
    base.update(on)
    return base


def make_fused_golden(path=None):
    import vl_merging_b200 as vlm

    path = path or os.path.join(GOLDEN, "fused_tiny.npz")
    cfg = vlm.vlmo_config("tiny")
    mine = vlm.init_synthetic_(vlm.VLMo(cfg).eval(), seed=1)
    ref_cfg = rh.make_config([TASK, "all_moe"], **TINY)
    ref = rh.build_model(ref_cfg)
    missing, unexpected = ref.load_state_dict(mine.state_dict(), strict=False)
    assert not unexpected, unexpected

    out, meta = {}, {"calib_batches": CALIB_BATCHES, "eval_batch": EVAL_BATCH}
    store = defaultdict(float)
    handles = rh.ref_register_gram_hooks(ref, store, use_moe=True)
    with torch.no_grad():
        for bs, seed, pad in CALIB_BATCHES:
            batch = vlm.synthetic_batch(bs, cfg, seed=seed, pad=pad)
            batch["image"] = [batch["image"]] if not isinstance(batch["image"], (list, tuple)) else batch["image"]
            ret = ref.infer(batch)
    for h in handles:
        h.remove()
    out["calib/last_cls"] = ret["cls_feats"].numpy()
    out["calib/last_raw_cls"] = ret["raw_cls_feats"].numpy()
    out["calib/last_text_feats"] = ret["text_feats"].numpy()[:, :4]
    out["calib/last_image_feats"] = ret["image_feats"].numpy()[:, :4]
    meta["gram_keys"] = list(store.keys())
    for k, g in store.items():
        s = gram_summary(g)
        out[f"gram/{k}/diag"], out[f"gram/{k}/rowsum"] = s["diag"], s["rowsum"]
        out[f"gram/{k}/fro_trace"] = np.array([s["fro"], s["trace"]])
    for k in FULL_GRAMS:
        out[f"gram_full/{k}"] = store[k].numpy()

    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    tmp = tempfile.mkdtemp()
    gram_path = os.path.join(tmp, "grams.pth")
    torch.save(store, gram_path)
    variants = {"vqa": loss_names(vqa=1), "nlvr2": loss_names(nlvr2=1)}
    meta["variants"] = variants
    ufo_cfg = rh.make_config([TASK, "ufo"], **TINY)
    bs, seed, pad = EVAL_BATCH
    eval_batch = vlm.synthetic_batch(bs, cfg, seed=seed, pad=pad)
    eval_batch["image"] = [eval_batch["image"]] if not isinstance(eval_batch["image"], (list, tuple)) else eval_batch["image"]
    for vname, names in variants.items():
        c = dict(ref_cfg, loss_names=names, scaling_for_non_diag=0.9, gram_matrices=gram_path)
        merged = rh.ref_regmean({k: v.clone() for k, v in sd.items()}, c)
        ufo = rh.build_model(ufo_cfg)
        missing, unexpected = ufo.load_state_dict(merged, strict=False)
        assert not [m for m in missing if "transformer.blocks" in m], missing
        with torch.no_grad():
            ret = ufo.infer(eval_batch)
        out[f"merged/{vname}/cls"] = ret["cls_feats"].numpy()
        out[f"merged/{vname}/raw_cls"] = ret["raw_cls_feats"].numpy()
        for k in ("transformer.blocks.0.attn.qkv.weight", "transformer.blocks.11.mlp.fc2.weight",
                  "transformer.blocks.10.attn.proj.weight", "transformer.blocks.11.norm2.weight"):
            out[f"merged/{vname}/tensor/{k}"] = merged[k].numpy()[:8]
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB; {len(store)} Grams")


if __name__ == "__main__":
    make_fused_golden()
