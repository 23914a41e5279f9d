"""The fused vision-language route (type_id 2) against the reference, through
tests/golden/fused_tiny.npz (written by oracle/make_golden_fused.py from the unmodified reference):
the stock-torch mirror `VLMo.infer`, the Grams its hooks collect (row slices for the shallow `l` / `v`
experts, all tokens for the `vl` experts) and the oracle's RegMean on those Grams for VQA / NLVR2 style
tasks.  CPU only: the Gram arithmetic here is the oracle's (fp64); the CUDA path is checked against the
same golden in tests/test_gpu_fused.py."""
import json
import os
import sys

import numpy as np
import pytest
import torch

import vl_merging_b200 as vlm
from vl_merging_b200.gram import select_hooked_modules

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import oracle  # noqa: E402

GOLDEN = os.path.join(HERE, "golden", "fused_tiny.npz")


@pytest.fixture(scope="module")
def golden():
    z = np.load(GOLDEN)
    return z, json.loads(bytes(z["meta"]).decode())


@pytest.fixture(scope="module")
def tiny():
    cfg = vlm.vlmo_config("tiny")
    return cfg, vlm.init_synthetic_(vlm.VLMo(cfg).eval(), seed=1)


@pytest.fixture(scope="module")
def calibrated(golden, tiny):
    """Our model's `infer` under the reference hook restated by the oracle -> (last outputs, Gram store, views seen)."""
    _, meta = golden
    cfg, model = tiny
    store = oracle.new_gram_store()
    hook = oracle.reference_hook_torch(store)
    views = {}

    def spy(module, input, output):
        x = input[0] if isinstance(input, tuple) else input
        views[module.module_name] = (tuple(x.shape), x.is_contiguous())
        hook(module, input, output)

    handles = []
    for name, module in select_hooked_modules(model, use_moe=True):
        module.module_name = name
        handles.append(module.register_forward_hook(spy))
    with torch.no_grad():
        for bs, seed, pad in meta["calib_batches"]:
            ret = model.infer(vlm.synthetic_batch(bs, cfg, seed=seed, pad=pad))
    for h in handles:
        h.remove()
    return ret, store, views


def test_infer_matches_reference_golden(golden, calibrated):
    z, _ = golden
    ret, _, _ = calibrated
    assert np.abs(ret["cls_feats"].numpy() - z["calib/last_cls"]).max() < 1e-5
    assert np.abs(ret["raw_cls_feats"].numpy() - z["calib/last_raw_cls"]).max() < 1e-5
    assert np.abs(ret["text_feats"].numpy()[:, :4] - z["calib/last_text_feats"]).max() < 1e-5
    assert np.abs(ret["image_feats"].numpy()[:, :4] - z["calib/last_image_feats"]).max() < 1e-5


def test_fused_route_fires_the_reference_gram_keys(golden, calibrated):
    _, meta = golden
    _, store, views = calibrated
    assert list(store.keys()) == meta["gram_keys"]
    assert len(store) == 88                                   # 10 layers x (v, l) x 4 + 2 layers x vl x 4
    # shallow experts are fed row SLICES of the joint sequence: non-contiguous (B, n, D) views
    assert views["transformer.blocks.0.attn.l"] == ((3, 40, 192), False)
    assert views["transformer.blocks.0.attn.v"] == ((3, 197, 192), False)
    assert views["transformer.blocks.4.mlp.v.fc1"] == ((3, 197, 192), False)
    assert views["transformer.blocks.4.mlp.v.fc2"][1] is True
    assert views["transformer.blocks.11.attn.vl"] == ((3, 237, 192), True)


def test_fused_route_grams_match_reference_golden(golden, calibrated):
    z, meta = golden
    _, store, _ = calibrated
    for k in meta["gram_keys"]:
        g = store[k].numpy()
        fro, trace = z[f"gram/{k}/fro_trace"]
        assert abs(np.trace(g) - trace) <= 1e-6 * abs(trace), k
        assert abs(np.linalg.norm(g) - fro) <= 1e-6 * fro, k
        assert np.abs(np.diag(g) - z[f"gram/{k}/diag"]).max() <= 1e-6 * np.abs(z[f"gram/{k}/diag"]).max(), k
    for k in (f for f in z.files if f.startswith("gram_full/")):
        name = k[len("gram_full/"):]
        want = z[k]
        assert np.linalg.norm(store[name].numpy() - want) <= 1e-6 * np.linalg.norm(want), name


@pytest.mark.parametrize("variant", ["vqa", "nlvr2"])
def test_regmean_on_fused_grams_matches_reference_golden(golden, tiny, calibrated, variant):
    """vqa: the deep layers take the `vl` expert alone (vilt_module.py:401-402); nlvr2: v, l and vl (:403-404)."""
    z, meta = golden
    cfg, model = tiny
    _, store, _ = calibrated
    sd = {k: v.numpy() for k, v in model.state_dict().items()}
    grams = {k: v.numpy() for k, v in store.items()}
    mcfg = {"loss_names": meta["variants"][variant], "scaling_for_non_diag": 0.9,
            "vlffn_start_layer_index": cfg["vlffn_start_layer_index"]}
    merged = oracle.regmean(sd, grams, mcfg)
    pre = f"merged/{variant}/tensor/"
    for k in (f for f in z.files if f.startswith(pre)):
        want, got = z[k], merged[k[len(pre):]][:8]
        assert np.linalg.norm(got - want) <= 1e-6 * np.linalg.norm(want), k
    ufo = vlm.VLMo(vlm.vlmo_config("tiny", use_moe=False)).eval()
    missing, unexpected = ufo.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in merged.items()}, strict=False)
    assert not [m for m in missing if "transformer.blocks" in m]
    bs, seed, pad = meta["eval_batch"]
    with torch.no_grad():
        ret = ufo.infer(vlm.synthetic_batch(bs, cfg, seed=seed, pad=pad))
    assert np.abs(ret["cls_feats"].numpy() - z[f"merged/{variant}/cls"]).max() < 1e-4
    assert np.abs(ret["raw_cls_feats"].numpy() - z[f"merged/{variant}/raw_cls"]).max() < 2e-4


@pytest.mark.reference
@pytest.mark.parametrize("kind", ["all_moe", "ufo"])
def test_infer_bit_exact_against_imported_reference(kind):
    """`VLMo.infer` vs the unmodified reference `infer` (vilt_module.py:1071-1156) on the same weights and a ragged
    batch: identical bits for both the modality-specific and the merged (shared-weight, split attention) layout."""
    import ref_harness as rh

    cfg = vlm.vlmo_config("tiny", use_moe=(kind == "all_moe"))
    mine = vlm.init_synthetic_(vlm.VLMo(cfg).eval(), seed=5)
    ref_cfg = rh.make_config(["task_finetune_irtr_coco_square_randaug_base_image384", kind],
                             vit="vit_tiny_patch16_224", hidden_size=192, num_heads=3, image_size=224,
                             load_path="", random_initialization=True, per_gpu_batchsize=2)
    ref = rh.build_model(ref_cfg)
    _, unexpected = ref.load_state_dict(mine.state_dict(), strict=False)
    assert not unexpected
    batch = vlm.synthetic_batch(3, cfg, seed=17, pad=True)
    rbatch = dict(batch, image=[batch["image"]] if not isinstance(batch["image"], (list, tuple)) else batch["image"])
    with torch.no_grad():
        a, b = ref.infer(rbatch), mine.infer(batch)
    for key in ("cls_feats", "raw_cls_feats", "text_feats", "image_feats"):
        assert torch.equal(a[key], b[key]), key
