"""modify_checkpoint_vlmo (vl-merging_b200/checkpoint.py) against the reference method's golden output
(tests/golden/ckpt_tiny.npz, oracle/make_golden_ckpt.py): bit-exact on CPU (same torch bicubic call)."""
import json
import os

import numpy as np
import torch

import vl_merging_b200 as vlm

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ckpt_tiny.npz")


def test_modify_checkpoint_vlmo_matches_reference(tmp_path):
    z = np.load(GOLDEN)
    meta = json.loads(bytes(z["meta"]).decode())
    src = vlm.init_synthetic_(vlm.VLMo(vlm.vlmo_config("tiny")).eval(), seed=3)
    for name, m in meta.items():
        ckpt = {"state_dict": {k: v.clone() for k, v in src.state_dict().items()}}
        ckpt["state_dict"]["text_embeddings.position_ids"] = torch.arange(40).expand((1, -1)).clone()
        path = tmp_path / f"{name}.ckpt"
        vlm.save_checkpoint(ckpt["state_dict"], path)                       # PL-style file round trip
        res = vlm.modify_checkpoint_vlmo(vlm.load_checkpoint(path), m["cfg"])
        assert list(res.keys()) == m["keys"]
        assert np.array_equal(res["relative_position_bias_table"].numpy(), z[f"{name}/relative_position_bias_table"])
        assert np.array_equal(res["text_embeddings.position_embeddings.weight"].numpy(), z[f"{name}/text_pos"])
        assert np.array_equal(res["text_embeddings.position_ids"].numpy(), z[f"{name}/position_ids"])
    big = z["resize_224_to_384/relative_position_bias_table"]
    assert big.shape == ((2 * 24 - 1) ** 2 + 3 + 2 * 196 + 2, 36)           # 27x27 -> 47x47 image distances


def test_resized_checkpoint_loads_and_merges_shapes():
    cfg384 = vlm.vlmo_config("tiny", image_size=384)
    src = vlm.init_synthetic_(vlm.VLMo(vlm.vlmo_config("tiny")).eval(), seed=3)
    sd = vlm.modify_checkpoint_vlmo({k: v.clone() for k, v in src.state_dict().items()}, cfg384)
    model = vlm.VLMo(cfg384)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected
    assert set(missing) <= {"relative_position_index", "text_relative_position_index",
                            "text_imag_relative_position_index"}  # index buffers are rebuilt from the config
    with torch.no_grad():
        out = model.infer_image_ft(vlm.synthetic_batch(1, cfg384, seed=1))
    assert out["image_feats"].shape == (1, 577, 192)


import pytest  # noqa: E402


@pytest.mark.reference
@pytest.mark.parametrize("image_size,max_text_len", [(288, 40), (320, 32), (448, 40), (384, 20), (224, 40)])
def test_modify_checkpoint_vlmo_against_imported_reference(image_size, max_text_len):
    """More target geometries than the committed golden holds, against the unmodified reference method executed on
    the spot (vilt_module.py:749-806): same keys in the same order, identical bits."""
    import ref_harness as rh

    src = vlm.init_synthetic_(vlm.VLMo(vlm.vlmo_config("tiny")).eval(), seed=4)        # "trained" at 224 px, 40 tokens
    cfg = rh.make_config(["task_finetune_irtr_coco_square_randaug_base_image384", "all_moe"], hidden_size=192,
                         num_heads=3, load_path="", random_initialization=True, per_gpu_batchsize=2,
                         vit="vit_tiny_patch16_224", image_size=image_size, max_text_len=max_text_len)
    ref = rh.build_model(cfg)

    def ckpt():
        sd = {k: v.clone() for k, v in src.state_dict().items()}
        sd["text_embeddings.position_ids"] = torch.arange(40).expand((1, -1)).clone()
        return {"state_dict": sd}

    want = ref.modify_checkpoint_vlmo(ckpt())
    got = vlm.modify_checkpoint_vlmo(ckpt(), {k: cfg[k] for k in ("image_size", "patch_size", "max_text_len",
                                                                "max_text_len_of_initckpt")})
    assert list(got.keys()) == list(want.keys())
    for k, w in want.items():
        assert torch.equal(got[k], w), k
