"""make_golden_ckpt.py — TEST INFRASTRUCTURE.  Golden vector for modify_checkpoint_vlmo: the UNMODIFIED
reference method (src/vilt/modules/vilt_module.py:749-806) applied to a 224-px tiny checkpoint by a 384-px
tiny reference model (27x27 -> 47x47 bicubic resize of the relative position bias table) and, second case,
by a model with a shorter max_text_len (position table truncation).  Writes tests/golden/ckpt_tiny.npz."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_harness as rh  # noqa: E402

TASK = "task_finetune_irtr_coco_square_randaug_base_image384"


def main():
    import vl_merging_b200 as vlm

    src = vlm.init_synthetic_(vlm.VLMo(vlm.vlmo_config("tiny")).eval(), seed=3)   # "trained" at 224 px
    out, meta = {}, {}
    cases = {
        "resize_224_to_384": dict(vit="vit_tiny_patch16_224", image_size=384),  # table size follows config image_size
        "truncate_text_40_to_32": dict(vit="vit_tiny_patch16_224", image_size=224, max_text_len=32),
    }
    for name, over in cases.items():
        cfg = rh.make_config([TASK, "all_moe"], hidden_size=192, num_heads=3, load_path="", random_initialization=True,
                             per_gpu_batchsize=2, **over)
        ref = rh.build_model(cfg)
        ckpt = {"state_dict": {k: v.clone() for k, v in src.state_dict().items()}}
        # checkpoints of the reference's era carry BertEmbeddings.position_ids (persistent buffer then)
        ckpt["state_dict"]["text_embeddings.position_ids"] = torch.arange(40).expand((1, -1)).clone()
        res = ref.modify_checkpoint_vlmo(ckpt)
        meta[name] = {"cfg": {k: cfg[k] for k in ("image_size", "patch_size", "max_text_len", "max_text_len_of_initckpt")},
                      "keys": list(res.keys())}
        out[f"{name}/relative_position_bias_table"] = res["relative_position_bias_table"].numpy()
        out[f"{name}/text_pos"] = res["text_embeddings.position_embeddings.weight"].numpy()
        out[f"{name}/position_ids"] = res["text_embeddings.position_ids"].numpy()
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(ROOT, "tests", "golden", "ckpt_tiny.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
