"""Load-time dispatch (src/vilt/modules/vilt_module.py:270-295) executed by the UNMODIFIED reference
in the build container (marker `reference`): a ufo model constructed with `load_path=<all_moe checkpoint>` and one of
`merge_weights / sum_task_vectors / regmean` switched on runs torch.load -> modify_checkpoint_vlmo -> the merge method
-> load_state_dict(strict=False).  Its resulting weights must equal the oracle's merge of the same checkpoint (and the
host plan's), i.e. the merged state_dict is consumed by the reference exactly as produced."""
from collections import defaultdict

import numpy as np
import pytest
import torch

import oracle
import vl_merging_b200 as vlm
from test_plan import interpret
from vl_merging_b200 import plan as P
from vl_merging_b200.gram import select_hooked_modules

TASK = "task_finetune_irtr_coco_square_randaug_base_image384"
TINY = dict(vit="vit_tiny_patch16_224", hidden_size=192, num_heads=3, image_size=224, per_gpu_batchsize=2)


@pytest.fixture(scope="module")
def artefacts(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("dispatch")
    cfg = vlm.vlmo_config("tiny")
    moe = vlm.init_synthetic_(vlm.VLMo(cfg).eval(), seed=1)
    sd = {k: v.clone() for k, v in moe.state_dict().items()}
    sd["text_embeddings.position_ids"] = torch.arange(40).expand((1, -1)).clone()
    torch.save({"state_dict": sd}, tmp / "all_moe.ckpt")
    central = vlm.init_synthetic_(vlm.VLMo(vlm.vlmo_config("tiny", use_moe=False)).eval(), seed=2)
    torch.save({"state_dict": central.state_dict()}, tmp / "central.ckpt")
    store = defaultdict(float)
    hook = oracle.reference_hook_torch(store)
    handles = []
    for name, module in select_hooked_modules(moe, use_moe=True):
        module.module_name = name
        handles.append(module.register_forward_hook(hook))
    with torch.no_grad():
        for seed in (1, 2):
            moe(vlm.synthetic_batch(4, cfg, seed=seed))
    for h in handles:
        h.remove()
    torch.save(store, tmp / "grams.pth")
    return tmp, cfg, sd, central.state_dict(), store


@pytest.mark.reference
@pytest.mark.parametrize("method", ["merge_weights", "sum_task_vectors", "regmean"])
def test_reference_constructor_consumes_the_merge(artefacts, method):
    import ref_harness as rh

    tmp, cfg, sd, central, store = artefacts
    over = {"merge_weights": dict(merge_weights=True, merge_ratio=0.3),
            "sum_task_vectors": dict(sum_task_vectors=True, sum_lambda=0.75, central_weight=str(tmp / "central.ckpt")),
            "regmean": dict(regmean=True, scaling_for_non_diag=0.9, gram_matrices=str(tmp / "grams.pth"))}[method]
    ref_cfg = rh.make_config([TASK, "ufo"], load_path=str(tmp / "all_moe.ckpt"), **TINY, **over)
    ref = rh.build_model(ref_cfg)                      # the constructor performs load + merge (:270-295)
    got = ref.state_dict()

    np_sd = {k: v.numpy() for k, v in sd.items()}
    mcfg = {k: ref_cfg[k] for k in ("vlffn_start_layer_index", "only_activate_used_experts", "merge_ratio", "sum_lambda",
                                    "scaling_for_non_diag", "loss_names")}
    if method == "merge_weights":
        want = oracle.merge_weights(np_sd, mcfg)
        planned = interpret(P.plan_merge_weights(np_sd.keys(), mcfg), np_sd)
    elif method == "sum_task_vectors":
        np_c = {k: v.numpy() for k, v in central.items()}
        want = oracle.sum_task_vectors(np_sd, {k: v.copy() for k, v in np_c.items()}, mcfg)
        planned = interpret(P.plan_sum_task_vectors(np_sd.keys(), np_c.keys(), mcfg), np_sd, central=np_c)
    else:
        grams = {k: v.numpy() for k, v in store.items()}
        want = oracle.regmean(np_sd, grams, mcfg)
        planned = interpret(P.plan_regmean(np_sd.keys(), grams.keys(), mcfg), np_sd, grams=grams,
                            alpha=mcfg["scaling_for_non_diag"])
    checked = 0
    for k, w in want.items():
        if "transformer.blocks." not in k or "gamma" in k:
            continue
        w32 = torch.from_numpy(np.asarray(w)).to(torch.float32)      # load_state_dict casts RegMean's fp64 linears
        p32 = torch.from_numpy(np.asarray(planned[k])).to(torch.float32)
        assert k in got, k
        if np.asarray(w).dtype == np.float32:
            assert torch.equal(got[k], w32) and torch.equal(p32, w32), k
        else:
            assert (got[k] - w32).norm() <= 1e-6 * w32.norm(), k
            assert (p32 - w32).norm() <= 1e-6 * w32.norm(), k
        checked += 1
    assert checked == 12 * 13      # qkv.weight, proj w+b, q_bias, v_bias, fc1 w+b, fc2 w+b, norm1 w+b, norm2 w+b
