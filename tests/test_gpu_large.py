"""Config 5 (ViT-L/16 multiway: 24 layers, width 1024 / 4096, `vl` experts from layer 21 — src/vilt/config.py:454-475)
and config 4 at its real size (5,000 images x 25,000 captions, src/vilt/modules/objectives.py:655-710) on the CUDA
path.  The reference's merge methods hard-code `range(12)` (vilt_module.py:395,553,665); the large-config oracle is
the same code with that literal replaced by num_layers."""
import numpy as np
import pytest
import torch

import oracle
import vl_merging_b200 as vlm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def large_sd():
    cfg = vlm.vlmo_config("large")
    assert (cfg["num_layers"], cfg["hidden_size"], cfg["vlffn_start_layer_index"]) == (24, 1024, 21)
    with torch.device("cuda"):
        model = vlm.VLMo(cfg)
    vlm.init_synthetic_(model, seed=1)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    experts = {k.split(".")[2] + k.split(".")[4] for k in sd if ".mlp." in k and k.endswith("fc1.weight")}
    assert len(experts) == 51                       # 24 x {v, l} + 3 x vl
    return cfg, sd


LARGE_CFG = dict(vlffn_start_layer_index=21, only_activate_used_experts=False, merge_ratio=0.3, sum_lambda=0.75,
                 loss_names={"irtr": 1.0, "vqa": 0, "nlvr2": 0})


def _blocks(d):
    return [k for k in d if "transformer.blocks." in k and "gamma" not in k]


def test_vitl_interpolation_bit_exact_vs_oracle(large_sd):
    """All 51 experts -> 24 x 13 merged tensors, 3-source layers (21-23) with the 2/3 a, 2/3 (1-a), 1/3 ratios."""
    cfg, sd = large_sd
    np_sd = {k: v.cpu().numpy() for k, v in sd.items()}
    want = oracle.merge_weights(np_sd, LARGE_CFG, num_layers=24)
    stats = {}
    got = vlm.merge_weights(sd, LARGE_CFG, num_layers=24, stats=stats)
    assert list(got.keys()) == list(want.keys())
    keys = _blocks(want)
    assert len(keys) == 24 * 13            # 7 key templates -> 13 tensors per layer (vilt_module.py:376-384)
    assert stats["merge_bytes"] == 4 * sum((3 if int(k.split(".")[2]) >= 21 else 2) * want[k].size + want[k].size for k in keys)
    for k in keys:
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k
    used = dict(LARGE_CFG, only_activate_used_experts=True)
    want = oracle.merge_weights(np_sd, used, num_layers=24)
    got = vlm.merge_weights(sd, used, num_layers=24)
    for k in keys[-39:]:                             # layers 21-23: the branch that differs
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k


def test_vitl_modality_arithmetic_bit_exact_vs_oracle(large_sd):
    cfg, sd = large_sd
    with torch.device("cuda"):
        central = vlm.init_synthetic_(vlm.VLMo(dict(cfg, use_moe=False)), seed=2)
    csd = {k: v.detach() for k, v in central.state_dict().items()}
    np_sd = {k: v.cpu().numpy() for k, v in sd.items()}
    want = oracle.sum_task_vectors(np_sd, {k: v.cpu().numpy() for k, v in csd.items()}, LARGE_CFG, num_layers=24)
    got = vlm.sum_task_vectors(sd, LARGE_CFG, num_layers=24, central_weight=csd)
    assert list(got.keys()) == list(want.keys())
    for k in _blocks(want):
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k


@pytest.mark.parametrize("precision", ["fp64", "int8x4"])
def test_vitl_width_regmean_chain_4096(large_sd, precision):
    """One ViT-L layer's four linear problems — (3072,1024), (1024,1024), (4096,1024), (1024,4096): the 4096-wide
    potrf / potrs and the 1024 / 4096 RHS shapes — with Grams cached on the device in fp64 mode from synthetic
    activations (post-GELU-like for fc2), against the oracle fed with numpy fp64 Grams of the same activations.
    Both RegMean-grade Gram modes (fp64 DMMA; the integer tensor cores at 1024 / 4096 columns).  BASELINE.json: RegMean 1e-4."""
    cfg, sd = large_sd
    layer = 22                                       # has v, l AND vl experts; IRTR uses v and l (vilt_module.py:399-400)
    sub = {k: v for k, v in sd.items() if "transformer.blocks." not in k or f".blocks.{layer}." in k}
    sub = {k.replace(f".blocks.{layer}.", ".blocks.0."): v for k, v in sub.items()}
    gen = torch.Generator(device="cuda").manual_seed(3)
    cache = vlm.GramCache(precision=precision)
    np_grams = {}
    for m, rows in (("v", 9232), ("l", 5120)):
        for name, d, positive in ((f"attn.{m}", 1024, False), (f"attn.{m}.proj", 1024, False),
                                  (f"mlp.{m}.fc1", 1024, False), (f"mlp.{m}.fc2", 4096, True)):
            x = torch.randn(rows, d, device="cuda", generator=gen)
            x = torch.nn.functional.gelu(x) if positive else x + 0.1
            key = f"transformer.blocks.0.{name}"
            cache.accumulate(key, x[: rows // 2])
            cache.accumulate(key, x[rows // 2:].reshape(2, -1, d))          # second call, 3-D like a real activation
            x64 = x.double().cpu().numpy()
            np_grams[key] = x64.T @ x64
    for alpha in (1.0, 0.9):
        mcfg = dict(vlffn_start_layer_index=0, loss_names={"irtr": 1.0, "vqa": 0, "nlvr2": 0}, scaling_for_non_diag=alpha)
        want = oracle.regmean({k: v.cpu().numpy() for k, v in sub.items()}, np_grams, mcfg, num_layers=1)
        got = vlm.regmean(sub, mcfg, gram_matrices=cache, num_layers=1)
        lin = [k for k in _blocks(want) if want[k].ndim == 2]
        assert sorted(want[k].shape for k in lin) == [(1024, 1024), (1024, 4096), (3072, 1024), (4096, 1024)]
        for k in _blocks(want):
            g, w = got[k].cpu().numpy(), want[k]
            assert g.dtype == w.dtype and g.shape == w.shape, k
            assert np.linalg.norm(g - w) <= 1e-4 * np.linalg.norm(w), (alpha, k, np.linalg.norm(g - w) / np.linalg.norm(w))


def test_irtr_5k_x_25k_similarity_and_recalls():
    """objectives.py:684-710 at the COCO 5k test size: the 5,000 x 25,000 score matrix against an fp32 numpy
    recomputation on the same features (1e-3, BASELINE.json) and the six recalls against the oracle's."""
    gen = torch.Generator(device="cuda").manual_seed(11)
    n_img, per = 5000, 5
    base = torch.randn(n_img, 768, device="cuda", generator=gen)
    img = torch.nn.functional.normalize(base + 2.0 * torch.randn(n_img, 768, device="cuda", generator=gen), dim=-1)
    txt = torch.nn.functional.normalize(base.repeat_interleave(per, 0) +
                                        3.0 * torch.randn(n_img * per, 768, device="cuda", generator=gen), dim=-1)
    iids = np.arange(n_img)
    tiids = np.arange(n_img * per) // per
    scores, recalls = vlm.irtr_recall(img, txt, iids, tiids)
    assert tuple(scores.shape) == (5000, 25000)
    ref_scores, ref_recalls = oracle.irtr_recall(img.cpu().numpy(), txt.cpu().numpy(), iids, tiids)
    assert np.abs(scores.cpu().numpy() - ref_scores).max() < 1e-3
    got = np.array([float(r) for r in recalls])
    assert 0.02 < got.min() and got.max() < 0.995                    # a non-trivial retrieval problem
    assert np.allclose(got, np.array(ref_recalls, dtype=np.float64), atol=2.0 / n_img), (got, ref_recalls)
