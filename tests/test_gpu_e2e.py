"""End to end on the GPU against golden vectors from the UNMODIFIED reference model (tiny width):
Gram caching through the hooks of the stock-torch forward, then the three merges, then the merged
model's IRTR features and similarity scores (BASELINE.json: Gram 1e-3, logits / scores 1e-3)."""
import json
import os

import numpy as np
import pytest
import torch

import vl_merging_b200 as vlm

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model_tiny.npz")


@pytest.fixture(scope="module")
def setup():
    z = np.load(GOLDEN)
    meta = json.loads(bytes(z["meta"]).decode())
    torch.backends.cuda.matmul.allow_tf32 = False      # the golden forward ran in plain fp32 on the CPU
    torch.backends.cudnn.allow_tf32 = False
    cfg = vlm.vlmo_config("tiny")
    model = vlm.init_synthetic_(vlm.VLMo(cfg).eval(), seed=1).cuda()
    cache = vlm.GramCache()
    cache.register(model, use_moe=True)
    cache.fp64 = vlm.GramCache(precision="fp64")       # the RegMean-grade mode, same hooks, same forward
    cache.fp64.register(model, use_moe=True)
    with torch.no_grad():
        for bs, seed, pad in meta["calib_batches"]:
            model(vlm.synthetic_batch(bs, cfg, seed=seed, pad=pad, device="cuda"))
    cache.remove_hooks()
    cache.fp64.remove_hooks()
    return z, meta, cfg, model, cache


def test_gram_caching_matches_reference_hooks(setup):
    """Against the golden summaries of the reference's own hooks (reference forward on CPU): diagonal,
    Frobenius norm and trace have no cancellation, so the 1e-3 Gram tolerance applies to them directly; the
    row sums do cancel, so they get the Cauchy-Schwarz form of the same bound; one Gram is compared in full."""
    z, meta, cfg, model, cache = setup
    grams = cache.state_dict()
    assert sorted(grams.keys()) == sorted(meta["gram_keys"]) and len(grams) == 96
    for k in meta["gram_keys"]:
        g = grams[k].numpy()
        fro, trace = z[f"gram/{k}/fro_trace"]
        assert np.linalg.norm(np.diag(g) - z[f"gram/{k}/diag"]) < 1e-3 * np.linalg.norm(z[f"gram/{k}/diag"]), k
        assert abs(np.linalg.norm(g) - fro) < 1e-3 * fro and abs(np.trace(g) - trace) < 1e-3 * trace, k
        assert np.linalg.norm(g.sum(1) - z[f"gram/{k}/rowsum"]) < 1e-3 * fro * np.sqrt(g.shape[0]), k
    k = "transformer.blocks.0.attn.v"
    full = z[f"gram_full/{k}"]
    assert np.linalg.norm(grams[k].numpy() - full) / np.linalg.norm(full) < 1e-3


def test_gram_caching_matches_fp64_hook_on_identical_activations(setup):
    """The reference hook's arithmetic (fp64 X^T X, accumulated) on the very activations our hook saw:
    all 96 Grams in full, relative Frobenius error <= 1e-3 (BASELINE.json)."""
    z, meta, cfg, model, cache = setup
    ref = {}
    mods = dict(model.named_modules())
    handles = []
    for name in meta["gram_keys"]:
        def probe(m, i, o, name=name):
            x = (i[0] if isinstance(i, tuple) else i).double()
            x = x.reshape(-1, x.shape[-1])
            ref[name] = ref.get(name, 0) + x.T @ x
        handles.append(mods[name].register_forward_hook(probe))
    with torch.no_grad():
        for bs, seed, pad in meta["calib_batches"]:
            model(vlm.synthetic_batch(bs, cfg, seed=seed, pad=pad, device="cuda"))
    for h in handles:
        h.remove()
    worst = max(((cache.gram(k).double() - ref[k]).norm() / ref[k].norm()).item() for k in meta["gram_keys"])
    assert worst < 1e-3, worst


@pytest.mark.parametrize("vname", ["interp", "arith", "regmean"])
def test_merged_model_matches_reference_merged_model(setup, vname):
    z, meta, cfg, model, cache = setup
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    mcfg = dict(vlffn_start_layer_index=cfg["vlffn_start_layer_index"], only_activate_used_experts=False,
                loss_names={"irtr": 1.0, "vqa": 0, "nlvr2": 0}, merge_ratio=0.5, sum_lambda=1,
                scaling_for_non_diag=1)
    mcfg.update(meta["variants"][vname]["cfg"])
    if vname == "interp":
        merged = vlm.merge_weights(sd, mcfg)
    elif vname == "arith":
        central = vlm.init_synthetic_(vlm.VLMo(vlm.vlmo_config("tiny", use_moe=False)), seed=2).cuda().state_dict()
        merged = vlm.sum_task_vectors(sd, mcfg, central_weight=central)
    else:
        # Grams straight from the device cache in the RegMean-grade (fp64) mode: BASELINE's 1e-4 against the merged
        # weights of the unmodified reference, whose Grams came from ITS forward on the CPU
        merged = vlm.regmean(sd, mcfg, gram_matrices=cache.fp64)
        fast = vlm.regmean(sd, mcfg, gram_matrices=cache)      # single-pass TF32 Grams (Gram tolerance 1e-3)
    for k in ("transformer.blocks.0.attn.qkv.weight", "transformer.blocks.11.mlp.fc2.weight", "transformer.blocks.5.norm1.bias"):
        want = z[f"merged/{vname}/tensor/{k}"]
        got = merged[k].cpu().numpy()[:8]
        tol = 0.0 if vname != "regmean" else 1e-4
        assert np.linalg.norm(got - want) <= tol * np.linalg.norm(want), (vname, k, np.linalg.norm(got - want) / np.linalg.norm(want))
        if vname == "regmean":
            err = np.linalg.norm(fast[k].cpu().numpy()[:8] - want) / np.linalg.norm(want)
            assert err <= 5e-3, (k, err)
    ufo = vlm.VLMo(vlm.vlmo_config("tiny", use_moe=False)).eval().cuda()
    missing, unexpected = ufo.load_state_dict(merged, strict=False)   # vilt_module.py:293
    assert not [m for m in missing if "transformer.blocks" in m]
    (ni, si), (nt, st) = meta["eval"]
    with torch.no_grad():
        i_cls = ufo.infer_image_ft(vlm.synthetic_batch(ni, cfg, seed=si, device="cuda"))["cls_feats"]
        t_cls = ufo.infer_text_ft(vlm.synthetic_batch(nt, cfg, seed=st, pad=True, device="cuda"))["cls_feats"]
    scores = (i_cls @ t_cls.t()).cpu().numpy()
    assert np.abs(i_cls.cpu().numpy() - z[f"merged/{vname}/img_cls"]).max() < 1e-3
    assert np.abs(t_cls.cpu().numpy() - z[f"merged/{vname}/txt_cls"]).max() < 1e-3
    assert np.abs(scores - z[f"merged/{vname}/scores"]).max() < 1e-3


def test_regmean_with_reference_format_gram_file_is_tight(setup, tmp_path):
    """Same Grams on both sides (exported in the reference's fp64 file format, fed to our regmean AND to
    the numpy oracle): isolates kernel (c) + solve from the TF32 Gram error.  Tolerance 1e-4 (BASELINE)."""
    import oracle

    z, meta, cfg, model, cache = setup
    cache.save(tmp_path / "grams.pth")
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    mcfg = dict(vlffn_start_layer_index=10, loss_names={"irtr": 1.0, "vqa": 0, "nlvr2": 0}, scaling_for_non_diag=0.9,
                gram_matrices=str(tmp_path / "grams.pth"))
    got = vlm.regmean(sd, mcfg)
    grams = {k: v.numpy() for k, v in torch.load(tmp_path / "grams.pth", weights_only=False).items()}
    want = oracle.regmean({k: v.numpy() for k, v in sd.items()}, grams, mcfg)
    for k, w in want.items():
        if "transformer.blocks." in k and "gamma" not in k:
            g = got[k].numpy()
            assert g.dtype == w.dtype
            assert np.linalg.norm(g - w) <= 1e-4 * np.linalg.norm(w), k
