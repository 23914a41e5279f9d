"""Config 3 end to end (BASELINE.json: RegMean merged weights within 1e-4 of the reference torch path on identical
synthetic inputs): the WHOLE chain on the device — forward hooks -> SYRK Grams (libvlmerge) -> regmean — against
the oracle's regmean (vilt_module.py:366-531 restated) fed with the fp64 Grams of the reference hook
(cache_gram_matrices.py:246-254: X.double().T @ X.double(), accumulated) taken on the very same activations.

The model is VLMo-base WIDTH (768 / 3072, 12 heads) with two layers, so every linear shape of the real merge
occurs — (2304,768), (768,768), (3072,768), (768,3072) — with >= 5,120 rows per expert (text: 10 batches x 16 x 40
= 6,400; image: 92,320)."""
import numpy as np
import pytest
import torch

import vl_merging_b200 as vlm

pytestmark = pytest.mark.gpu

N_BATCH, BS = 10, 16


def _linear_keys(want):
    return [k for k in want if "transformer.blocks." in k and k.endswith(".weight")
            and (".qkv." in k or ".proj." in k or ".fc1." in k or ".fc2." in k)]


@pytest.fixture(scope="module")
def chain():
    import oracle  # noqa: F401  (checker only)

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = vlm.vlmo_config("base", num_layers=2, vlffn_start_layer_index=2)
    model = vlm.init_synthetic_(vlm.VLMo(cfg).eval(), seed=1).cuda()
    caches = {
        "fp64": vlm.GramCache(precision="fp64"),
        "int8x4": vlm.GramCache(precision="int8x4"),
        "tf32": vlm.GramCache(),
        "tf32x3": vlm.GramCache(precision="tf32x3"),
        "tf32x3_grouped": vlm.GramCache(precision="tf32x3", defer_bytes=64 << 20),
    }
    for c in caches.values():
        names = c.register(model, use_moe=True)
    ref = {}
    mods = dict(model.named_modules())
    handles = []

    def probe(module, inputs, output):   # the reference hook, on the device in fp64
        x = (inputs[0] if isinstance(inputs, tuple) else inputs).double()
        x = x.reshape(-1, x.shape[-1])
        ref[module.module_name] = ref.get(module.module_name, 0) + x.T @ x

    for n in names:
        handles.append(mods[n].register_forward_hook(probe))
    with torch.no_grad():
        for b in range(N_BATCH):
            model(vlm.synthetic_batch(BS, cfg, seed=100 + b, pad=(b % 2 == 1), device="cuda"))
    for h in handles:
        h.remove()
    for c in caches.values():
        c.remove_hooks()
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    np_sd = {k: v.cpu().numpy() for k, v in sd.items()}
    np_ref = {k: v.cpu().numpy() for k, v in ref.items()}
    return cfg, sd, np_sd, caches, ref, np_ref


def test_rows_per_expert(chain):
    cfg, sd, np_sd, caches, ref, np_ref = chain
    live = caches["tf32x3"].live_names()
    assert len(live) == 16 and sorted(live) == sorted(ref)          # 2 layers x {v, l} x 4 hooked inputs
    assert min(caches["tf32x3"].rows[n] for n in live) >= 5120


def test_fp64_gram_equals_reference_hook(chain):
    """fp64 mode = the reference hook's arithmetic: only the summation order differs."""
    cfg, sd, np_sd, caches, ref, np_ref = chain
    assert all(caches["fp64"].gram(k).dtype == torch.float64 for k in ref)
    worst = max(((caches["fp64"].gram(k) - g).norm() / g.norm()).item() for k, g in ref.items())
    assert worst < 1e-13, worst
    for k, g in caches["fp64"].state_dict().items():      # the reference's file format: fp64 CPU, full symmetric
        assert g.dtype == torch.float64 and g.device.type == "cpu" and torch.equal(g, g.T)


def test_int8_gram_error(chain):
    """Integer tensor cores: exact products and sums of the quantised activations; what is left is the 2^-27
    quantisation against the column maximum and the dropped digit products (< 2^-26)."""
    cfg, sd, np_sd, caches, ref, np_ref = chain
    assert all(caches["int8x4"].gram(k).dtype == torch.float64 for k in ref)
    worst = max(((caches["int8x4"].gram(k) - g).norm() / g.norm()).item() for k, g in ref.items())
    assert worst < 1e-7, worst


@pytest.mark.parametrize("mode", ["tf32x3", "tf32x3_grouped"])
def test_split_gram_error(chain, mode):
    """Gram error of the split mode: what is left is the tensor core's truncating fp32 accumulation (a near-uniform
    shrink bounded by the segment length, which cancels in (sum W G)(sum G)^-1); the operand-rounding NOISE of the
    single pass, which the inverse amplifies, is gone."""
    cfg, sd, np_sd, caches, ref, np_ref = chain
    worst = max(((caches[mode].gram(k).double() - g).norm() / g.norm()).item() for k, g in ref.items())
    assert worst < 5e-5, worst
    single = max(((caches["tf32"].gram(k).double() - g).norm() / g.norm()).item() for k, g in ref.items())
    assert single < 1e-3, single     # BASELINE.json's Gram tolerance for the single-pass mode


def _regmean_errors(chain, mode, alpha):
    import oracle

    cfg, sd, np_sd, caches, ref, np_ref = chain
    mcfg = dict(vlffn_start_layer_index=2, loss_names={"irtr": 1.0, "vqa": 0, "nlvr2": 0}, scaling_for_non_diag=alpha)
    want = oracle.regmean(np_sd, np_ref, mcfg, num_layers=2)
    got = vlm.regmean(sd, mcfg, gram_matrices=caches[mode], num_layers=2)
    return {k: float(np.linalg.norm(got[k].cpu().numpy() - want[k]) / np.linalg.norm(want[k])) for k in _linear_keys(want)}


@pytest.mark.parametrize("alpha", [1.0, 0.9])
@pytest.mark.parametrize("mode", ["fp64", "int8x4"])
def test_regmean_chain_within_1e4_of_fp64_gram_oracle(chain, mode, alpha):
    import oracle

    cfg, sd, np_sd, caches, ref, np_ref = chain
    mcfg = dict(vlffn_start_layer_index=2, loss_names={"irtr": 1.0, "vqa": 0, "nlvr2": 0}, scaling_for_non_diag=alpha)
    want = oracle.regmean(np_sd, np_ref, mcfg, num_layers=2)
    got = vlm.regmean(sd, mcfg, gram_matrices=caches[mode], num_layers=2)
    keys = _linear_keys(want)
    assert len(keys) == 8
    errs = {}
    for k in keys:
        g = got[k].cpu().numpy()
        assert g.dtype == np.float64 and g.shape == want[k].shape
        errs[k] = float(np.linalg.norm(g - want[k]) / np.linalg.norm(want[k]))
    print(f"{mode} alpha={alpha} regmean errors:", {k.split("blocks.")[1]: f"{e:.2e}" for k, e in errs.items()})
    assert max(errs.values()) <= 1e-4, errs          # BASELINE.json: RegMean 1e-4
    for k, w in want.items():                 # biases / LayerNorms: plain means, exact
        if "transformer.blocks." in k and k not in keys and "gamma" not in k:
            assert np.array_equal(got[k].cpu().numpy(), w), k


@pytest.mark.parametrize("alpha", [1.0, 0.9])
def test_tensor_core_gram_modes_are_reported(chain, alpha):
    """The tcgen05 Gram modes through the same chain.  Their truncating fp32 accumulation leaves a ~2e-5 non-uniform
    shrink in the Gram; where the summed Gram has a small eigen-direction (the LayerNorm-fed qkv / fc1 inputs lie
    near an affine hyperplane) regmean's inverse amplifies it far beyond 1e-4, which is why the RegMean-grade mode is
    fp64.  tf32x3 removes the operand-rounding noise, which is what dominates on the well-conditioned fc2 inputs.
    Recorded and bounded loosely: these modes serve the 1e-3 Gram tolerance, not RegMean's 1e-4."""
    errs = {m: _regmean_errors(chain, m, alpha) for m in ("tf32", "tf32x3", "tf32x3_grouped")}
    for m, e in errs.items():
        print(f"{m} alpha={alpha} regmean errors:", {k.split("blocks.")[1]: f"{v:.2e}" for k, v in e.items()})
        assert max(e.values()) < 0.2, (m, e)
    fc2 = [k for k in errs["tf32"] if ".fc2." in k]
    assert all(errs["tf32x3"][k] <= 1e-4 for k in fc2), errs["tf32x3"]


def test_regmean_warns_about_tf32_grams_without_regularisation(chain):
    """scaling_for_non_diag = 1 on single-pass TF32 Grams is the combination that misses 1e-4: regmean says so and
    names the RegMean-grade modes; it stays quiet for those modes and for scaling < 1."""
    import warnings

    cfg, sd, np_sd, caches, ref, np_ref = chain
    mcfg = dict(vlffn_start_layer_index=2, loss_names={"irtr": 1.0, "vqa": 0, "nlvr2": 0}, scaling_for_non_diag=1.0)
    with pytest.warns(RuntimeWarning, match="int8x4"):
        vlm.regmean(sd, mcfg, gram_matrices=caches["tf32"], num_layers=2)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        vlm.regmean(sd, mcfg, gram_matrices=caches["int8x4"], num_layers=2)
        vlm.regmean(sd, {**mcfg, "scaling_for_non_diag": 0.9}, gram_matrices=caches["tf32"], num_layers=2)
