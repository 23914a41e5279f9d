"""GPU parity of kernel (b) and (c) through the public merge API (which calls the C ABI) against the
golden vectors produced by the unmodified reference, plus size-independent properties at the
VLMo-base size.  Tolerances: BASELINE.json asks 1e-6 relative for interpolation / arithmetic (we
require bit-exact fp32) and 1e-4 for RegMean (we require 1e-9: both sides are fp64)."""
import numpy as np
import pytest
import torch

import oracle
import vl_merging_b200 as vlm
from golden_io import MergeGolden

pytestmark = pytest.mark.gpu
G = MergeGolden()


def _t(d, device="cpu"):
    return {k: torch.from_numpy(np.array(v)).to(device) for k, v in d.items()}  # np.array keeps 0-dim shapes


def _run(vname, device_inputs):
    sd, cfg, grams = G.inputs(vname)
    method = G.variants[vname]["method"]
    dev = "cuda" if device_inputs else "cpu"
    tsd = _t(sd, dev)
    if method == "merge_weights":
        return vlm.merge_weights(tsd, cfg), tsd
    if method == "sum_task_vectors":
        return vlm.sum_task_vectors(tsd, cfg, central_weight={"state_dict": _t(G.central, dev)}), tsd
    return vlm.regmean(tsd, cfg, gram_matrices=_t(grams, dev)), tsd


@pytest.mark.parametrize("device_inputs", [False, True], ids=["host-inputs", "device-inputs"])
@pytest.mark.parametrize("vname", list(G.variants))
def test_merge_methods_match_reference_golden(vname, device_inputs):
    got, tsd = _run(vname, device_inputs)
    assert list(got.keys()) == G.variants[vname]["keys"]
    for k, w in G.expected(vname).items():
        g = got[k]
        assert g.device.type == ("cuda" if device_inputs else "cpu")
        g = g.cpu().numpy()
        assert g.dtype == w.dtype and g.shape == w.shape, k
        if w.dtype == np.float32:
            assert np.array_equal(g, w), k
        else:
            assert np.linalg.norm(g - w) / np.linalg.norm(w) < 1e-9, k
    for k, v in got.items():  # pass-through entries are the caller's own tensors
        if "transformer.blocks." not in k or "gamma" in k:
            assert v is tsd[k]


def test_merge_does_not_mutate_inputs():
    sd, cfg, _ = G.inputs("arith_l0.75")
    tsd, central = _t(sd), _t(G.central)
    before = {k: v.clone() for k, v in central.items()}
    vlm.sum_task_vectors(tsd, cfg, central_weight=central)
    assert all(torch.equal(before[k], central[k]) for k in central)  # the reference mutates its loaded copy
    assert all(torch.equal(torch.from_numpy(sd[k]), tsd[k]) for k in sd)


def test_file_based_config_like_the_reference(tmp_path):
    """central_weight / gram_matrices given as paths of torch.save'd files, as in the reference CLI."""
    from collections import defaultdict

    sd, cfg, grams = G.inputs("regmean_s0.9")
    gd = defaultdict(float)
    gd.update(_t(grams))
    torch.save(gd, tmp_path / "grams.pth")
    torch.save({"state_dict": _t(G.central)}, tmp_path / "central.pth")
    got = vlm.regmean(_t(sd), dict(cfg, gram_matrices=str(tmp_path / "grams.pth")))
    for k, w in G.expected("regmean_s0.9").items():
        assert np.linalg.norm(got[k].numpy() - w) / np.linalg.norm(w) < 1e-9
    sd, cfg, _ = G.inputs("arith_l0.75")
    got = vlm.Merger(dict(cfg, central_weight=str(tmp_path / "central.pth"), sum_task_vectors=True)).apply(_t(sd))
    for k, w in G.expected("arith_l0.75").items():
        assert np.array_equal(got[k].numpy(), w)


def test_singular_gram_sum_raises_like_torch_inverse():
    sd, cfg, grams = G.inputs("regmean_s1.0")
    bad = dict(grams)
    for k in bad:
        if k.startswith("transformer.blocks.0.mlp") and k.endswith("fc2"):
            bad[k] = np.zeros_like(bad[k])
    with pytest.raises(torch.linalg.LinAlgError):
        vlm.regmean(_t(sd), cfg, gram_matrices=_t(bad))


def test_cholesky_rejected_sum_is_solved_by_lu_like_torch_inverse():
    """The reference inverts the summed Gram with torch.inverse (LU, vilt_module.py:432,483): it returns a result for
    any numerically non-singular matrix.  A sum that is symmetric but NOT positive definite (one negative
    eigenvalue — what rounding can do to an ill-conditioned Gram sum) makes potrf fail; regmean then solves that
    problem with pivoted LU, warns, and still matches the oracle's explicit inverse."""
    sd, cfg, grams = G.inputs("regmean_s1.0")
    rng = np.random.default_rng(5)
    bad = dict(grams)
    hit = [k for k in bad if k.startswith("transformer.blocks.1.") and k.endswith(".proj")]
    assert hit
    d = bad[hit[0]].shape[0]
    q, _ = np.linalg.qr(rng.standard_normal((d, d)))          # same eigenvectors for both experts
    for k in hit:
        lam = np.linspace(1.0, 3.0, d)
        lam[d // 2] = -0.75 if k.endswith("attn.v.proj") else 0.25     # the SUM keeps one negative eigen-direction
        bad[k] = (q * lam) @ q.T
        bad[k] = (bad[k] + bad[k].T) / 2
    want = oracle.regmean(sd, bad, cfg)
    for streams in (1, 4):
        stats = {}
        with pytest.warns(RuntimeWarning, match="not positive definite"):
            got = vlm.regmean(_t(sd, "cuda"), cfg, gram_matrices=_t(bad, "cuda"), solve_streams=streams, stats=stats)
        assert stats["lu_fallbacks"] == 1
        for k, w in want.items():
            if "transformer.blocks." in k and "gamma" not in k:
                g = got[k].cpu().numpy()
                assert np.linalg.norm(g - w) <= 1e-9 * max(np.linalg.norm(w), 1e-30), k


@pytest.mark.parametrize("device_inputs", [False, True], ids=["host-inputs", "device-inputs"])
def test_load_time_dispatch_all_three_branches(device_inputs, tmp_path):
    """vilt_module.py:284-291: `merge_weights` / `sum_task_vectors` / `regmean` config switches select the method;
    Merger.apply is that dispatch on the CUDA path, checked against the reference goldens for each branch."""
    dev = "cuda" if device_inputs else "cpu"
    torch.save({"state_dict": _t(G.central)}, tmp_path / "central.pth")
    for vname, switch in (("interp_a0.5", "merge_weights"), ("arith_l0.75", "sum_task_vectors"), ("regmean_s0.9", "regmean")):
        sd, cfg, grams = G.inputs(vname)
        assert G.variants[vname]["method"] == switch
        extra = {switch: True}
        if switch == "sum_task_vectors":
            extra["central_weight"] = str(tmp_path / "central.pth")
        if switch == "regmean":
            from collections import defaultdict
            gd = defaultdict(float)
            gd.update(_t(grams))
            torch.save(gd, tmp_path / "grams.pth")
            extra["gram_matrices"] = str(tmp_path / "grams.pth")
        got = vlm.Merger(dict(cfg, **extra)).apply(_t(sd, dev))
        assert list(got.keys()) == G.variants[vname]["keys"]
        for k, w in G.expected(vname).items():
            g = got[k].cpu().numpy()
            if w.dtype == np.float32:
                assert np.array_equal(g, w), (vname, k)
            else:
                assert np.linalg.norm(g - w) / np.linalg.norm(w) < 1e-9, (vname, k)
    sd, cfg, _ = G.inputs("interp_a0.5")
    tsd = _t(sd, dev)
    assert vlm.Merger(dict(cfg)).apply(tsd) is tsd          # no switch set: the checkpoint is loaded as it is


@pytest.mark.parametrize("device_inputs", [False, True], ids=["host-inputs", "device-inputs"])
def test_regmean_concurrent_streams_equal_sequential(device_inputs):
    """The per-linear problems spread over several streams (default) vs one after the other: same kernels on the
    same data, bit-identical results; the singular case raises in both modes."""
    sd, cfg, grams = G.inputs("regmean_s0.9")
    dev = "cuda" if device_inputs else "cpu"
    seq = vlm.regmean(_t(sd, dev), cfg, gram_matrices=_t(grams, dev), solve_streams=1)
    for n in (2, 4, 7):
        con = vlm.regmean(_t(sd, dev), cfg, gram_matrices=_t(grams, dev), solve_streams=n)
        assert list(con) == list(seq)
        for k in seq:
            assert torch.equal(con[k], seq[k]), (n, k)
    bad = dict(grams)
    for k in bad:
        if k.startswith("transformer.blocks.0.mlp") and k.endswith("fc2"):
            bad[k] = np.zeros_like(bad[k])
    sd1, cfg1, _ = G.inputs("regmean_s1.0")
    for n in (1, 4):
        with pytest.raises(torch.linalg.LinAlgError):
            vlm.regmean(_t(sd1, dev), cfg1, gram_matrices=_t(bad, dev), solve_streams=n)
    ok = vlm.regmean(_t(sd, dev), cfg, gram_matrices=_t(grams, dev))        # and the library is healthy afterwards
    assert all(torch.equal(ok[k], seq[k]) for k in seq)


# ---- properties at the VLMo-base size (184 M expert parameters on the device) ---------------------

@pytest.fixture(scope="module")
def base_sd():
    cfg = vlm.vlmo_config("base")
    with torch.device("cuda"):
        model = vlm.VLMo(cfg)
    vlm.init_synthetic_(model, seed=1)
    return {k: v.detach() for k, v in model.state_dict().items()}


BASE_CFG = dict(vlffn_start_layer_index=10, only_activate_used_experts=True, merge_ratio=0.5, sum_lambda=0.75,
                loss_names={"irtr": 1, "vqa": 0, "nlvr2": 0})


def test_base_size_interpolation_properties(base_sd):
    stats = {}
    half = vlm.merge_weights(base_sd, BASE_CFG, stats=stats)
    assert stats["merge_bytes"] == 3 * 12 * 7_087_104 * 4  # 2 experts read + 1 written, 12 layers (IRTR-used experts)
    one = vlm.merge_weights(base_sd, dict(BASE_CFG, merge_ratio=1.0))
    zero = vlm.merge_weights(base_sd, dict(BASE_CFG, merge_ratio=0.0))
    for i in (0, 5, 11):
        for t in ("attn.qkv.weight", "mlp.fc2.weight", "norm1.bias"):
            k = f"transformer.blocks.{i}.{t}"
            v = base_sd[k.replace("attn.", "attn.v.").replace("mlp.", "mlp.v.").replace("norm1.", "norm1.v.")]
            l_ = base_sd[k.replace("attn.", "attn.l.").replace("mlp.", "mlp.l.").replace("norm1.", "norm1.l.")]
            assert torch.equal(one[k], v) and torch.equal(zero[k], l_)         # end points are exact copies
            assert torch.equal(half[k], 0.5 * v + 0.5 * l_)                    # same rounding as torch
    same = {k: v for k, v in base_sd.items()}
    for k in list(same):
        if ".l." in k:
            same[k] = same[k.replace(".l.", ".v.")]
    merged = vlm.merge_weights(same, dict(BASE_CFG, merge_ratio=0.3))
    k = "transformer.blocks.3.mlp.fc1.weight"
    want = np.float32(0.3) * same["transformer.blocks.3.mlp.v.fc1.weight"] + np.float32(0.7) * same["transformer.blocks.3.mlp.v.fc1.weight"]
    assert torch.equal(merged[k], want)


def test_base_size_matches_oracle_on_sampled_tensors(base_sd):
    """Oracle (numpy restatement) on a handful of full-size tensors; the other 150 are covered by properties."""
    central = {k.replace(".v.", "."): (v * 0.9 + 0.01) for k, v in base_sd.items() if ".v." in k and "blocks" in k}
    got = vlm.sum_task_vectors(base_sd, BASE_CFG, central_weight=central)
    sub = {}
    picks = ["transformer.blocks.2.attn.{m}qkv.weight", "transformer.blocks.11.mlp.{m}fc1.bias", "transformer.blocks.7.norm2.{m}weight"]
    np_sd = {k: v.cpu().numpy() for k, v in base_sd.items() if any(k == p.format(m=m) for p in picks for m in ("v.", "l.", "vl."))}
    np_central = {p.format(m=""): central[p.format(m="")].cpu().numpy() for p in picks}
    for p in picks:
        acc = np_central[p.format(m="")].copy()
        for m in ("v.", "l."):
            acc = acc + np.float32(0.75) * (np_sd[p.format(m=m)] - acc)   # oracle.sum_task_vectors' update
        assert np.array_equal(got[p.format(m="")].cpu().numpy(), acc), p


def test_randomised_segment_tables_through_the_c_abi():
    """Seeded sweep straight at vlm_merge_plan_*: random segment counts, sizes (incl. 0, 1, chunk boundaries),
    source counts, modes and misalignments; expected values from torch with the reference's rounding order."""
    import ctypes

    from vl_merging_b200 import _lib

    L = _lib.lib()
    rng = np.random.default_rng(7)
    for trial in range(12):
        nseg = int(rng.integers(1, 9))
        segs = (_lib.MergeSeg * nseg)()
        keep, expect = [], []
        for s in range(nseg):
            n = int(rng.choice([0, 1, 3, 4, 5, 4095, 4096, 4097, 8192, 50001]))
            n_src = int(rng.integers(1, 5))
            mode = int(rng.integers(0, 3))
            off = int(rng.choice([0, 0, 1, 2, 3]))
            srcs = [torch.randn(n + 8, device="cuda") for _ in range(n_src)]
            dst = torch.full((n + 8,), float("nan"), device="cuda")
            coefs = [float(c) for c in rng.uniform(-1.5, 1.5, size=n_src)]
            segs[s].dst = dst.data_ptr() + 4 * off
            for j in range(n_src):
                segs[s].src[j] = srcs[j].data_ptr() + 4 * off
                segs[s].coef[j] = coefs[j]
            segs[s].n, segs[s].n_src, segs[s].mode = n, n_src, mode
            xs = [t[off: off + n] for t in srcs]
            c32 = [torch.tensor(c, dtype=torch.float32, device="cuda") for c in coefs]
            if mode == _lib.MERGE_WSUM:
                acc = c32[0] * xs[0]
                for c, x in zip(c32[1:], xs[1:]):
                    acc = acc + c * x
            elif mode == _lib.MERGE_SEQ_LERP:
                acc = xs[0].clone()
                for c, x in zip(c32[1:], xs[1:]):
                    acc = acc + c * (x - acc)
            else:
                acc = xs[0].clone()
                for x in xs[1:]:
                    acc = acc + x
                acc = acc / torch.full_like(acc, float(n_src))   # true division (tensor / python scalar multiplies by 1/n on CUDA)
            keep.append((srcs, dst, off, n))
            expect.append(acc)
        plan = ctypes.c_void_p()
        _lib.check(L.vlm_merge_plan_create(segs, nseg, ctypes.byref(plan)))
        _lib.check(L.vlm_merge_plan_run(plan, None))
        torch.cuda.synchronize()
        assert L.vlm_merge_plan_bytes(plan) == sum((segs[s].n_src + 1) * segs[s].n * 4 for s in range(nseg))
        L.vlm_merge_plan_destroy(plan)
        for (srcs, dst, off, n), want in zip(keep, expect):
            assert torch.equal(dst[off: off + n], want), trial
            assert torch.isnan(dst[:off]).all() and torch.isnan(dst[off + n:]).all()   # nothing written out of range


def test_base_size_host_path_is_pipelined_and_bit_identical(base_sd):
    """CPU checkpoint in, CPU tensors out (the reference's calling convention) at the VLMo-base size: the
    executor overlaps H2D / kernel / D2H over groups of targets; every merged tensor must equal the
    device-resident result bit for bit."""
    host_sd = {k: v.cpu() for k, v in base_sd.items()}
    stats = {}
    got = vlm.merge_weights(host_sd, BASE_CFG, stats=stats)
    want = vlm.merge_weights(base_sd, BASE_CFG)
    assert stats.get("pipelined_groups", 0) >= 2
    assert stats["h2d_bytes"] == 2 * 12 * 7_087_104 * 4 and stats["d2h_bytes"] >= 12 * 7_087_104 * 4
    assert list(got.keys()) == list(want.keys())
    for k, w in want.items():
        if "transformer.blocks." in k and "gamma" not in k:
            assert got[k].device.type == "cpu" and torch.equal(got[k], w.cpu()), k
