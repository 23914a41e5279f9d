"""Host logic of the merge methods (vl-merging_b200/plan.py) against the reference's golden outputs:
the MergeOps are executed here by a tiny numpy interpreter, so this runs without a GPU."""
import numpy as np
import pytest

import vl_merging_b200 as vlm
from vl_merging_b200 import plan as P
from golden_io import MergeGolden

G = MergeGolden()


def interpret(ops, sd, central=None, grams=None, alpha=None):
    out = {k: v for k, v in sd.items() if P.is_passthrough_key(k)}
    for op in ops:
        if op.passthrough:
            out[op.dst] = sd[op.passthrough]
        elif op.regmean is not None:
            summed, acc = 0, 0
            for wk, gk in op.regmean:
                g = np.asarray(grams[gk], np.float64)
                g = alpha * g + (1 - alpha) * np.diag(np.diag(g))
                summed = summed + g
                acc = acc + sd[wk].astype(np.float64) @ g
            out[op.dst] = acc @ np.linalg.inv(summed)
        elif op.mode == P.WSUM:
            acc = np.float32(op.coefs[0]) * sd[op.srcs[0]]
            for c, k in zip(op.coefs[1:], op.srcs[1:]):
                acc = acc + np.float32(c) * sd[k]
            out[op.dst] = acc
        elif op.mode == P.SEQ_LERP:
            acc = central[op.dst].copy()
            for c, k in zip(op.coefs, op.srcs):
                acc = acc + np.float32(c) * (sd[k] - acc)
            out[op.dst] = acc
        else:
            acc = sd[op.srcs[0]]
            for k in op.srcs[1:]:
                acc = acc + sd[k]
            out[op.dst] = acc / np.float32(len(op.srcs))
    return out


@pytest.mark.parametrize("vname", list(G.variants))
def test_plan_reproduces_reference(vname):
    sd, cfg, grams = G.inputs(vname)
    method = G.variants[vname]["method"]
    if method == "merge_weights":
        ops = P.plan_merge_weights(sd.keys(), cfg)
        got = interpret(ops, sd)
    elif method == "sum_task_vectors":
        ops = P.plan_sum_task_vectors(sd.keys(), G.central.keys(), cfg)
        got = interpret(ops, sd, central=G.central)
    else:
        ops = P.plan_regmean(sd.keys(), grams.keys(), cfg)
        got = interpret(ops, sd, grams=grams, alpha=cfg["scaling_for_non_diag"])
    assert list(got.keys()) == G.variants[vname]["keys"]
    for k, w in G.expected(vname).items():
        if w.dtype == np.float32:
            assert np.array_equal(got[k], w), k
        else:
            assert np.linalg.norm(got[k] - w) / np.linalg.norm(w) < 1e-9, k


def test_missing_everything_is_a_keyerror_like_the_reference():
    sd, cfg, _ = G.inputs("interp_a0.5")
    broken = {k: v for k, v in sd.items() if ".blocks.5.mlp.l.fc1.weight" not in k}
    with pytest.raises(KeyError):
        P.plan_merge_weights(broken.keys(), cfg)


def test_only_used_experts_without_a_task_fails_like_the_reference():
    sd, cfg, _ = G.inputs("interp_a0.5")
    cfg = dict(cfg, only_activate_used_experts=True, loss_names=dict(cfg["loss_names"], irtr=0))
    with pytest.raises(TypeError):
        P.plan_merge_weights(sd.keys(), cfg)


def test_regmean_without_any_gram_raises():
    sd, cfg, grams = G.inputs("regmean_s1.0")
    with pytest.raises(KeyError):
        P.plan_regmean(sd.keys(), {}, cfg)


def test_vit_large_layer_count():
    """the reference hard-codes range(12); ViT-L callers pass num_layers=24."""
    keys = set()
    for i in range(24):
        for m in (["v", "l"] if i < 21 else ["v", "l", "vl"]):
            for src, _ in P.layer_targets(i):
                keys.add(src.replace("{m}", m))
    cfg = dict(vlffn_start_layer_index=21, only_activate_used_experts=False, merge_ratio=0.5,
               loss_names={"irtr": 1, "vqa": 0, "nlvr2": 0})
    assert len(P.plan_merge_weights(keys, cfg, num_layers=12)) == 12 * 13
    ops = P.plan_merge_weights(keys, cfg, num_layers=24)
    assert len(ops) == 24 * 13
    assert len(ops[-1].srcs) == 3 and abs(sum(ops[-1].coefs) - 1) < 1e-12


def test_package_imports_without_cuda():
    assert vlm.__version__
    assert callable(vlm.merge_weights) and callable(vlm.GramCache)


def test_independent_unshares_views_of_one_host_buffer():
    import torch

    from vl_merging_b200.merge import independent

    flat = torch.arange(12, dtype=torch.float32)
    sd = {"a": flat[:4].view(2, 2), "b": flat[4:].view(2, 4), "own": torch.ones(3), "n": 5}
    out = independent(sd)
    assert list(out) == list(sd) and out["n"] == 5 and out["own"] is sd["own"]
    for k in ("a", "b"):
        assert torch.equal(out[k], sd[k]) and out[k].untyped_storage().data_ptr() != flat.untyped_storage().data_ptr()
        assert out[k].untyped_storage().nbytes() == out[k].numel() * 4


def test_device_resolution_follows_local_rank(monkeypatch):
    import torch

    from vl_merging_b200 import merge

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "device_count", lambda: 8)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    cpu_sd = {"w": torch.zeros(2)}
    monkeypatch.setenv("LOCAL_RANK", "5")
    assert merge._resolve_device(cpu_sd, None) == torch.device("cuda", 5)      # not everybody on cuda:0
    monkeypatch.setenv("LOCAL_RANK", "11")                                      # out of range: ignore
    assert merge._resolve_device(cpu_sd, None) == torch.device("cuda", 0)
    monkeypatch.delenv("LOCAL_RANK")
    assert merge._resolve_device(cpu_sd, None) == torch.device("cuda", 0)
    assert merge._resolve_device(cpu_sd, "cuda:3") == torch.device("cuda", 3)   # an explicit device always wins
