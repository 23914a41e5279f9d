"""IRTR evaluation (vl-merging_b200/irtr.py) against the numpy restatement of objectives.py:684-710 and,
end to end, against the reference-merged model's golden scores (tests/golden/model_tiny.npz)."""
import json
import os

import numpy as np
import torch

import oracle
import vl_merging_b200 as vlm

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model_tiny.npz")


def test_recall_matches_oracle():
    rng = np.random.default_rng(3)
    ni, per, d = 40, 5, 32
    img = rng.standard_normal((ni, d)).astype(np.float32)
    txt = (np.repeat(img, per, axis=0) + 1.5 * rng.standard_normal((ni * per, d))).astype(np.float32)  # noisy captions
    iids, tiids = np.arange(ni), np.repeat(np.arange(ni), per)
    scores, got = vlm.irtr_recall(torch.from_numpy(img), torch.from_numpy(txt), iids, tiids)
    want_scores, want = oracle.irtr_recall(img, txt, iids, tiids)
    assert np.allclose(scores.numpy(), want_scores, rtol=1e-5, atol=1e-5)
    assert np.allclose([float(g) for g in got], [float(w) for w in want], atol=1e-7)
    assert 0.0 < float(got[0]) < 1.0 and float(got[2]) >= float(got[1]) >= float(got[0])   # r@1 <= r@5 <= r@10


def test_features_and_scores_of_reference_merged_model():
    """CPU leg of config 4: oracle-merged weights (bit-identical to the reference's, tests/test_oracle.py) in our
    stock-torch ufo model reproduce the reference ufo model's features and similarity matrix."""
    z = np.load(GOLDEN)
    meta = json.loads(bytes(z["meta"]).decode())
    cfg = vlm.vlmo_config("tiny")
    src = vlm.init_synthetic_(vlm.VLMo(cfg).eval(), seed=1)
    sd = {k: v.numpy() for k, v in src.state_dict().items()}
    mcfg = dict(vlffn_start_layer_index=10, only_activate_used_experts=False, merge_ratio=0.5,
                loss_names={"irtr": 1.0, "vqa": 0, "nlvr2": 0})
    merged = {k: torch.from_numpy(np.asarray(v)) for k, v in oracle.merge_weights(sd, mcfg).items()}
    ufo = vlm.VLMo(vlm.vlmo_config("tiny", use_moe=False)).eval()
    ufo.load_state_dict(merged, strict=False)
    (ni, si), (nt, st) = meta["eval"]
    img, txt = vlm.irtr_features(ufo, [vlm.synthetic_batch(ni, cfg, seed=si)], [vlm.synthetic_batch(nt, cfg, seed=st, pad=True)])
    scores, recalls = vlm.irtr_recall(img, txt, np.arange(ni), np.arange(nt) % ni)
    assert np.abs(scores.numpy() - z["merged/interp/scores"]).max() < 1e-5
    assert len(recalls) == 6
