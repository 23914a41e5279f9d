"""IRTR recall against the UNMODIFIED reference function `compute_irtr_recall`
(src/vilt/modules/objectives.py:572-710), executed in the build container (marker `reference`) on a fake data module:
tiny datasets whose items carry their own index, and tower stand-ins that look features up by that index — so the
reference's loaders, preload loops, similarity matrix, top-k and recall arithmetic all run as written, and both the
oracle (`oracle.irtr_recall`) and the product function (`vlm.irtr_recall`) must reproduce its six numbers."""
import os
import types

import numpy as np
import pytest
import torch

import oracle
import vl_merging_b200 as vlm


class _Items(torch.utils.data.Dataset):
    """Captions (image_only=False: one item per caption) or images (image_only=True)."""

    def __init__(self, img_index, image_only):
        self.img_index, self.image_only = list(img_index), image_only

    def __len__(self):
        return len(self.img_index)

    def __getitem__(self, i):
        return {"pos": i, "img_index": self.img_index[i]}

    def collate(self, batch, mlm_collator=None):
        pos = torch.tensor([b["pos"] for b in batch])
        out = {"img_index": [b["img_index"] for b in batch]}
        if self.image_only:
            out["image"] = [pos.float().view(-1, 1, 1, 1).expand(-1, 3, 2, 2).contiguous()]
        else:
            ids = torch.zeros(len(batch), 4, dtype=torch.long)
            ids[:, 0] = pos
            out.update(text_ids=ids, text_masks=torch.ones_like(ids), text_labels=torch.full_like(ids, -100))
        return out


@pytest.mark.reference
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_recall_matches_the_reference_function(seed):
    import torch.distributed as dist

    import ref_harness as rh

    rh.import_reference()
    import vilt.modules.objectives as objectives

    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", str(29600 + seed))
        dist.init_process_group("gloo", rank=0, world_size=1)
        created = True
    try:
        rng = np.random.default_rng(seed)
        n_img, per_img, dim = 37, 3, 16
        iids = rng.permutation(1000)[:n_img].tolist()                     # arbitrary image ids
        tiids = [iids[i] for i in range(n_img) for _ in range(per_img)]
        order = rng.permutation(len(tiids))
        tiids = [tiids[i] for i in order]
        img_feats = torch.from_numpy(rng.standard_normal((n_img, dim)).astype(np.float32))
        # captions near their image (so recalls are neither 0 nor 1) plus noise
        base = {iid: img_feats[k] for k, iid in enumerate(iids)}
        txt_feats = torch.stack([base[t] for t in tiids]) + 1.5 * torch.from_numpy(
            rng.standard_normal((len(tiids), dim)).astype(np.float32))
        img_feats = img_feats / img_feats.norm(dim=-1, keepdim=True)
        txt_feats = txt_feats / txt_feats.norm(dim=-1, keepdim=True)

        dm = types.SimpleNamespace(
            tokenizer=None, mlm_collator=None,
            make_no_false_test_dset=lambda image_only=False: _Items(iids if image_only else tiids, image_only))
        module = types.SimpleNamespace(
            trainer=types.SimpleNamespace(datamodule=types.SimpleNamespace(dms=[dm])),
            device=torch.device("cpu"),
            infer_text_ft=lambda b: {"cls_feats": txt_feats[b["text_ids"][:, 0]]},
            infer_image_ft=lambda b: {"cls_feats": img_feats[b["image"][0][:, 0, 0, 0].long()]})
        ref = [float(v) for v in objectives.compute_irtr_recall(module, split="test")]
    finally:
        if created:
            dist.destroy_process_group()

    _, want = oracle.irtr_recall(img_feats.numpy(), txt_feats.numpy(), iids, tiids)
    assert [float(v) for v in want] == pytest.approx(ref, abs=1e-7)
    _, got = vlm.irtr_recall(img_feats, txt_feats, iids, tiids)
    assert [float(v) for v in got] == pytest.approx(ref, abs=1e-7)
    assert 0.0 < ref[2] < 1.0 or 0.0 < ref[5] < 1.0        # the case is not degenerate
