"""irtr.py — the step right after the merge: COCO-style image<->text retrieval of a
(merged) model, objectives.compute_irtr_recall (src/vilt/modules/objectives.py:572-710) without its data
loading.  Stock torch: the two towers run over all texts and all images (optionally under autocast, as the
reference does, :657,669), `scores = img_cls_feats @ txt_cls_feats.t()` (:684), top-1/5/10 both ways (:688-708).

The reference evaluates the full set on every rank; here a process group shards the batches over the ranks
and all-gathers the 768-d features (one exchange step), after which every rank holds the same score matrix.
"""
import torch


@torch.no_grad()
def irtr_features(model, image_batches, text_batches, autocast_dtype=None, group=None):
    """Runs infer_image_ft over image_batches and infer_text_ft over text_batches (lists of batch dicts in the
    reference's collate layout) and returns (img_cls_feats [Ni, D], txt_cls_feats [Nt, D])."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if group is not None else 1
    rank = dist.get_rank(group) if group is not None else 0

    def run(batches, fn):
        feats = []
        for i, b in enumerate(batches):
            if i % world != rank:
                continue
            if autocast_dtype is None:
                feats.append(fn(b)["cls_feats"])
            else:
                with torch.autocast("cuda", dtype=autocast_dtype):
                    feats.append(fn(b)["cls_feats"])
        if world == 1:
            return torch.cat(feats)
        # one all-gather of equal-width, zero-padded shards; batch i came from rank i % world, in order
        dev = next(model.parameters()).device
        width = model.cfg["hidden_size"]
        local = torch.cat(feats).float() if feats else torch.zeros(0, width, device=dev)
        counts = torch.zeros(world, len(batches), dtype=torch.int64, device=dev)
        for j, f in zip([i for i in range(len(batches)) if i % world == rank], feats):
            counts[rank, j] = f.shape[0]
        dist.all_reduce(counts, group=group)
        rows = int(counts.sum(1).max().item())
        padded = torch.zeros(rows, width, device=dev)
        padded[: local.shape[0]] = local
        gathered = torch.empty(world * rows, width, device=dev)
        dist.all_gather_into_tensor(gathered, padded, group=group)
        out, cursor = [], [0] * world
        for i in range(len(batches)):
            r, n = i % world, int(counts[i % world, i].item())
            out.append(gathered[r * rows + cursor[r]: r * rows + cursor[r] + n])
            cursor[r] += n
        return torch.cat(out)

    txt = run(text_batches, model.infer_text_ft)
    img = run(image_batches, model.infer_image_ft)
    return img, txt


def irtr_recall(img_cls_feats, txt_cls_feats, iids, tiids):
    """objectives.py:684-710.  iids [Ni]: image index of every image; tiids [Nt]: image index of every caption.
    Returns (scores, (ir_r1, ir_r5, ir_r10, tr_r1, tr_r5, tr_r10)) — tr = text retrieval (per image, over
    captions), ir = image retrieval (per caption, over images)."""
    scores = img_cls_feats @ txt_cls_feats.t()
    iids = torch.as_tensor(iids, device=scores.device)
    tiids = torch.as_tensor(tiids, device=scores.device)
    out = {}
    for k in (1, 5, 10):
        # (k is clamped for toy-sized sets; the reference's topk would raise there)
        top = scores.topk(min(k, scores.shape[1]), dim=1).indices   # per image: best captions
        out[f"tr_r{k}"] = (iids.unsqueeze(1) == tiids[top]).float().max(dim=1)[0].mean()
        top = scores.topk(min(k, scores.shape[0]), dim=0).indices   # per caption: best images
        out[f"ir_r{k}"] = (tiids.unsqueeze(0) == iids[top]).float().max(dim=0)[0].mean()
    return scores, (out["ir_r1"], out["ir_r5"], out["ir_r10"], out["tr_r1"], out["tr_r5"], out["tr_r10"])
