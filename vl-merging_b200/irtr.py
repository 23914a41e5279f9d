"""irtr.py — the step right after the merge: COCO-style image<->text retrieval of a
(merged) model, objectives.compute_irtr_recall (src/vilt/modules/objectives.py:572-710) without its data
loading.  Stock torch: the two towers run over all texts and all images (optionally under autocast, as the
reference does, :657,669), `scores = img_cls_feats @ txt_cls_feats.t()` (:684), top-1/5/10 both ways (:688-708).

The reference evaluates the full set on every rank; here a process group shards the batches over the ranks
and all-gathers the 768-d features (one exchange step), after which every rank holds the same score matrix.
"""
import torch

from . import _lib


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


def _split3(x, left):
    """fp32 features as three bf16 planes (x = p1 + p2 + p3 to 24 bits) laid side by side so that ONE bf16 product
    of the widened matrices is the fp32-accurate score: [a1 a1 a2 a1 a3 a2] . [b1 b2 b1 b3 b1 b2] = a1b1 + a1b2 +
    a2b1 + a1b3 + a3b1 + a2b2 (largest terms first; the dropped ones are below 2^-24 of the score)."""
    p1 = x.bfloat16()
    r = x - p1.float()
    p2 = r.bfloat16()
    p3 = (r - p2.float()).bfloat16()
    order = (p1, p1, p2, p1, p3, p2) if left else (p1, p2, p1, p3, p1, p2)
    return torch.cat(order, dim=1)


def _fusable(a, b):
    """Operands for vlm_sim_topk: fp16 / bf16 CUDA features as they are; fp32 features that fp16 holds exactly
    (produced under autocast and up-cast afterwards) narrowed back losslessly; any other fp32 features as 3-way
    bf16 splits (6 x the tensor work, fp32 accuracy)."""
    if not (a.is_cuda and b.is_cuda and a.dim() == 2 and b.dim() == 2 and a.shape[1] == b.shape[1]) or a.shape[1] % 8:
        return None
    if a.dtype == b.dtype and a.dtype in (torch.float16, torch.bfloat16):
        return a.contiguous(), b.contiguous()
    if a.dtype == torch.float32 and b.dtype == torch.float32:
        ha, hb = a.half(), b.half()
        if torch.equal(ha.float(), a) and torch.equal(hb.float(), b):
            return ha.contiguous(), hb.contiguous()
        return _split3(a, True), _split3(b, False)
    return None


def sim_topk(a, b, k=10):
    """Row-wise top-k of a @ b.T without forming it (vlm_sim_topk: TMA + tcgen05 kind::f16, running top-10 in the
    accumulator epilogue).  a [m, d], b [n, d]: fp16 or bf16 CUDA tensors.  Returns (values [m, k] fp32 descending,
    indices [m, k] int64); among equal scores the lower index comes first (a stable descending sort)."""
    if not (1 <= k <= 10):
        raise ValueError("sim_topk keeps the ten best per row: 1 <= k <= 10")
    if a.dtype != b.dtype or a.dtype not in (torch.float16, torch.bfloat16) or not a.is_cuda or a.device != b.device:
        raise RuntimeError("sim_topk needs fp16 / bf16 features on one CUDA device (there is no CPU fallback)")
    lib = _lib.lib()
    a, b = a.contiguous(), b.contiguous()
    m, n, d = a.shape[0], b.shape[0], a.shape[1]
    with torch.cuda.device(a.device):
        splits = lib.vlm_sim_topk_splits(m, n)
        val = torch.empty(m, splits, 10, dtype=torch.float32, device=a.device)
        idx = torch.empty(m, splits, 10, dtype=torch.int32, device=a.device)
        code = _lib.VLM_F16 if a.dtype == torch.float16 else _lib.VLM_BF16
        _lib.check(lib.vlm_sim_topk(a.data_ptr(), m, a.stride(0), b.data_ptr(), n, b.stride(0), d, code, val.data_ptr(),
                                    idx.data_ptr(), splits, torch.cuda.current_stream(a.device).cuda_stream))
    val, idx = val.view(m, splits * 10), idx.view(m, splits * 10)
    if splits > 1:   # partial lists are in increasing column order: a stable sort keeps the lower index first
        val, order = torch.sort(val, dim=1, descending=True, stable=True)
        idx = torch.gather(idx, 1, order)
    return val[:, :k].contiguous(), idx[:, :k].long()


def irtr_recall_fused(img_cls_feats, txt_cls_feats, iids, tiids):
    """objectives.py:684-710 without the score matrix: two vlm_sim_topk launches (per image over captions, per
    caption over images) and the recall arithmetic on the 10 indices per row.  Features: fp16 / bf16 as they are;
    fp32 (what infer_*_ft return even under autocast: the normalisation promotes) as 3-way bf16 splits, which keeps
    fp32 accuracy.  Returns ((ir_r1, ir_r5, ir_r10, tr_r1, tr_r5, tr_r10), (top captions per image [Ni, 10],
    top images per caption [Nt, 10]))."""
    pair = _fusable(img_cls_feats, txt_cls_feats)
    if pair is None:
        raise RuntimeError("irtr_recall_fused needs 2-D CUDA features of one floating dtype (fp16 / bf16 / fp32) with a "
                           "width that is a multiple of 8; use irtr_recall for anything else")
    img, txt = pair
    iids = torch.as_tensor(iids, device=img.device)
    tiids = torch.as_tensor(tiids, device=img.device)
    _, by_image = sim_topk(img, txt, min(10, txt.shape[0]))
    if img.shape[1] != img_cls_feats.shape[1]:     # split operands: the roles of the two plane orders swap too
        txt, img = _split3(txt_cls_feats, True), _split3(img_cls_feats, False)
    _, by_caption = sim_topk(txt, img, min(10, img.shape[0]))
    out = {}
    for k in (1, 5, 10):
        top = by_image[:, :k]
        out[f"tr_r{k}"] = (iids.unsqueeze(1) == tiids[top]).float().max(dim=1)[0].mean()
        top = by_caption[:, :k]
        out[f"ir_r{k}"] = (tiids.unsqueeze(1) == iids[top]).float().max(dim=1)[0].mean()
    return (out["ir_r1"], out["ir_r5"], out["ir_r10"], out["tr_r1"], out["tr_r5"], out["tr_r10"]), (by_image, by_caption)


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
