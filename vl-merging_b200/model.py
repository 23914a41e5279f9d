"""model.py — the frozen VLMo multiway transformer in STOCK torch.

The Gram hooks need a forward to hang on, and the merged (ufo) weights need a forward to be checked
in; per the scope contract that forward stays stock PyTorch (no custom kernels here).  This module
mirrors the reference's module tree and parameter names, so that
  * a reference state_dict loads into it and vice versa (tests/test_model_vs_reference.py),
  * the module names the Gram-hook registration rule matches on are the reference's
    (src/cache_gram_matrices.py:264-281): `transformer.blocks.{i}.attn.{m}`, `...attn.{m}.proj`,
    `...mlp.{m}.fc1`, `...mlp.{m}.fc2`.

Reference it follows: src/vilt/modules/vision_transformer.py:272-363 (Mlp, Attention), :366-691
(Block: moe_forward / separate_plain_forward for type_id 0 and 1), :952-991 (visual_embed);
src/vilt/modules/vilt_module.py:122-186 (relative position index tables), :1226-1285
(infer_text_ft), :1378-1464 (infer_image_ft), :1071-1156 (infer: the fused vision-language route,
type_id 2, vision_transformer.py:495-523, :560-681).  The two fine-tuning towers (type_id 0 = image,
1 = text) serve IRTR calibration and evaluation; `infer` serves Gram caching for VQA / NLVR2 style
tasks, where the `vl` experts run and the shallow layers hand ROW SLICES of the joint sequence to the
`l` / `v` experts.  Task heads and losses are out of scope.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

MODALITIES = ("v", "l", "vl")


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Attention(nn.Module):
    """BEiT-style attention: fused qkv without its own bias, separate q/v biases, additive relative
    position bias, fp32 logits.  The qkv projection is applied with F.linear, not self.qkv(x) — as in
    the reference (vision_transformer.py:337), which is why the qkv Gram is hooked on THIS module."""

    def __init__(self, dim, num_heads, attn_impl="reference"):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.attn_impl = attn_impl  # "reference": the reference's explicit fp32 softmax; "sdpa": torch's fused kernel
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        self.q_bias = nn.Parameter(torch.zeros(dim))
        self.v_bias = nn.Parameter(torch.zeros(dim))
        self.proj = nn.Linear(dim, dim)

    def forward(self, x, mask=None, relative_position_bias=None):
        b, n, c = x.shape
        bias = torch.cat((self.q_bias, torch.zeros_like(self.v_bias), self.v_bias))
        qkv = F.linear(x, self.qkv.weight, bias).reshape(b, n, 3, self.num_heads, -1).permute(2, 0, 3, 1, 4)
        if self.attn_impl == "sdpa":
            # same math through F.scaled_dot_product_attention (still stock torch): the relative position
            # bias and the key-padding mask become one additive mask
            add = relative_position_bias.unsqueeze(0).to(qkv.dtype) if relative_position_bias is not None else None
            if mask is not None and not bool(mask.all()):
                pad = torch.zeros(b, 1, 1, n, dtype=qkv.dtype, device=x.device).masked_fill(
                    ~mask.bool()[:, None, None, :], float("-inf"))
                add = pad if add is None else add + pad
            x = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2], attn_mask=add, scale=self.scale)
            return self.proj(x.transpose(1, 2).reshape(b, n, c))
        q, k, v = qkv[0] * self.scale, qkv[1], qkv[2]
        attn = q.float() @ k.float().transpose(-2, -1)
        if relative_position_bias is not None:
            attn = attn + relative_position_bias.unsqueeze(0)
        if mask is not None:
            attn = attn.masked_fill(~mask.bool()[:, None, None, :], float("-inf"))
        attn = attn.softmax(dim=-1).type_as(x)
        x = (attn @ v).transpose(1, 2).reshape(b, n, c)
        return self.proj(x)


class Block(nn.Module):
    """One multiway block.  experts = ("v","l") / ("v","l","vl") gives the modality-specific (all_moe)
    layout with per-expert attention, MLP and both LayerNorms; experts = None the shared (ufo) one."""

    def __init__(self, dim, num_heads, mlp_ratio, experts, attn_impl="reference", joint=False, max_text_len=40):
        super().__init__()
        self.experts = experts
        self.joint = joint                # layer >= vlffn_start_layer_index: "vl" is among this layer's tasks
        self.max_text_len = max_text_len
        hidden = int(dim * mlp_ratio)
        ln = lambda: nn.LayerNorm(dim, eps=1e-6)  # noqa: E731
        if experts is None:
            self.attn, self.norm1 = Attention(dim, num_heads, attn_impl), ln()
            self.mlp, self.norm2 = Mlp(dim, hidden), ln()
        else:
            self.attn = nn.ModuleDict({m: Attention(dim, num_heads, attn_impl) for m in experts})
            self.norm1 = nn.ModuleDict({m: ln() for m in experts})
            self.mlp = nn.ModuleDict({m: Mlp(dim, hidden) for m in experts})
            self.norm2 = nn.ModuleDict({m: ln() for m in experts})
        self.gamma_1 = nn.Parameter(0.1 * torch.ones(dim))
        self.gamma_2 = nn.Parameter(0.1 * torch.ones(dim))

    def forward(self, x, mask, type_id, relative_position_bias):
        if type_id == 2 and not self.joint:
            return self._forward_split(x, mask, relative_position_bias)
        if self.experts is None:
            attn, norm1, mlp, norm2 = self.attn, self.norm1, self.mlp, self.norm2
        else:
            m = MODALITIES[type_id]
            attn, norm1, mlp, norm2 = self.attn[m], self.norm1[m], self.mlp[m], self.norm2[m]
        x = x + self.gamma_1 * attn(norm1(x), mask=mask, relative_position_bias=relative_position_bias)
        x = x + self.gamma_2 * mlp(norm2(x))
        return x

    def _forward_split(self, x, mask, bias):
        """type_id 2 on a layer WITHOUT a `vl` expert: text tokens [0, L) go through the `l` expert, image tokens
        [L, N) through the `v` expert, each attending only to its own modality (vision_transformer.py:510-516,
        :619-637, :667-677; shared weights: :540-556, :586-598).  As in the reference the LayerNorm output is
        concatenated first and the experts receive SLICES of it, so their forward hooks see non-contiguous
        (B, n, D) views — which GramCache reads in place (vlm_syrk_accum_strided)."""
        L = self.max_text_len
        moe = self.experts is not None
        pick = (lambda mod, m: mod[m]) if moe else (lambda mod, m: mod)
        h = torch.cat([pick(self.norm1, "l")(x[:, :L]), pick(self.norm1, "v")(x[:, L:])], dim=1)
        a_t = pick(self.attn, "l")(h[:, :L], mask=mask[:, :L], relative_position_bias=bias[:, :L, :L])
        a_i = pick(self.attn, "v")(h[:, L:], mask=mask[:, L:], relative_position_bias=bias[:, L:, L:])
        x = x + self.gamma_1 * torch.cat([a_t, a_i], dim=1)
        h = torch.cat([pick(self.norm2, "l")(x[:, :L]), pick(self.norm2, "v")(x[:, L:])], dim=1)
        x = x + self.gamma_2 * torch.cat([pick(self.mlp, "l")(h[:, :L]), pick(self.mlp, "v")(h[:, L:])], dim=1)
        return x


class PatchEmbed(nn.Module):
    def __init__(self, patch_size, dim):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x)


class Transformer(nn.Module):
    def __init__(self, cfg, experts_for_layer):
        super().__init__()
        dim = cfg["hidden_size"]
        self.patch_embed = PatchEmbed(cfg["patch_size"], dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.mask_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.blocks = nn.ModuleList(
            [Block(dim, cfg["num_heads"], cfg["mlp_ratio"], experts_for_layer(i), cfg.get("attn_impl", "reference"),
                   joint=i >= cfg["vlffn_start_layer_index"], max_text_len=cfg["max_text_len"])
             for i in range(cfg["num_layers"])])
        self.norm = nn.LayerNorm(dim, eps=1e-6)

    def visual_embed(self, img):
        x = self.patch_embed(img).flatten(2).transpose(1, 2)
        x = torch.cat((self.cls_token.expand(x.shape[0], -1, -1), x), dim=1)
        return x, torch.ones(x.shape[0], x.shape[1], device=x.device)


class TextEmbeddings(nn.Module):
    """word + token_type(0) + absolute position, LayerNorm(eps=1e-12): what HF BertEmbeddings computes
    in this image (transformers 5.x), which is what the reference instantiates (vilt_module.py:63)."""

    def __init__(self, vocab_size, dim, max_len):
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab_size, dim, padding_idx=0)
        self.position_embeddings = nn.Embedding(max_len, dim)
        self.token_type_embeddings = nn.Embedding(2, dim)
        self.LayerNorm = nn.LayerNorm(dim, eps=1e-12)

    def forward(self, ids):
        pos = torch.arange(ids.shape[1], device=ids.device)
        x = self.word_embeddings(ids) + self.token_type_embeddings(torch.zeros_like(ids))
        return self.LayerNorm(x + self.position_embeddings(pos)[None])


class _Head(nn.Module):
    def __init__(self, dim, bias):
        super().__init__()
        setattr(self, "fc" if not bias else "dense", nn.Linear(dim, dim, bias=bias))


DEFAULT_CONFIG = dict(
    hidden_size=768, num_heads=12, num_layers=12, mlp_ratio=4, image_size=384, patch_size=16,
    max_text_len=40, max_text_len_of_initckpt=196, vocab_size=30522, vlffn_start_layer_index=10,
    use_moe=True, attn_impl="reference",
)


def vlmo_config(name="base", **overrides):
    """'base' = task_finetune_irtr_coco_square_randaug_base_image384 (src/vilt/config.py:478-496),
    'large' = task_finetune_irtr_f30k_square_randaug_large_image384 (:454-475), 'tiny' = the
    vit_tiny_patch16_224 factory (vision_transformer.py:1260-1266) for CPU tests."""
    cfg = dict(DEFAULT_CONFIG)
    if name == "large":
        cfg.update(hidden_size=1024, num_heads=16, num_layers=24, vlffn_start_layer_index=21)
    elif name == "tiny":
        cfg.update(hidden_size=192, num_heads=3, image_size=224)
    elif name != "base":
        raise ValueError(name)
    cfg.update(overrides)
    return cfg


class VLMo(nn.Module):
    """ViLTransformerSS restricted to the IRTR fine-tuning towers.  cfg['use_moe'] True = all_moe
    (modality-specific experts), False = ufo (modality-agnostic, the merge target)."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = dict(cfg)
        dim, heads, layers = cfg["hidden_size"], cfg["num_heads"], cfg["num_layers"]
        self.num_layers = layers
        self.text_embeddings = TextEmbeddings(cfg["vocab_size"], dim, cfg["max_text_len"])
        self.token_type_embeddings = nn.Embedding(2, dim)
        vl0 = cfg["vlffn_start_layer_index"]
        if cfg["use_moe"]:
            experts = lambda i: ("v", "l") if i < vl0 else ("v", "l", "vl")  # noqa: E731
        else:
            experts = lambda i: None  # noqa: E731
        self.transformer = Transformer(cfg, experts)
        self.pooler = _Head(dim, bias=True)
        self.ifm_text_proj = _Head(dim, bias=False)
        self.ifm_image_proj = _Head(dim, bias=False)
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))

        # relative position bias: one table for all layers, [distances, heads*layers] (vilt_module.py:122-186)
        w = cfg["image_size"] // cfg["patch_size"]
        max_text, init_text = cfg["max_text_len"], cfg["max_text_len_of_initckpt"]
        n_img = (2 * w - 1) * (2 * w - 1) + 3
        n_all = n_img + 2 * init_text + 2
        self.relative_position_bias_table = nn.Parameter(torch.zeros(n_all, heads * layers))
        ys, xs = torch.meshgrid(torch.arange(w), torch.arange(w), indexing="ij")
        coords = torch.stack((ys.flatten(), xs.flatten()))
        rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0)
        idx = torch.zeros(w * w + 1, w * w + 1, dtype=torch.long)
        idx[1:, 1:] = (rel[..., 0] + w - 1) * (2 * w - 1) + rel[..., 1] + w - 1
        idx[0, :] = n_img - 3
        idx[:, 0] = n_img - 2
        idx[0, 0] = n_img - 1
        self.register_buffer("relative_position_index", idx)
        pos = torch.arange(max_text - 1)
        tidx = torch.zeros(max_text, max_text, dtype=torch.long)
        tidx[1:, 1:] = (pos[None, :] - pos[:, None]) - (2 - init_text) + n_img + 2
        tidx[0, :] = n_all - 3
        tidx[:, 0] = n_all - 2
        tidx[0, 0] = n_all - 1
        self.register_buffer("text_relative_position_index", tidx)
        # joint text+image index (float, as the reference builds it, vilt_module.py:176-186); only the fused
        # vl route reads it, but it is part of the checkpoint layout
        t2i = torch.full((max_text, w * w + 1), float(n_img))
        i2t = torch.full((w * w + 1, max_text), float(n_img + 1))
        self.register_buffer("text_imag_relative_position_index",
                             torch.cat((torch.cat((tidx.float(), t2i), 1), torch.cat((i2t, idx.float()), 1)), 0))

    # ---- forward (fine-tuning towers) ------------------------------------------------------------
    def _rel_pos_bias(self, index):
        bias = F.embedding(index, self.relative_position_bias_table).permute(2, 0, 1).contiguous()
        return torch.chunk(bias, self.num_layers, dim=0)

    def infer_text_ft(self, batch):
        ids, masks = batch["text_ids"], batch["text_masks"]
        x = self.text_embeddings(ids) + self.token_type_embeddings(torch.zeros_like(masks))
        biases = self._rel_pos_bias(self.text_relative_position_index)
        for i, blk in enumerate(self.transformer.blocks):
            x = blk(x, masks, 1, biases[i])
        feats = self.transformer.norm(x)
        cls = self.ifm_text_proj.fc(feats[:, 0])
        return {"text_feats": feats, "cls_feats": cls / cls.norm(dim=-1, keepdim=True), "raw_cls_feats": x[:, 0]}

    def infer_image_ft(self, batch, image_token_type_idx=1):
        img = batch["image"][0] if isinstance(batch["image"], (list, tuple)) else batch["image"]
        x, masks = self.transformer.visual_embed(img)
        masks = masks.long()
        x = x + self.token_type_embeddings(torch.full_like(masks, image_token_type_idx))
        biases = self._rel_pos_bias(self.relative_position_index)
        for i, blk in enumerate(self.transformer.blocks):
            x = blk(x, masks, 0, biases[i])
        feats = self.transformer.norm(x)
        cls = self.ifm_image_proj.fc(feats[:, 0])
        return {"image_feats": feats, "cls_feats": cls / cls.norm(dim=-1, keepdim=True), "raw_cls_feats": x[:, 0]}

    def infer(self, batch, image_token_type_idx=1):
        """The fused vision-language route (vilt_module.py:1071-1156): text and image tokens in ONE sequence,
        every block called with type_id 2 — `vl` experts on all 40 + 577 tokens where a layer has them, row
        slices through the `l` / `v` experts below vlffn_start_layer_index."""
        img = batch["image"][0] if isinstance(batch["image"], (list, tuple)) else batch["image"]
        ids, text_masks = batch["text_ids"], batch["text_masks"]
        text = self.text_embeddings(ids)
        image, image_masks = self.transformer.visual_embed(img)
        image_masks = image_masks.type_as(text_masks)
        text = text + self.token_type_embeddings(torch.zeros_like(text_masks))
        image = image + self.token_type_embeddings(torch.full_like(image_masks, image_token_type_idx))
        x = torch.cat([text, image], dim=1)
        masks = torch.cat([text_masks, image_masks], dim=1)
        biases = self._rel_pos_bias(self.text_imag_relative_position_index.long())
        for i, blk in enumerate(self.transformer.blocks):
            x = blk(x, masks, 2, biases[i])
        x = self.transformer.norm(x)
        n_text = text.shape[1]
        return {"text_feats": x[:, :n_text], "image_feats": x[:, n_text:],
                "cls_feats": torch.tanh(self.pooler.dense(x[:, 0])), "raw_cls_feats": x[:, 0]}

    def forward(self, batch):
        """One IRTR calibration step (objectives.py:372-379): both towers, similarity logits."""
        img = self.infer_image_ft(batch)["cls_feats"]
        txt = self.infer_text_ft(batch)["cls_feats"]
        return self.logit_scale.exp() * img @ txt.t()


# ---- deterministic synthetic weights ----------------------------------------------------------------

def _hash_uniform(numel, seed, device="cpu"):
    """Uniform (-0.5, 0.5) from an integer hash of the element index: exact on every machine
    (no dependence on torch's RNG stream or CPU vector width)."""
    i = torch.arange(numel, dtype=torch.int64, device=device)
    x = ((i + 1) * 0x9E3779B1 + (seed % (1 << 24)) * 0x85EBCA77 + 0x165667B1) % (1 << 32)
    x = (x ^ (x >> 15)) * 0x2C1B3C6D % (1 << 32)
    x = (x ^ (x >> 12)) * 0x297A2D39 % (1 << 32)
    x = x ^ (x >> 15)
    return ((x % (1 << 24)).to(torch.float32) + 0.5) / float(1 << 24) - 0.5


@torch.no_grad()
def init_synthetic_(model, seed=1):
    """Random-init stand-in for a trained checkpoint (there is no network for real ones): every tensor
    gets hash-uniform noise at the scale the reference initialises it with; LayerNorm weights ~ 1,
    layer-scale gammas ~ 0.1, relative position bias small."""
    import zlib

    for name, p in model.named_parameters():  # parameters only: index buffers are functions of the config
        # per-tensor stream keyed by the NAME, so adding or removing tensors never reshuffles the others
        u = _hash_uniform(p.numel(), seed * 1000003 + zlib.crc32(name.encode()) % 999983, device=p.device).reshape(p.shape)
        if "norm" in name.lower() and name.endswith("weight"):
            p.copy_(1.0 + 0.2 * u)
        elif "gamma_" in name:
            p.copy_(0.1 + 0.05 * u)
        elif name == "logit_scale":
            p.fill_(math.log(1 / 0.07))
        elif name.endswith("bias") or "cls_token" in name or "mask_token" in name:
            p.copy_(0.1 * u)
        elif name == "relative_position_bias_table":
            p.copy_(0.5 * u)
        else:
            p.copy_(0.07 * u)  # std ~ 0.02, like trunc_normal_(std=0.02)
    return model


def synthetic_batch(batch_size, cfg, seed=0, device="cpu", dtype=torch.float32, pad=False):
    """Synthetic calibration batch in the reference's collate layout (base_dataset.py:204-253):
    image U(-1,1) [B,3,H,W], text_ids with [CLS]=101 first, text_masks, text_labels = -100."""
    n = batch_size * 3 * cfg["image_size"] ** 2
    image = (2.0 * _hash_uniform(n, 7919 + seed)).reshape(batch_size, 3, cfg["image_size"], cfg["image_size"])
    t = cfg["max_text_len"]
    ids = ((_hash_uniform(batch_size * t, 104729 + seed) + 0.5) * (cfg["vocab_size"] - 1000)).long() + 999
    ids = ids.reshape(batch_size, t)
    ids[:, 0] = 101
    masks = torch.ones(batch_size, t, dtype=torch.long)
    if pad:  # ragged lengths >= 8; padded positions keep id 0 and still count in the Grams
        lens = 8 + ((_hash_uniform(batch_size, 31 + seed) + 0.5) * (t - 8)).long()
        masks = (torch.arange(t)[None, :] < lens[:, None]).long()
        ids = ids * masks
    return {
        "image": [image.to(device=device, dtype=dtype)],
        "text_ids": ids.to(device),
        "text_masks": masks.to(device),
        "text_labels": torch.full((batch_size, t), -100, dtype=torch.long, device=device),
    }
