"""CPU ORACLE - TEST INFRASTRUCTURE ONLY (never imported by the product path).

A plain-PyTorch functional restatement of the OA-Transformer video-text dual-encoder hot path, written from the
reference's algorithm (file:line cited per function, paths relative to /root/reference/OATrans). It is the checker
for the CUDA path: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs are the only
importers.

Pinning (see oracle/make_golden.py and tests/test_oracle_vs_golden.py): the reference ships no golden vectors or
known-answer tests, so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, generated
in the authoring container by importing the unmodified reference modules (oracle/ref_shim.py) and committed as
tests/golden/*.pt. Rows of SURVEY.md section 8a marked "spec-by-extension" (object tokens X1-X3, object->patch
attention X4) have no runnable reference; there this file IS the specification ("parity unpinned" for those rows,
except the fragments that do run: the bbox->patch-mask bookkeeping and the einsum pooling / sigmoid similarity).

Third-party arithmetic: the text tower is HuggingFace DistilBERT (reference pins transformers==4.6.0,
environment.yml:115; this container has 5.5.0). `distilbert` below restates the published architecture and is pinned
against the installed DistilBertModel run through the reference's own call site (model/oa_model.py:110-115).

Two arithmetic modes:
  cfg.bf16 = False : fp32 everywhere (what the reference computes).
  cfg.bf16 = True  : every matrix-product OPERAND is rounded to bf16 (fp32 accumulate); residual stream, LayerNorm,
                     softmax, loss stay fp32; dgrad results are rounded to bf16. This mirrors the storage points of
                     the CUDA path exactly, so it isolates accumulation-order differences (gate 1e-3).
  cfg.split = True : (with bf16) the rows the logits depend on DIRECTLY - the CLS row of every video-tower linear, every
                     text-tower linear, both projections - use split-bf16 operands in the forward pass: x = hi + lo
                     with hi = bf16(x), lo = bf16(x - hi), same for w, product = hi.hi + hi.lo + lo.hi (three bf16
                     MMAs, fp32 accumulate; relative error ~2^-17 instead of 2^-9). The other F*n token rows reach
                     the CLS row only through attention averages over hundreds of keys, where operand rounding noise
                     averages out. The backward pass is the plain bf16 one. This is the CUDA path's contract.
"""
import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class OracleCfg:
    heads: int = 12
    bf16: bool = False
    split: bool = True              # split-bf16 (3-term) forward products on the CLS / text / projection rows (bf16 mode)
    ln_eps_video: float = 1e-6      # video_transformer.py:228
    ln_eps_text: float = 1e-12      # HF DistilBERT Embeddings / TransformerBlock LayerNorm
    text_layers: int = 6
    patch: int = 16
    modality_token: bool = False    # oa_video_transformer_region.py:257-261


# ----------------------------------------------------------------------------------------------------------------
# bf16 operand emulation
# ----------------------------------------------------------------------------------------------------------------
def _r(x):
    return x.to(torch.bfloat16).to(torch.float32)


class _RoundSTE(torch.autograd.Function):
    """bf16 rounding with a straight-through gradient (storage point in the CUDA path)."""

    @staticmethod
    def forward(ctx, x):
        return _r(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _LinearBF16(torch.autograd.Function):
    """y = round(x) @ round(w)^T (+ b); backward consumes a bf16-rounded dy, returns bf16-rounded dx, fp32 dw/db."""

    @staticmethod
    def forward(ctx, x, w, b):
        xr, wr = _r(x), _r(w)
        ctx.save_for_backward(xr, wr)
        ctx.has_bias = b is not None
        y = xr @ wr.t()
        return y + b if b is not None else y

    @staticmethod
    def backward(ctx, g):
        xr, wr = ctx.saved_tensors
        gr = _r(g)
        dx = _r(gr @ wr)
        g2 = gr.reshape(-1, gr.shape[-1])
        dw = g2.t() @ xr.reshape(-1, xr.shape[-1])
        db = g2.sum(0) if ctx.has_bias else None
        return dx, dw, db


class _LinearSplit(torch.autograd.Function):
    """Forward with split-bf16 operands: (x_hi + x_lo)(w_hi + w_lo)^T without the lo.lo term; backward as _LinearBF16."""

    @staticmethod
    def forward(ctx, x, w, b):
        xh, wh = _r(x), _r(w)
        xl, wl = _r(x - xh), _r(w - wh)
        ctx.save_for_backward(xh, wh)
        ctx.has_bias = b is not None
        y = xh @ wh.t() + (xh @ wl.t() + xl @ wh.t())
        return y + b if b is not None else y

    backward = staticmethod(_LinearBF16.backward)


def linear(x, w, b, cfg, split=False):
    if cfg.bf16:
        return (_LinearSplit if (split and cfg.split) else _LinearBF16).apply(x, w, b)
    return F.linear(x, w, b)


def linear_cls(x, w, b, cfg):
    """Linear over (B, T, D) tokens whose CLS row (token 0) takes the split-bf16 product (see the module docstring)."""
    if not (cfg.bf16 and cfg.split):
        return linear(x, w, b, cfg)
    return torch.cat([linear(x[:, :1], w, b, cfg, split=True), linear(x[:, 1:], w, b, cfg)], dim=1)


def _ste(x, cfg):
    return _RoundSTE.apply(x) if cfg.bf16 else x


def layer_norm(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


class _GeluBF16(torch.autograd.Function):
    """GELU of the fp32 pre-activation; the CUDA path stores the derivative GELU'(u) in bf16 for the backward pass."""

    @staticmethod
    def forward(ctx, u):
        cdf = 0.5 * (1.0 + torch.erf(u * 0.7071067811865476))
        pdf = 0.3989422804014327 * torch.exp(-0.5 * u * u)
        ctx.save_for_backward(_r(cdf + u * pdf))
        return u * cdf

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        return g * d


def gelu(x, cfg):
    # exact erf GELU (nn.GELU default, video_transformer.py:35-49; DistilBERT activation "gelu")
    return _GeluBF16.apply(x) if cfg.bf16 else F.gelu(x)


def _softmax_attention(q, k, v, cfg, add_mask=None, drop=None):
    """softmax(q k^T (+mask)) v over the last two dims (video_transformer.py:28-32). q is already scaled.
    drop: multiplier (0 or 1 / (1 - p)) on the softmax weights = nn.functional.dropout with a given mask."""
    q, k, v = _ste(q, cfg), _ste(k, cfg), _ste(v, cfg)
    s = q @ k.transpose(-1, -2)
    if add_mask is not None:
        s = s + add_mask
    p = s.softmax(dim=-1)
    if drop is not None:
        p = p * drop
    return _ste(p, cfg) @ v


# ----------------------------------------------------------------------------------------------------------------
# video tower
# ----------------------------------------------------------------------------------------------------------------
def video_tokens(video, p, cfg, objects=None, prefix="video_model."):
    """Token assembly. Restates VideoPatchEmbed.forward (video_transformer.py:71-76) and forward_features :303-325:
    Conv2d(k16,s16) == per-patch linear map; tokens are frame-major: index 1 + f*n + row*14 + col; CLS gets
    pos_embed[0]; patch (f, i) gets pos_embed[1+i] + temporal_embed[f].
    Extension X1/X2: `objects` (B, F, O, 2054) -> object_embed Linear(2054, 768)
    (oa_video_transformer_region.py:250), appended after the 196 patches of each frame; an object token gets
    temporal_embed[f] only (box geometry is inside the feature), plus token_type 1 (patches type 0) when
    cfg.modality_token."""
    B, Fr, C, H, W = video.shape
    ps = cfg.patch
    gh, gw = H // ps, W // ps
    N = gh * gw
    w = p[prefix + "patch_embed.proj.weight"]           # (D, C, ps, ps)
    D = w.shape[0]
    x = video.reshape(B * Fr, C, gh, ps, gw, ps).permute(0, 2, 4, 1, 3, 5).reshape(B * Fr * N, C * ps * ps)
    x = linear(x, w.reshape(D, -1), p[prefix + "patch_embed.proj.bias"], cfg).reshape(B, Fr, N, D)
    pos = p[prefix + "pos_embed"]                        # (1, 1+N, D)
    tem = p[prefix + "temporal_embed"]                   # (1, F_max, D)
    assert Fr <= tem.shape[1]                            # video_transformer.py:73
    x = x + pos[:, 1:, :].unsqueeze(1) + tem[:, :Fr, :].unsqueeze(2)
    if cfg.modality_token and (prefix + "token_type_embeddings.weight") in p:
        x = x + p[prefix + "token_type_embeddings.weight"][0]
    if objects is not None:
        O = objects.shape[2]
        o = linear(objects.reshape(B * Fr * O, -1), p[prefix + "object_embed.weight"], p[prefix + "object_embed.bias"],
                   cfg).reshape(B, Fr, O, D)
        o = o + tem[:, :Fr, :].unsqueeze(2)
        if cfg.modality_token and (prefix + "token_type_embeddings.weight") in p:
            o = o + p[prefix + "token_type_embeddings.weight"][1]
        x = torch.cat([x, o], dim=2)
    n = x.shape[2]
    cls = (p[prefix + "cls_token"] + pos[:, 0:1, :]).expand(B, 1, D)
    return torch.cat([cls, x.reshape(B, Fr * n, D)], dim=1), n


def divided_attention_core(q, k, v, mode, Fr, n, cfg):
    """The attention arithmetic of VarAttention.forward (video_transformer.py:108-131) on already-projected,
    already-scaled q and k, v of shape (B, h, T, d): CLS query over all keys; patch queries over [CLS] + their group.
    Returns (B, T, h*d)."""
    B, h, T, d = q.shape
    cls_out = _softmax_attention(q[:, :, 0:1], k, v, cfg)                  # (B, h, 1, d)

    def grp(t):
        t = t[:, :, 1:].reshape(B, h, Fr, n, d)
        return t if mode == "space" else t.transpose(2, 3)                 # (B, h, G, L, d)

    q_, k_, v_ = grp(q), grp(k), grp(v)
    G = q_.shape[2]
    kc = k[:, :, 0:1].unsqueeze(2).expand(B, h, G, 1, d)
    vc = v[:, :, 0:1].unsqueeze(2).expand(B, h, G, 1, d)
    out = _softmax_attention(q_, torch.cat([kc, k_], dim=3), torch.cat([vc, v_], dim=3), cfg)
    if mode != "space":
        out = out.transpose(2, 3)
    out = torch.cat([cls_out, out.reshape(B, h, Fr * n, d)], dim=2)       # CLS first (:128)
    return out.permute(0, 2, 1, 3).reshape(B, T, h * d)                    # '(b h) n d -> b n (h d)' (:131)


def divided_attention(x, p, pre, mode, Fr, n, cfg):
    """VarAttention.forward (video_transformer.py:99-135). mode 'space': patch queries of frame f attend to
    [CLS] + the n tokens of frame f ('b (f n) d -> (b f) n d'); mode 'time': token (f, i) attends to [CLS] + tokens
    (f', i) for all f' ('b (f n) d -> (b n) f d'). The CLS query attends to every key (:110). q is scaled by
    head_dim^-0.5 before any product (:105)."""
    B, T, D = x.shape
    h = cfg.heads
    d = D // h
    qkv = linear_cls(x, p[pre + "qkv.weight"], p[pre + "qkv.bias"], cfg).reshape(B, T, 3, h, d)
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))        # (B, h, T, d)
    out = divided_attention_core(q * (d ** -0.5), k, v, mode, Fr, n, cfg)
    return linear_cls(_ste(out, cfg), p[pre + "proj.weight"], p[pre + "proj.bias"], cfg)   # attention output is stored bf16


def mlp(x, p, pre, cfg):
    """Mlp.forward (video_transformer.py:45-51), dropout p=0."""
    u = linear_cls(x, p[pre + "fc1.weight"], p[pre + "fc1.bias"], cfg)
    return linear_cls(gelu(u, cfg), p[pre + "fc2.weight"], p[pre + "fc2.bias"], cfg)


def space_time_block(x, p, pre, Fr, n, cfg):
    """SpaceTimeBlock.forward (video_transformer.py:161-176), 'frozen-in-time' residual wiring: the space residual
    skips from the block input x, not from the time residual."""
    e = cfg.ln_eps_video
    t_out = divided_attention(layer_norm(x, p[pre + "norm3.weight"], p[pre + "norm3.bias"], e), p, pre + "timeattn.",
                              "time", Fr, n, cfg)
    t_res = x + t_out
    s_out = divided_attention(layer_norm(t_res, p[pre + "norm1.weight"], p[pre + "norm1.bias"], e), p, pre + "attn.",
                              "space", Fr, n, cfg)
    s_res = x + s_out
    return s_res + mlp(layer_norm(s_res, p[pre + "norm2.weight"], p[pre + "norm2.bias"], e), p, pre + "mlp.", cfg)


def video_tower(video, p, cfg, objects=None, prefix="video_model.", depth=None, return_tokens=False, region_layer=None):
    """SpaceTimeTransformer.forward_features (video_transformer.py:303-351): tokens -> blocks -> final LN -> CLS;
    return_tokens: (x[:, 0], x[:, 1:]) as :351 returns. region_layer=K: the region variant
    (oa_video_transformer_region.py:364-376) returns (norm(x)[:, 0], region_norm(x after K blocks)[:, 1:]), K = 6."""
    x, n = video_tokens(video, p, cfg, objects, prefix)
    Fr = video.shape[1]
    if depth is None:
        depth = 1 + max(int(k[len(prefix) + 7:].split(".")[0]) for k in p if k.startswith(prefix + "blocks."))
    region = None
    for i in range(depth):
        x = space_time_block(x, p, "%sblocks.%d." % (prefix, i), Fr, n, cfg)
        if region_layer is not None and i + 1 == region_layer:
            region = layer_norm(x, p[prefix + "region_norm.weight"], p[prefix + "region_norm.bias"],
                                cfg.ln_eps_video)[:, 1:]
    x = layer_norm(x, p[prefix + "norm.weight"], p[prefix + "norm.bias"], cfg.ln_eps_video)
    if region_layer is not None:
        return x[:, 0], region
    return (x[:, 0], x[:, 1:]) if return_tokens else x[:, 0]


# ----------------------------------------------------------------------------------------------------------------
# text tower (HF DistilBERT, call site model/oa_model.py:110-115)
# ----------------------------------------------------------------------------------------------------------------
def distilbert(input_ids, attention_mask, p, cfg, prefix="text_model.", drop=None):
    """DistilBertModel forward: word + position embeddings, LN(1e-12); 6 post-LN blocks of
    [q/k/v/out linears, softmax(q k^T / sqrt(d) + key-padding mask) v, LN(x + attn), lin2(GELU(lin1)), LN(x + ffn)].
    Returns last_hidden_state (B, L, 768). Eval mode unless `drop` is given: a dict site -> multiplier tensor (0 or
    1 / (1 - p)) for DistilBERT's three dropout sites - 0: after the embedding LayerNorm (B, L, D); 1 + 3 i: softmax
    weights of layer i (B, h, L, L); 2 + 3 i: FFN output of layer i before the residual add (B, L, D) - i.e.
    nn.Dropout in training mode with the masks injected (the reference trains with text_model.train(), oa_model.py:28)."""
    B, L = input_ids.shape
    h = cfg.heads
    e = cfg.ln_eps_text
    x = p[prefix + "embeddings.word_embeddings.weight"][input_ids] + \
        p[prefix + "embeddings.position_embeddings.weight"][:L].unsqueeze(0)
    x = layer_norm(x, p[prefix + "embeddings.LayerNorm.weight"], p[prefix + "embeddings.LayerNorm.bias"], e)
    if drop is not None and 0 in drop:
        x = x * drop[0]
    D = x.shape[-1]
    d = D // h
    add_mask = None
    if attention_mask is not None:
        add_mask = torch.zeros(B, 1, 1, L, dtype=x.dtype, device=x.device)
        add_mask = add_mask.masked_fill(attention_mask.reshape(B, 1, 1, L) == 0, torch.finfo(x.dtype).min)
    for i in range(cfg.text_layers):
        pre = "%stransformer.layer.%d." % (prefix, i)

        def heads(t):
            return t.reshape(B, L, h, d).permute(0, 2, 1, 3)

        q = heads(linear(x, p[pre + "attention.q_lin.weight"], p[pre + "attention.q_lin.bias"], cfg, True)) * (d ** -0.5)
        k = heads(linear(x, p[pre + "attention.k_lin.weight"], p[pre + "attention.k_lin.bias"], cfg, True))
        v = heads(linear(x, p[pre + "attention.v_lin.weight"], p[pre + "attention.v_lin.bias"], cfg, True))
        dm = drop.get(1 + 3 * i) if drop is not None else None
        ctx = _ste(_softmax_attention(q, k, v, cfg, add_mask, dm).permute(0, 2, 1, 3).reshape(B, L, D), cfg)  # stored bf16
        sa = linear(ctx, p[pre + "attention.out_lin.weight"], p[pre + "attention.out_lin.bias"], cfg, True)
        x = layer_norm(sa + x, p[pre + "sa_layer_norm.weight"], p[pre + "sa_layer_norm.bias"], e)
        f = linear(gelu(linear(x, p[pre + "ffn.lin1.weight"], p[pre + "ffn.lin1.bias"], cfg, True), cfg),
                   p[pre + "ffn.lin2.weight"], p[pre + "ffn.lin2.bias"], cfg, True)
        if drop is not None and (2 + 3 * i) in drop:
            f = f * drop[2 + 3 * i]
        x = layer_norm(f + x, p[pre + "output_layer_norm.weight"], p[pre + "output_layer_norm.bias"], e)
    return x


# ----------------------------------------------------------------------------------------------------------------
# dual encoder, similarity, loss
# ----------------------------------------------------------------------------------------------------------------
def compute_text(text, p, cfg, drop=None):
    """FrozenInTime.compute_text (oa_model.py:106-123): last_hidden_state[:, 0] -> ReLU -> Linear(768, 256)."""
    hid = distilbert(text["input_ids"], text.get("attention_mask"), p, cfg, drop=drop)[:, 0]
    return linear(F.relu(hid.float()), p["txt_proj.1.weight"], p["txt_proj.1.bias"], cfg, True)


def compute_video(video, p, cfg, objects=None):
    """FrozenInTime.compute_video (oa_model.py:129-133): CLS feature -> Linear(768, 256)."""
    return linear(video_tower(video, p, cfg, objects), p["vid_proj.0.weight"], p["vid_proj.0.bias"], cfg, True)


def dual_encoder(data, p, cfg):
    """FrozenInTime.forward (oa_model.py:97-104) -> (text_embeddings, video_embeddings)."""
    return compute_text(data["text"], p, cfg), compute_video(data["video"], p, cfg, data.get("object"))


def region_mem_forward(data, p, cfg, region_layer=6):
    """FrozenInTime.forward of the region-sensitive variant (model/oa_model_region_mem.py:105-123):
    data['video'] (B, 2F, 3, H, W) viewed as 2B clips (even = anchor frames, odd = video, :111-117); the region tower
    returns (norm(x)[:, 0], region_norm(x after 6 blocks)[:, 1:]) (oa_video_transformer_region.py:364-376); vid_proj on
    both (:140-145); video_embeddings = (vid_proj(cls) + mean(vid_proj(regions), 1)) / 2 of the video clips (:119);
    region_sim = sigmoid(einsum('b k f, b n f -> b k n', txt_proj_2(text_region_embedding), anchor regions)) (:118,147-151).
    Returns (text_embeddings, video_embeddings, region_sim)."""
    text_e = compute_text(data["text"], p, cfg)
    v = data["video"]
    v = v.reshape(v.shape[0] * 2, -1, v.shape[2], v.shape[3], v.shape[4])
    cls, region = video_tower(v, p, cfg, region_layer=region_layer)
    cls_p = linear(cls, p["vid_proj.0.weight"], p["vid_proj.0.bias"], cfg, True)
    reg_p = linear(region, p["vid_proj.0.weight"], p["vid_proj.0.bias"], cfg)
    obj_region = reg_p[0::2]
    video_e, video_region = cls_p[1::2], reg_p[1::2]
    tre = linear(F.relu(data["text_region_embedding"].float()), p["txt_proj_2.1.weight"], p["txt_proj_2.1.bias"], cfg)
    video_e = (video_e + video_region.mean(dim=1)) / 2
    region_sim = torch.sigmoid(torch.einsum("bkf,bnf->bkn", tre, obj_region))
    return text_e, video_e, region_sim


def region_loss(region_sim, patch_mask, weight=0.1):
    """trainer/trainer_region_mem.py:161-167: 0.1 * BCELoss(reduction='sum')(region_sim rows, mask rows) / rows."""
    rs = region_sim.reshape(-1, region_sim.shape[-1])
    pm = patch_mask.reshape(-1, patch_mask.shape[-1]).to(rs.dtype)
    return weight * F.binary_cross_entropy(rs, pm, reduction="sum") / rs.shape[0]


def global_local_loss(text, pad_text, video, region_feat, tags_feat, temperature=0.05):
    """trainer/trainer_global_local.py:187-208: short-text and tag-padded-text InfoNCE against the video embeddings plus
    the fine-grained InfoNCE between mean-pooled region and tag features."""
    return norm_softmax_loss(sim_matrix(text, video), temperature) + \
        norm_softmax_loss(sim_matrix(pad_text, video), temperature) + \
        norm_softmax_loss(sim_matrix(region_feat.mean(dim=1), tags_feat.mean(dim=1)), temperature)


def sim_matrix(a, b, eps=1e-8):
    """model/model.py:164-172: rows L2-normalised with the norm clamped at eps, then a_n @ b_n^T."""
    an = a / a.norm(dim=1, keepdim=True).clamp_min(eps)
    bn = b / b.norm(dim=1, keepdim=True).clamp_min(eps)
    return an @ bn.t()


def norm_softmax_loss(x, temperature=0.05):
    """NormSoftmaxLoss.forward (model/loss.py:13-25): symmetric InfoNCE, mean over the diagonal of both
    log-softmaxes (rows: text->video, columns: video->text)."""
    i = F.log_softmax(x / temperature, dim=1)
    j = F.log_softmax(x.t() / temperature, dim=1)
    return -i.diag().mean() - j.diag().mean()


def gathered_loss(text_local, video_local, rank, world, all_text, all_video, temperature=0.05):
    """trainer_dist.py:158-163 with AllGather_multi (:29-45) semantics: the local rows are spliced into the gathered
    (detached) global matrices, so autograd returns exactly the local slice of the global gradient, unreduced."""
    B = text_local.shape[0]
    t = torch.cat([all_text[:rank * B].detach(), text_local, all_text[(rank + 1) * B:].detach()], 0)
    v = torch.cat([all_video[:rank * B].detach(), video_local, all_video[(rank + 1) * B:].detach()], 0)
    return norm_softmax_loss(sim_matrix(t, v), temperature)


# ----------------------------------------------------------------------------------------------------------------
# object -> patch attention (X4) and the integer bookkeeping around it
# ----------------------------------------------------------------------------------------------------------------
def patch_masks_from_bbox(bboxs, patch_rows=14):
    """base/base_dataset_global_local.py:348-356: boxes (x1, y1, x2, y2) in [0,1] -> binary masks over the row-major
    patch grid: rows int(y1*g) .. ceil(y2*g)-1, cols int(x1*g) .. ceil(x2*g)-1. Arithmetic in float64 like numpy."""
    b = np.asarray(bboxs, dtype=np.float64)[:, :4] * patch_rows
    masks = np.zeros((len(b), patch_rows, patch_rows), dtype=np.float64)
    for i in range(len(b)):
        masks[i, int(b[i, 1]):math.ceil(b[i, 3]), int(b[i, 0]):math.ceil(b[i, 2])] = 1
    return masks.reshape(len(b), patch_rows * patch_rows)


def patch_masks_same_class(bboxs, object_indexs, indexs, patch_rows=14):
    """base/base_dataset_region_mem.py:233-247 with the random.sample draw passed in as `indexs`: mask j is the union of
    the patch rectangles of every box whose class equals that of box indexs[j]. Returns (masks, selected classes)."""
    b = np.array(bboxs, dtype=np.float64, copy=True)
    b[:, :4] = b[:, :4] * patch_rows
    masks = np.zeros((len(indexs), patch_rows, patch_rows), dtype=np.float64)
    sel = []
    for j, i in enumerate(indexs):
        sel.append(object_indexs[i])
        for idx in range(len(b)):
            if object_indexs[idx] == object_indexs[i]:
                masks[j, int(b[idx, 1]):math.ceil(b[idx, 3]), int(b[idx, 0]):math.ceil(b[idx, 2])] = 1
    return masks.reshape(len(indexs), patch_rows ** 2), sel


def object_tags_masks(token_lens, indices):
    """base/base_dataset_global_local.py:395-405: running end offsets of each tag's tokens and the total length."""
    ends, end = [], 0
    for item in indices:
        end += int(token_lens[item])
        ends.append(float(end))
    return torch.tensor(ends, dtype=torch.float32), int(end)


def region_features_topk(x, bbox, conf, ids, image_w, image_h, top_k=10, v=1):
    """base/base_dataset.py:611-649 (read_object_from_disk once the .npz is loaded). numpy float32 in, torch fp32 out.
    np.pad(a, (0, res), 'edge') pads BOTH axes of a 2-D array: with fewer than top_k regions the feature block widens
    by res columns - kept, because that is what the reference returns."""
    order = np.argsort(conf)[::-1]
    boxes, feats = bbox[order], x[order]
    if v == 2:
        _, uniq = np.unique(ids, return_index=True)
        boxes, feats = boxes[uniq], feats[uniq]
    if boxes.shape[0] < top_k:
        res = top_k - boxes.shape[0]
        boxes, feats = np.pad(boxes, (0, res), 'edge'), np.pad(feats, (0, res), 'edge')
    boxes, feats = boxes[:top_k, :], feats[:top_k, :]
    bw, bh = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
    sw, sh, sx, sy = bw / image_w, bh / image_h, boxes[:, 0] / image_w, boxes[:, 1] / image_h
    spatial = np.stack([sx, sy, sx + sw, sy + sh, sw, sh], axis=1)
    return torch.cat([torch.from_numpy(feats), torch.from_numpy(spatial)], dim=1)


def object_patch_attention(q, k, v=None, mode="softmax", masks=None):
    """One op, three score->weight modes:
      'mask'    : weights = binary patch masks; out = masks @ v          (oa_model_global_local.py:178)
      'sigmoid' : weights = sigmoid(q k^T)                               (oa_model_region_mem.py:147-151)
      'softmax' : weights = softmax(q k^T * C^-0.5)                      (Visualization/.../visualize.py:155-168)
    q (B,O,C), k (B,L,C), v (B,L,Cv). Returns (weights (B,O,L), out (B,O,Cv) or None)."""
    if mode == "mask":
        w = masks.to(v.dtype if v is not None else masks.dtype)
    else:
        s = torch.einsum("bkf,bnf->bkn", q, k)
        w = torch.sigmoid(s) if mode == "sigmoid" else (s * q.shape[-1] ** -0.5).softmax(dim=-1)
    out = torch.einsum("bol,blc->boc", w, v) if v is not None else None
    return w, out


# ----------------------------------------------------------------------------------------------------------------
# retrieval metrics (model/metric.py:16-121 t2v, :123-212 v2t, :281-291 cols2metrics) for square sim matrices
# ----------------------------------------------------------------------------------------------------------------
def _cols2metrics(cols, num_queries):
    m = {"R1": 100 * float(np.sum(cols == 0)) / num_queries}
    for k in (5, 10, 50):
        m["R%d" % k] = 100 * float(np.sum(cols < k)) / num_queries
    m["MedR"] = float(np.median(cols) + 1)
    m["MeanR"] = float(np.mean(cols) + 1)
    stats = np.array([m["R1"], m["R5"], m["R10"]], dtype=np.float64)
    m["geometric_mean_R1-R5-R10"] = float(np.exp(np.mean(np.log(stats)))) if np.all(stats > 0) else 0.0
    return m


def t2v_metrics(sims):
    """Rank of the ground-truth video for each text query (rows = queries); ties broken optimistically
    (metric.py:62-69: the first matching column after a descending sort of -sims)."""
    sims = np.asarray(sims, dtype=np.float64)
    nq = sims.shape[0]
    dists = -sims
    sorted_d = np.sort(dists, axis=1)
    gt = np.diag(dists)[:, None]
    rows, cols = np.where((sorted_d - gt) == 0)
    _, idx = np.unique(rows, return_index=True)          # optimistic: first occurrence per row
    return _cols2metrics(cols[idx], nq)


def v2t_metrics(sims):
    """Rank of the ground-truth caption for each video query (columns of sims); ties averaged (metric.py:153,183)."""
    sims = np.asarray(sims, dtype=np.float64).T
    nq = sims.shape[0]
    dists = -sims
    ranks = np.zeros(nq)
    for i in range(nq):
        row = dists[i]
        srt = np.sort(row)
        loc = np.where((srt - row[i]) == 0)[0]
        ranks[i] = loc.mean()  # averaging tie rule
    return _cols2metrics(ranks, nq)


from oa_transformer_b200.synth import synth_objects, synth_text  # noqa: E402,F401  (shared input generators)
