"""Deterministic synthetic weights and inputs (no network, no datasets): shared by bench.py, smoke(), the tests and -
re-exported through oracle/weights.py - the golden-fixture generator, so that every side sees identical tensors.

Weights: tensor i (in the given order) is drawn from torch.Generator(seed * 100003 + i): LayerNorm gains (1-D
'.weight' whose name contains 'norm') are 1 + 0.1 N(0,1); every other floating tensor is scale * N(0,1).
Inputs follow SURVEY.md section 8d: video ~ N(0,1) (post-Normalize statistics, base/base_dataset.py:176); region
features [2048 ROI | x1,y1,x2,y2,w,h] with boxes in [0,1] (base/base_dataset.py:593-650); token ids
[CLS] ... [SEP] with an all-ones or ragged attention mask.
"""
import torch


def fill_seeded(spec, seed, scale=0.02):
    """spec: dict name -> tensor (state_dict) or name -> shape tuple. Returns dict name -> fp32 tensor."""
    out = {}
    for i, (name, v) in enumerate(spec.items()):
        if torch.is_tensor(v):
            if not v.is_floating_point():
                out[name] = v.clone()
                continue
            shape = tuple(v.shape)
        else:
            shape = tuple(v)
        g = torch.Generator().manual_seed(seed * 100003 + i)
        t = torch.randn(shape, generator=g)
        if len(shape) == 1 and name.endswith(".weight") and "norm" in name.lower():
            out[name] = 1.0 + 0.1 * t
        else:
            out[name] = scale * t
    return out


def video_tower_spec(depth=12, dim=768, frames=8, grid=14, patch=16, objects=False, prefix="video_model."):
    """Name -> shape of SpaceTimeTransformer's state_dict, optionally with the object_embed
    extension (oa_video_transformer_region.py:250)."""
    s = {}
    s[prefix + "cls_token"] = (1, 1, dim)
    s[prefix + "pos_embed"] = (1, grid * grid + 1, dim)
    s[prefix + "temporal_embed"] = (1, frames, dim)
    s[prefix + "patch_embed.proj.weight"] = (dim, 3, patch, patch)
    s[prefix + "patch_embed.proj.bias"] = (dim,)
    if objects:
        s[prefix + "object_embed.weight"] = (dim, 2054)
        s[prefix + "object_embed.bias"] = (dim,)
    for i in range(depth):
        b = "%sblocks.%d." % (prefix, i)
        s[b + "norm1.weight"] = (dim,)
        s[b + "norm1.bias"] = (dim,)
        for a in ("attn.", "timeattn."):
            s[b + a + "qkv.weight"] = (3 * dim, dim)
            s[b + a + "qkv.bias"] = (3 * dim,)
            s[b + a + "proj.weight"] = (dim, dim)
            s[b + a + "proj.bias"] = (dim,)
        s[b + "norm2.weight"] = (dim,)
        s[b + "norm2.bias"] = (dim,)
        s[b + "mlp.fc1.weight"] = (4 * dim, dim)
        s[b + "mlp.fc1.bias"] = (4 * dim,)
        s[b + "mlp.fc2.weight"] = (dim, 4 * dim)
        s[b + "mlp.fc2.bias"] = (dim,)
        s[b + "norm3.weight"] = (dim,)
        s[b + "norm3.bias"] = (dim,)
    s[prefix + "norm.weight"] = (dim,)
    s[prefix + "norm.bias"] = (dim,)
    return s


def text_tower_spec(layers=6, dim=768, hidden=3072, vocab=30522, max_pos=512, prefix="text_model."):
    """Name -> shape of HF DistilBertModel's state_dict."""
    s = {}
    s[prefix + "embeddings.word_embeddings.weight"] = (vocab, dim)
    s[prefix + "embeddings.position_embeddings.weight"] = (max_pos, dim)
    s[prefix + "embeddings.LayerNorm.weight"] = (dim,)
    s[prefix + "embeddings.LayerNorm.bias"] = (dim,)
    for i in range(layers):
        b = "%stransformer.layer.%d." % (prefix, i)
        for lin in ("q_lin", "k_lin", "v_lin", "out_lin"):
            s[b + "attention.%s.weight" % lin] = (dim, dim)
            s[b + "attention.%s.bias" % lin] = (dim,)
        s[b + "sa_layer_norm.weight"] = (dim,)
        s[b + "sa_layer_norm.bias"] = (dim,)
        s[b + "ffn.lin1.weight"] = (hidden, dim)
        s[b + "ffn.lin1.bias"] = (hidden,)
        s[b + "ffn.lin2.weight"] = (dim, hidden)
        s[b + "ffn.lin2.bias"] = (dim,)
        s[b + "output_layer_norm.weight"] = (dim,)
        s[b + "output_layer_norm.bias"] = (dim,)
    return s


def dual_encoder_spec(frames=8, objects=False, proj=256, **kw):
    s = {}
    s.update(text_tower_spec())
    s.update(video_tower_spec(frames=frames, objects=objects, **kw))
    s["txt_proj.1.weight"] = (proj, 768)
    s["txt_proj.1.bias"] = (proj,)
    s["vid_proj.0.weight"] = (proj, 768)
    s["vid_proj.0.bias"] = (proj,)
    return s


def synth_objects(B, Fr, O, gen):
    """Region features in the format of base/base_dataset.py:593-650: [2048 ROI feats | x1,y1,x2,y2,w,h] with boxes
    scaled to [0,1]; rows are already in confidence order."""
    feat = torch.randn(B, Fr, O, 2048, generator=gen).abs()
    x1 = torch.rand(B, Fr, O, 1, generator=gen) * 0.7
    y1 = torch.rand(B, Fr, O, 1, generator=gen) * 0.7
    w = 0.1 + torch.rand(B, Fr, O, 1, generator=gen) * 0.2
    h = 0.1 + torch.rand(B, Fr, O, 1, generator=gen) * 0.2
    return torch.cat([feat, x1, y1, x1 + w, y1 + h, w, h], dim=-1)


def synth_text(B, L, gen, vocab=30522, ragged=False):
    ids = torch.randint(1000, vocab, (B, L), generator=gen)
    ids[:, 0] = 101
    mask = torch.ones(B, L, dtype=torch.long)
    if ragged:
        lens = torch.randint(4, L + 1, (B,), generator=gen)
        lens[0] = L
        for b in range(B):
            ids[b, lens[b] - 1] = 102
            ids[b, lens[b]:] = 0
            mask[b, lens[b]:] = 0
    else:
        ids[:, -1] = 102
    return {"input_ids": ids, "attention_mask": mask}
