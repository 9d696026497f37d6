"""SpaceTimeTransformer: the parameter container of the space-time ViT, with the reference's constructor arguments,
attribute names and state_dict keys (OATrans/model/video_transformer.py:179-357), executed by
the liboat video engine instead of eager PyTorch.

Differences that are part of the design, not of the contract:
  * forward() returns (cls_feature, None) by default: only the CLS row is normalised by the final LayerNorm because that
    is the only row FrozenInTime consumes (oa_model.py:130-131). forward(x, return_tokens=True) returns
    (x[:, 0], x[:, 1:]) exactly like video_transformer.py:346-351 (final norm over every token row, differentiable).
  * object_tokens=True adds `object_embed` Linear(2054, embed_dim) (oa_video_transformer_region.py:250) and, with
    modality_token=True, `token_type_embeddings` (ibid. :257-261); forward then takes region features.
"""
from functools import partial

import torch
from torch import nn

from ..engine import VideoEngine
from ..functional import run_tower


class _Params(nn.Module):
    """A bag of parameters under a module name (keeps the reference's dotted state_dict keys)."""


def _linear_params(out_f, in_f, bias=True):
    m = nn.Linear(in_f, out_f, bias=bias)
    return m


class _VarAttention(nn.Module):
    def __init__(self, dim, num_heads, qkv_bias, initialize="random"):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        if initialize == "zeros":                      # video_transformer.py:89-95
            self.qkv.weight.data.fill_(0)
            self.qkv.bias.data.fill_(0)
            self.proj.weight.data.fill_(1)
            self.proj.bias.data.fill_(0)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _SpaceTimeBlock(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias, norm_layer, time_init):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = _VarAttention(dim, num_heads, qkv_bias)
        self.timeattn = _VarAttention(dim, num_heads, qkv_bias, initialize=time_init)
        self.norm2 = norm_layer(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))
        self.norm3 = norm_layer(dim)


class _PatchEmbed(nn.Module):
    def __init__(self, img_size, patch_size, in_chans, embed_dim, num_frames):
        super().__init__()
        img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        patch_size = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.img_size, self.patch_size = img_size, patch_size
        self.num_patches = (img_size[1] // patch_size[1]) * (img_size[0] // patch_size[0]) * num_frames
        self.num_frames, self.embed_dim = num_frames, embed_dim
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class SpaceTimeTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4., qkv_bias=True, qk_scale=None, representation_size=None,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0., hybrid_backbone=None, norm_layer=None,
                 num_frames=8, time_init='rand', attention_style='frozen-in-time', object_tokens=False,
                 modality_token=False):
        super().__init__()
        if hybrid_backbone is not None:
            raise NotImplementedError('hybrid backbone not implemented')
        if attention_style != 'frozen-in-time':
            raise NotImplementedError
        if drop_rate or attn_drop_rate or drop_path_rate:
            raise NotImplementedError("the CUDA path implements the shipped configuration: all drop rates 0")
        if qk_scale is not None or embed_dim != num_heads * 64:
            raise NotImplementedError("head_dim must be 64 with the default qk scale")
        if representation_size:
            raise NotImplementedError("representation layer is not on the hot path")
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.num_frames = num_frames
        self.num_heads = num_heads
        self.attention_style = attention_style
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        self.patch_embed = _PatchEmbed(img_size, patch_size, in_chans, embed_dim, num_frames)
        self.patches_per_frame = self.patch_embed.num_patches // num_frames
        self.object_tokens = object_tokens
        self.modality_token = modality_token
        if object_tokens:
            self.object_embed = nn.Linear(2054, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patches_per_frame + 1, embed_dim))
        self.temporal_embed = nn.Parameter(torch.zeros(1, num_frames, embed_dim))
        if modality_token:
            self.token_type_embeddings = nn.Embedding(2, embed_dim)
            nn.init.normal_(self.token_type_embeddings.weight, std=0.02)
        self.blocks = nn.ModuleList([
            _SpaceTimeBlock(embed_dim, num_heads, mlp_ratio, qkv_bias, norm_layer, time_init) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.eps = self.norm.eps
        self.pre_logits = nn.Identity()
        self.head = nn.Linear(self.num_features, num_classes) if num_classes > 0 else nn.Identity()
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.trunc_normal_(self.cls_token, std=.02)
        self._engine = None

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}

    def engine(self, device):
        if self._engine is None or self._engine.device != device:
            self._engine = VideoEngine(device, heads=self.num_heads, eps=self.eps, patch=self.patch_embed.patch_size[0])
        return self._engine

    def tower_params(self, prefix="video_model."):
        skip = ("head.", "pre_logits.", "fc.")
        return [(prefix + n, p) for n, p in self.named_parameters() if not n.startswith(skip)]

    def forward_features(self, x, objects=None, aug=False, return_tokens=False):
        named = self.tower_params()
        if return_tokens:
            cls_feature, tok = run_tower(self.engine(x.device), named, video=x, objects=objects, proj=None, tokens="final")
            return cls_feature, tok[:, 1:]
        out = run_tower(self.engine(x.device), named, video=x, objects=objects, proj=None)
        return out, None

    def forward(self, x, objects=None, aug=False, return_tokens=False):
        return self.forward_features(x, objects=objects, aug=aug, return_tokens=return_tokens)
