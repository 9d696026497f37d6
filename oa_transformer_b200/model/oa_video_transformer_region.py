"""Region variant of the space-time ViT (OATrans/model/oa_video_transformer_region.py:195-376): the same tower plus
`region_norm`, returning `(norm(x)[:, 0], region_norm(x after K = 6 blocks)[:, 1:])` (:364-376). `object_embed`
(Linear(2054, embed_dim), :250) and the optional `token_type_embeddings` (:257-261) are declared, as in the reference,
and - as in the reference's forward - not used by this class."""
from functools import partial

from torch import nn

from ..functional import run_tower
from .video_transformer import SpaceTimeTransformer as _SpaceTimeTransformer


class SpaceTimeTransformer(_SpaceTimeTransformer):
    REGION_LAYER = 6                 # "v2: extract region feature from k-th layer", K = 6 (:364)

    def __init__(self, *args, modality_token=False, two_outputs=False, norm_layer=None, **kwargs):
        super().__init__(*args, norm_layer=norm_layer, **kwargs)
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        self.object_embed = nn.Linear(2054, self.embed_dim)
        self.modality_token = modality_token
        if modality_token:
            self.token_type_embeddings = nn.Embedding(2, self.embed_dim)
            nn.init.normal_(self.token_type_embeddings.weight, std=0.02)
        self.two_outputs = two_outputs
        self.region_norm = norm_layer(self.embed_dim)

    def tower_params(self, prefix="video_model."):
        skip = ("head.", "pre_logits.", "fc.", "object_embed.", "token_type_embeddings.")    # unused by forward
        return [(prefix + n, p) for n, p in self.named_parameters() if not n.startswith(skip)]

    def run(self, x, extra_named=(), proj=None):
        """(cls feature or its projection, token features (B, T, D) after region_norm at layer K)."""
        layer = min(self.REGION_LAYER, len(self.blocks))
        return run_tower(self.engine(x.device), self.tower_params() + list(extra_named), video=x, proj=proj,
                         tokens="region", region_layer=layer)

    def forward_features(self, x):
        cls_token, tok = self.run(x)
        return cls_token, tok[:, 1:]

    def forward(self, x):
        return self.forward_features(x)
