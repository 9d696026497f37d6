"""Region-sensitive dual encoder: OATrans/model/oa_model_region_mem.py:18-151 on liboat.

forward(data) -> (text_embeddings (B, P), video_embeddings (B, P), region_sim (B, K, L)):
  * data['video'] (B, 2F, 3, H, W) is viewed as 2B clips, even clips = anchor frames, odd clips = the video (:111-117);
  * the video tower returns the CLS feature and region_norm'd patch features of layer 6; `vid_proj` is applied to both
    (:140-145);
  * video_embeddings = (vid_proj(cls) + mean_l vid_proj(patches)) / 2 of the video clips (:119);
  * region_sim = sigmoid(txt_proj_2(text_region_embedding) . vid_proj(patches of the anchor clips)^T) (:118,147-151),
    trained against bbox patch masks with 0.1 * BCE(sum) / rows (trainer/trainer_region_mem.py:161-167).
"""
import torch.nn as nn

from .. import functional as OF
from .oa_model import FrozenInTime as _FrozenInTime
from .oa_video_transformer_region import SpaceTimeTransformer


def init_weights(m):
    if type(m) == nn.Linear:
        nn.init.xavier_uniform_(m.weight)


class FrozenInTime(_FrozenInTime):
    VIDEO_TOWER = SpaceTimeTransformer

    def __init__(self, video_params, object_params, text_params, projection_dim=256, load_checkpoint=None,
                 projection='minimal', load_temporal_fix='zeros'):
        super().__init__(video_params, object_params, text_params, projection_dim=projection_dim,
                         load_checkpoint=None, projection=projection, load_temporal_fix=load_temporal_fix)
        self.txt_proj_2 = nn.Sequential(nn.ReLU(), nn.Linear(512, projection_dim))      # :70-72
        self.txt_proj.apply(init_weights)
        self.vid_proj.apply(init_weights)
        self.txt_proj_2.apply(init_weights)
        self.sigmod = nn.Sigmoid()
        if load_checkpoint not in ["", None]:
            self._load_checkpoint(load_checkpoint)

    def forward(self, data, aug=False, return_embeds=True):
        text_embeddings = self.compute_text(data['text'])
        video_data = data['video']
        video_data = video_data.view(video_data.size(0) * 2, -1, video_data.size(2), video_data.size(3),
                                     video_data.size(4))
        vision_embeddings, vision_region_feature = self.compute_video(video_data)
        object_region_feature = vision_region_feature[0::2].contiguous()
        video_embeddings, video_region_feature = vision_embeddings[1::2], vision_region_feature[1::2]
        text_region_embedding = OF.linear(data['text_region_embedding'].float(), self.txt_proj_2[1].weight,
                                          self.txt_proj_2[1].bias, relu=True)
        video_embeddings = OF.token_pool(video_embeddings, video_region_feature, 0.5, 0.5)
        region_sim = self.compute_region_sim(object_region_feature, text_region_embedding)
        return text_embeddings, video_embeddings, region_sim

    def compute_video(self, video_data, aug=False, object_data=None):
        """-> (vid_proj(cls) (2B, P), vid_proj(region features) (2B, L, P)) (:140-145)."""
        named = [("vid_proj." + n, p) for n, p in self.vid_proj.named_parameters()]
        cls_proj, tok = self.video_model.run(video_data.float(), extra_named=named,
                                             proj=("vid_proj.0.weight", "vid_proj.0.bias"))
        tok_proj = OF.linear(tok, self.vid_proj[0].weight, self.vid_proj[0].bias)       # every row, then drop the CLS row
        return cls_proj, tok_proj[:, 1:]

    def compute_region_sim(self, video_feats, text_feats):
        """sigmoid(einsum('b k f, b n f -> b k n', text_feats, video_feats)) (:147-151)."""
        weights, _ = OF.object_patch_attention(text_feats, video_feats.contiguous(), None, mode="sigmoid")
        return weights
