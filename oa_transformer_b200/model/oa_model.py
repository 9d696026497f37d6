"""FrozenInTime: the dual encoder of OATrans/model/oa_model.py:10-133 with the same constructor, attributes,
methods, error behaviour and state_dict keys, executed by liboat.

  text : HF DistilBERT parameters (text_model.*, loaded with AutoModel.from_pretrained exactly like oa_model.py:27)
         -> TextEngine -> last_hidden_state[:, 0] -> ReLU -> Linear(768, projection_dim)      (oa_model.py:106-123)
  video: SpaceTimeTransformer parameters (video_model.*) -> VideoEngine -> CLS -> Linear(768, projection_dim) (:129-133)

Extensions selected by config keys that old configs do not carry (so they still load unchanged):
  video_params['model'] == 'SpaceTimeObjectTransformer' or object_params['input_objects'] truthy: per-frame
  object-region tokens are appended to each frame's patch tokens (data['object'], fp32 (B, F, O, 2054) or (B, O, 2054)
  for single-frame loaders) - SURVEY.md section 8a rows X1-X3.
  text_params['random_init'] (+ 'config': DistilBertConfig overrides) / video_params['vit_checkpoint'] /
  video_params['depth' | 'embed_dim' | 'num_heads' | 'patch_size' | 'img_size']: build without the pretrained files and
  at other sizes (benchmarks, fixtures).
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..base import BaseModel
from ..engine import TextEngine
from ..functional import run_tower
from ..utils import state_dict_data_parallel_fix
from .model import sim_matrix as _unused  # noqa: F401  (kept importable from here like the reference)
from .video_transformer import SpaceTimeTransformer


MAX_TEXT_TOKENS = 256      # oat_attn_fwd mode 2 keeps one caption's keys in shared memory


class FrozenInTime(BaseModel):
    VIDEO_TOWER = SpaceTimeTransformer       # variants swap the tower class (model/oa_model_region_mem.py)

    def __init__(self,
                 video_params,
                 object_params,
                 text_params,
                 projection_dim=256,
                 load_checkpoint=None,
                 projection='minimal',
                 load_temporal_fix='zeros'):
        super().__init__()
        self.video_params = video_params
        self.text_params = text_params
        self.object_params = object_params
        self.load_temporal_fix = load_temporal_fix
        if not text_params['pretrained']:
            raise NotImplementedError("Huggingface text models require pretrained init.")

        from transformers import AutoModel
        if text_params.get('random_init', False):
            from transformers import DistilBertConfig, DistilBertModel
            self.text_model = DistilBertModel(DistilBertConfig(**text_params.get('config', {})))
        else:
            self.text_model = AutoModel.from_pretrained(text_params['model'])
        self.text_model.train()
        if not text_params['model'].split('/')[-1].startswith('distilbert'):
            raise NotImplementedError("the CUDA text tower implements DistilBERT (the shipped configs' text model)")

        self.use_objects = bool(object_params.get('input_objects', False)) or \
            video_params['model'] == "SpaceTimeObjectTransformer"
        if video_params['model'] in ["SpaceTimeTransformer", "SpaceTimeObjectTransformer"]:
            num_frames = video_params.get('num_frames', 4)
            time_init = video_params.get('time_init', 'zeros')
            attention_style = video_params.get('attention_style', 'frozen-in-time')
            arch_config = video_params.get('arch_config', 'base_patch16_224')
            modality_token = video_params.get('modality_token', False)
            if arch_config == 'base_patch16_224':
                vit_path = video_params.get('vit_checkpoint', "pretrained/jx_vit_base_p16_224-80ecf9dd.pth")
                vit_model = torch.load(vit_path, map_location="cpu", weights_only=False) \
                    if vit_path and os.path.exists(vit_path) else None
                if vit_model is None and not video_params.get('allow_missing_vit', False):
                    raise FileNotFoundError(vit_path)
                model = self.VIDEO_TOWER(num_frames=num_frames, time_init=time_init,
                                         attention_style=attention_style, object_tokens=self.use_objects,
                                         modality_token=modality_token,
                                         img_size=video_params.get('img_size', 224),
                                         **{k: video_params[k] for k in ('depth', 'embed_dim', 'num_heads', 'patch_size')
                                            if k in video_params})
            else:
                raise NotImplementedError
            model.head = nn.Identity()
            model.pre_logits = nn.Identity()
            ftr_dim = model.embed_dim
            if load_checkpoint in ["", None] and vit_model is not None:
                model.load_state_dict(vit_model, strict=False)
            self.video_model = model
            self.video_model.fc = nn.Identity()
        elif video_params['model'] == "":
            print("no vision model available!")
        else:
            raise NotImplementedError(f"{video_params['model']} not implemented")

        if projection == 'minimal':
            txt_proj = nn.Sequential(nn.ReLU(), nn.Linear(self.text_model.config.hidden_size, projection_dim))
            if video_params['model'] != "":
                vid_proj = nn.Sequential(nn.Linear(ftr_dim, projection_dim))
        elif projection != '':
            raise NotImplementedError("only the 'minimal' projection is on the CUDA path")
        else:
            raise NotImplementedError
        self.txt_proj = txt_proj
        if video_params['model'] != "":
            self.vid_proj = vid_proj

        self._text_engine = None
        if load_checkpoint not in ["", None]:
            self._load_checkpoint(load_checkpoint)

    def _load_checkpoint(self, load_checkpoint):
        # weights_only=False: reference checkpoints pickle their config next to the tensors (base_trainer.py:163-175)
        checkpoint = torch.load(load_checkpoint, map_location="cpu", weights_only=False)
        state_dict = checkpoint['state_dict']
        new_state_dict = state_dict_data_parallel_fix(state_dict, self.state_dict())
        new_state_dict = self._inflate_positional_embeds(new_state_dict)
        self.load_state_dict(new_state_dict, strict=False)

    def set_device(self, device):
        self.device = device

    def forward(self, data, aug=False, return_embeds=True):
        # The towers are independent (oa_model.py:97-104 runs text first). The video tower is enqueued first here: its
        # long GEMMs let the host run ahead, so the ~200 small launches of the text tower are already queued when the
        # GPU reaches them (matters right after a host sync, e.g. the per-step loss read-back).
        video_embeddings = self.compute_video(data['video'], aug=aug, object_data=data.get('object'))
        text_embeddings = self.compute_text(data['text'])
        if return_embeds:
            return text_embeddings, video_embeddings
        from .model import sim_matrix
        return sim_matrix(text_embeddings, video_embeddings)

    # ------------------------------------------------------------------ text
    def _text_named(self):
        named = [("text_model." + n, p) for n, p in self.text_model.named_parameters()]
        named += [("txt_proj." + n, p) for n, p in self.txt_proj.named_parameters()]
        return named

    def compute_text(self, text_data, pad=False):
        if not self.text_params['model'].split('/')[-1].startswith('distilbert'):
            raise NotImplementedError
        ids = text_data['input_ids']
        if ids.shape[1] > MAX_TEXT_TOKENS:
            raise ValueError("the CUDA text-attention kernel handles up to %d tokens per caption, got %d: tokenize with "
                             "truncation=True, max_length=%d" % (MAX_TEXT_TOKENS, ids.shape[1], MAX_TEXT_TOKENS))
        if self._text_engine is None or self._text_engine.device != ids.device:
            self._text_engine = TextEngine(ids.device, heads=self.text_model.config.n_heads)
        # `self.text_model.train()` (oa_model.py:28) leaves DistilBERT's dropout on whenever the module is in training mode:
        # embedding, attention-weight and FFN-output dropout at the config's rates, fresh Philox masks every call
        # (seeded from torch's CPU generator, so torch.manual_seed reproduces a run); model.eval() turns it off
        dropout = None
        if self.training:
            cfg = self.text_model.config
            pd, pa = float(getattr(cfg, "dropout", 0.0)), float(getattr(cfg, "attention_dropout", 0.0))
            if pd > 0.0 or pa > 0.0:
                dropout = {"p": pd, "p_attn": pa, "seed": int(torch.randint(0, 2 ** 62, (1,)).item())}
        return run_tower(self._text_engine, self._text_named(), input_ids=ids,
                         attention_mask=text_data.get('attention_mask'), dropout=dropout)

    # ------------------------------------------------------------------ video
    def _video_named(self):
        return self.video_model.tower_params() + [("vid_proj." + n, p) for n, p in self.vid_proj.named_parameters()]

    def compute_video(self, video_data, aug=False, object_data=None):
        objects = None
        if self.use_objects:
            if object_data is None:
                raise ValueError("this model was built with object tokens: data['object'] is required")
            objects = object_data
            if objects.dim() == 3:        # (B, O, 2054) single-frame loaders (base_dataset.py:356)
                objects = objects.unsqueeze(1).expand(-1, video_data.shape[1], -1, -1)
            objects = objects.float()
        return run_tower(self.video_model.engine(video_data.device), self._video_named(), video=video_data.float(),
                         objects=objects)

    # ------------------------------------------------------------------ checkpoints
    def _inflate_positional_embeds(self, new_state_dict):
        """oa_model.py:148-189: adapt temporal_embed when the checkpoint was trained with another num_frames."""
        curr_keys = list(self.state_dict().keys())
        if 'video_model.temporal_embed' in new_state_dict and 'video_model.temporal_embed' in curr_keys:
            load_temporal_embed = new_state_dict['video_model.temporal_embed']
            load_num_frames = load_temporal_embed.shape[1]
            curr_num_frames = self.video_params['num_frames']
            embed_dim = load_temporal_embed.shape[2]
            if load_num_frames != curr_num_frames:
                if load_num_frames > curr_num_frames:
                    new_temporal_embed = load_temporal_embed[:, :curr_num_frames, :]
                else:
                    if self.load_temporal_fix == 'zeros':
                        new_temporal_embed = torch.zeros([load_temporal_embed.shape[0], curr_num_frames, embed_dim])
                        new_temporal_embed[:, :load_num_frames] = load_temporal_embed
                    elif self.load_temporal_fix in ['interp', 'bilinear']:
                        mode = 'bilinear' if self.load_temporal_fix == 'bilinear' else 'nearest'
                        new_temporal_embed = F.interpolate(load_temporal_embed.unsqueeze(0),
                                                           (curr_num_frames, embed_dim), mode=mode).squeeze(0)
                    else:
                        raise NotImplementedError
                new_state_dict['video_model.temporal_embed'] = new_temporal_embed
        if 'video_model.pos_embed' in new_state_dict and 'video_model.pos_embed' in curr_keys:
            if new_state_dict['video_model.pos_embed'].shape[1] != self.state_dict()['video_model.pos_embed'].shape[1]:
                raise NotImplementedError(
                    'Loading models with different spatial resolution / patch number not yet implemented, sorry.')
        return new_state_dict
