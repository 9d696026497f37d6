"""Synthetic stand-in for OATrans/data_loader (datasets and video decoding are outside the hot-path scope): a loader
class with the constructor keywords of MultiDistTextObjectVideoDataLoader that yields seeded batches in the format
the trainer consumes - data['video'] fp32 (B,F,3,H,W), data['text'] list[str] or a token dict, data['object']
fp32 (B,F,O,2054), data['meta']."""
import torch

from ..synth import synth_objects, synth_text


class _Sampler:
    def set_epoch(self, epoch):
        self.epoch = epoch


class SyntheticTextObjectVideoDataLoader:
    def __init__(self, dataset_name="synthetic", text_params=None, video_params=None, object_params=None,
                 data_dir="", object_dir="", batch_size=16, split="train", num_workers=0, shuffle=True, n_samples=64,
                 num_objects=0, text_len=32, pre_tokenized=True, args=None, **_unused):
        video_params = video_params or {}
        self.dataset_name, self.split = dataset_name, split
        self.batch_size = batch_size
        self.n_samples = n_samples
        self.frames = video_params.get("num_frames", 4)
        self.res = video_params.get("input_res", 224)
        self.num_objects, self.text_len, self.pre_tokenized = num_objects, text_len, pre_tokenized
        self.rank = getattr(args, "rank", 0) if args is not None else 0
        self.train_sampler = _Sampler()

    def __len__(self):
        return max(1, self.n_samples // self.batch_size)

    def __iter__(self):
        for i in range(len(self)):
            g = torch.Generator().manual_seed(1234 + 1000 * self.rank + i + (0 if self.split == "train" else 7777))
            B = self.batch_size
            data = {"video": torch.randn(B, self.frames, 3, self.res, self.res, generator=g),
                    "meta": {"dataset": [self.dataset_name] * B, "paths": ["synthetic/%d" % (i * B + j) for j in range(B)]}}
            if self.num_objects:
                data["object"] = synth_objects(B, self.frames, self.num_objects, g)
            if self.pre_tokenized:
                data["text"] = synth_text(B, self.text_len, g)
            else:
                data["text"] = ["a synthetic caption number %d" % (i * B + j) for j in range(B)]
            yield data


MultiDistTextObjectVideoDataLoader = SyntheticTextObjectVideoDataLoader
TextObjectVideoDataLoader = SyntheticTextObjectVideoDataLoader
