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
                 num_objects=0, text_len=32, pre_tokenized=True, args=None, region_mem=False, num_regions=5,
                 **_unused):
        video_params = video_params or {}
        self.dataset_name, self.split = dataset_name, split
        self.batch_size = batch_size
        self.n_samples = n_samples
        self.frames = video_params.get("num_frames", 4)
        self.res = video_params.get("input_res", 224)
        self.num_objects, self.text_len, self.pre_tokenized = num_objects, text_len, pre_tokenized
        self.region_mem, self.num_regions = region_mem, num_regions
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
            if self.region_mem:
                # the region-sensitive loaders (base/base_dataset_region_mem.py) return the anchor frame(s) first and
                # the clip after them, CLIP text-region embeddings (B, 5, 512) and bbox patch masks (B, 1, 5, 196)
                data["video"] = torch.randn(B, 2 * self.frames, 3, self.res, self.res, generator=g)
                data["text_region_embedding"] = torch.randn(B, self.num_regions, 512, generator=g)
                grid = (self.res // 16) ** 2
                data["patch_masks"] = (torch.rand(B, 1, self.num_regions, self.frames * grid, generator=g) > 0.7).double()
            if self.pre_tokenized:
                data["text"] = synth_text(B, self.text_len, g)
            else:
                data["text"] = ["a synthetic caption number %d" % (i * B + j) for j in range(B)]
            yield data


MultiDistTextObjectVideoDataLoader = SyntheticTextObjectVideoDataLoader
TextObjectVideoDataLoader = SyntheticTextObjectVideoDataLoader


class DevicePrefetcher:
    """Host -> device staging one batch ahead on a copy stream (what `pin_memory=True` + `.to(device,
    non_blocking=True)` in the reference trainer, trainer/trainer_dist.py:150-156, aims at): while step i computes,
    the pinned tensors of batch i+1 are already crossing PCIe into the other of two device buffer sets.

        pf = DevicePrefetcher(iterable_of_host_batches, device)
        for data in pf: ...            # data tensors live on `device`; valid until the next-but-one batch is requested
    """

    def __init__(self, batches, device, pool=None):
        """pool: optional dict kept by the caller; the two device buffer sets live in it, so prefetchers created one
        after the other (one per epoch) re-use the same device memory instead of allocating 2 x batch bytes again."""
        self.it = iter(batches)
        self.device = device
        self.pool = pool
        self.copy_stream = torch.cuda.Stream(device=device)
        self.slots = [None, None]          # device buffer sets
        self.ready = [None, None]          # copy-finished events
        self.free = [None, None]           # compute-finished events (buffer may be overwritten)
        self.k = 0
        self._stage(0)

    @staticmethod
    def _tensors(d, prefix=()):
        for key, v in d.items():
            if isinstance(v, dict):
                yield from DevicePrefetcher._tensors(v, prefix + (key,))
            elif torch.is_tensor(v):
                yield prefix + (key,), v

    def _stage(self, slot):
        try:
            host = next(self.it)
        except StopIteration:
            self.slots[slot] = None
            return
        if self.free[slot] is not None:
            self.copy_stream.wait_event(self.free[slot])
        prev = self.slots[slot]["dev"] if isinstance(self.slots[slot], dict) and "dev" in self.slots[slot] else {}
        if not prev and self.pool is not None:
            prev = self.pool.get(slot, {})
        dev = {}
        with torch.cuda.stream(self.copy_stream):
            for path, t in self._tensors(host):
                buf = prev.get(path)
                if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
                    buf = torch.empty(t.shape, dtype=t.dtype, device=self.device)
                buf.copy_(t, non_blocking=True)
                dev[path] = buf
        ev = torch.cuda.Event()
        ev.record(self.copy_stream)
        self.ready[slot] = ev
        self.slots[slot] = {"host": host, "dev": dev}
        if self.pool is not None:
            self.pool[slot] = dev

    def __iter__(self):
        return self

    def __next__(self):
        slot = self.k & 1
        cur = self.slots[slot]
        if cur is None:
            raise StopIteration
        main = torch.cuda.current_stream(self.device)
        main.wait_event(self.ready[slot])
        # the other buffer set was consumed by the previous step: mark it free once the work queued so far is done
        other = slot ^ 1
        ev = torch.cuda.Event()
        ev.record(main)
        self.free[other] = ev
        self._stage(other)
        self.k += 1

        def build(d, prefix=()):
            out = {}
            for key, v in d.items():
                if isinstance(v, dict):
                    out[key] = build(v, prefix + (key,))
                elif torch.is_tensor(v):
                    out[key] = cur["dev"][prefix + (key,)]
                else:
                    out[key] = v
            return out
        return build(cur["host"])
