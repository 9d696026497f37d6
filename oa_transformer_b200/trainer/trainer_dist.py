"""Multi_Trainer_dist / AllGather_multi with the call contract of OATrans/trainer/trainer_dist.py.

The six hot lines of `_train_epoch` (:158-163) are unchanged in meaning:
    text_embeds, video_embeds = self.model(data, aug=True)
    video_embeds = self.allgather(video_embeds, self.n_gpu, self.args)
    text_embeds  = self.allgather(text_embeds,  self.n_gpu, self.args)
    output = sim_matrix(text_embeds, video_embeds);  loss = self.loss(output);  loss.backward()
but every one of those callables is a liboat launch sequence. Validation gathers embeddings, builds the full
similarity matrix and evaluates the retrieval metrics (t2v / v2t) like `_valid_epoch` (:201-281)."""
import time

import numpy as np
import torch
import torch.distributed as dist

from ..base import Multi_BaseTrainer_dist
from ..functional import AllGatherPairSlice, AllGatherSlice
from ..model.model import sim_matrix
from ..utils import inf_loop


class AllGather_multi(torch.autograd.Function):
    """apply(tensor, n_gpu, args): all-gather over args.world_size ranks, concatenated in rank order; the backward is
    the slice of rows owned by args.rank, with no reduction (trainer_dist.py:29-45)."""

    @staticmethod
    def forward(ctx, tensor, n_gpu, args):
        ctx.rank, ctx.bs = args.rank, tensor.shape[0]
        return AllGatherSlice.forward(ctx, tensor, args.rank, args.world_size)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output[ctx.bs * ctx.rank: ctx.bs * (ctx.rank + 1)], None, None


def allgather_pair(video_embeds, text_embeds, n_gpu, args):
    """Both AllGather_multi calls of trainer_dist.py:159-160 as one packed collective -> (video_all, text_all)."""
    return AllGatherPairSlice.apply(video_embeds, text_embeds, args.rank, args.world_size)


class Multi_Trainer_dist(Multi_BaseTrainer_dist):
    def __init__(self, args, model, loss, metrics, optimizer, config, data_loader, valid_data_loader=None,
                 lr_scheduler=None, len_epoch=None, writer=None, visualizer=None, tokenizer=None,
                 max_samples_per_epoch=50000):
        super().__init__(args, model, loss, metrics, optimizer, config, writer)
        self.config, self.args = config, args
        self.data_loader = data_loader
        if len_epoch is None:
            self.len_epoch = min(len(x) for x in data_loader)
        else:
            self.data_loader = [inf_loop(x) for x in data_loader]
            self.len_epoch = len_epoch
        self.valid_data_loader = valid_data_loader
        self.do_validation = self.valid_data_loader is not None
        self.lr_scheduler = lr_scheduler
        self.visualizer = visualizer
        self.batch_size = self.data_loader[0].batch_size
        self.log_step = max(1, int(np.sqrt(self.batch_size)))
        self.total_batch_sum = sum(x.batch_size for x in self.data_loader)
        self.tokenizer = tokenizer
        self.max_samples_per_epoch = max_samples_per_epoch
        self.n_gpu = self.args.world_size
        self.allgather = AllGather_multi.apply

    def _to_device(self, data):
        if self.tokenizer is not None and not isinstance(data['text'], dict):
            # trainer_dist.py:152 (padding=True, truncation=True); the CUDA text attention takes at most 256 tokens
            data['text'] = self.tokenizer(data['text'], return_tensors='pt', padding=True, truncation=True,
                                          max_length=min(getattr(self.tokenizer, "model_max_length", 256), 256))
        data['text'] = {k: v.to(self.device) for k, v in data['text'].items()}
        data['video'] = data['video'].to(self.device)
        if 'object' in data:
            data['object'] = data['object'].to(self.device)
        return data

    def _adjust_learning_rate(self, optimizer, epoch, args):
        lr = getattr(args, "learning_rate1", None)
        if lr is None:
            return
        for milestone in getattr(args, "schedule", []):
            lr *= 0.1 if epoch >= milestone else 1.
        for group in optimizer.param_groups:
            group['lr'] = lr

    def _train_epoch(self, epoch):
        self.model.train()
        total_loss = [0] * len(self.data_loader)
        for loader in self.data_loader:
            loader.train_sampler.set_epoch(epoch)
        begin = time.time()
        for batch_idx, data_li in enumerate(zip(*self.data_loader)):
            if (batch_idx + 1) * self.total_batch_sum > self.max_samples_per_epoch:
                break
            for dl_idx, data in enumerate(data_li):
                data = self._to_device(data)
                self.optimizer.zero_grad()
                with torch.set_grad_enabled(True):
                    text_embeds, video_embeds = self.model(data, aug=True)
                    # trainer_dist.py:159-160: two AllGather_multi calls; here one packed NCCL all-gather
                    video_embeds, text_embeds = allgather_pair(video_embeds, text_embeds, self.n_gpu, self.args)
                    output = sim_matrix(text_embeds, video_embeds)
                    loss = self.loss(output)
                loss.backward()
                self.optimizer.step()
                value = loss.detach().item()
                total_loss[dl_idx] += value
                if self.writer is not None and self.args.rank == 0:
                    self.writer.log_scalar(f'loss_train_{dl_idx}', value)
                if batch_idx % self.log_step == 0 and self.args.local_rank == 0:
                    self.logger.debug('Train Epoch: {} dl{} [{}/{}] Loss: {:.6f} ({:.2f}s)'.format(
                        epoch, dl_idx, batch_idx, self.len_epoch, value, time.time() - begin))
                    begin = time.time()
                self.optimizer.zero_grad()
            if batch_idx == self.len_epoch:
                break
        log = {f'loss_{i}': total_loss[i] / self.len_epoch for i in range(len(self.data_loader))}
        if self.do_validation:
            val_log = self._valid_epoch(epoch)
            if self.args.rank == 0:
                log.update(val_log)
        self._adjust_learning_rate(self.optimizer, epoch, self.args)
        return log

    def _valid_epoch(self, epoch):
        self.model.eval()
        model = self.model.module if hasattr(self.model, "module") else self.model
        n_dl = len(self.valid_data_loader)
        total_val_loss = [0] * n_dl
        text_arr = {i: [] for i in range(n_dl)}
        vid_arr = {i: [] for i in range(n_dl)}
        world = self.args.world_size
        with torch.no_grad():
            for dl_idx, dl in enumerate(self.valid_data_loader):
                for data in dl:
                    data = self._to_device(data)
                    text_embed, vid_embed = model(data, return_embeds=True)
                    if world > 1 and dist.is_initialized():
                        t_all = torch.empty((world * text_embed.shape[0], text_embed.shape[1]), device=self.device)
                        v_all = torch.empty_like(t_all)
                        dist.all_gather_into_tensor(t_all, text_embed.contiguous())
                        dist.all_gather_into_tensor(v_all, vid_embed.contiguous())
                    else:
                        t_all, v_all = text_embed, vid_embed
                    text_arr[dl_idx].append(t_all.cpu())
                    vid_arr[dl_idx].append(v_all.cpu())
                    total_val_loss[dl_idx] += self.loss(sim_matrix(t_all, v_all)).item()
        # nested_metrics[dl_idx][metric_name] = {R1, R5, ...} exactly as trainer_dist.py:251-279, so that
        # Multi_BaseTrainer_dist.train flattens it to the scalar keys val_{dl}_{metric}_{R1,...} a monitor can name
        nested_metrics = {i: {} for i in range(n_dl)}
        for dl_idx in range(n_dl):
            text_embeds = torch.cat(text_arr[dl_idx]).to(self.device)
            vid_embeds = torch.cat(vid_arr[dl_idx]).to(self.device)
            # the reference moves the matrix to the host first (trainer_dist.py:255-264); the metric functions here count
            # the ranks on the device for a CUDA tensor (same numbers, model/metric.py)
            sims = sim_matrix(text_embeds, vid_embeds).detach()
            for metric in self.metrics:
                nested_metrics[dl_idx][metric.__name__] = {k: float(v) for k, v in metric(sims).items()}
        log = {f'val_loss_{i}': total_val_loss[i] / max(1, len(self.valid_data_loader[i])) for i in range(n_dl)}
        log['nested_val_metrics'] = nested_metrics
        return log
