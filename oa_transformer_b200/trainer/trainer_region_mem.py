"""Multi_Trainer_dist of OATrans/trainer/trainer_region_mem.py: the dual-encoder step plus the region loss.

    text_embeds, video_embeds, region_sim = self.model(data, aug=True)                     (:150)
    four AllGather_multi: video / text embeddings, region_sim (B, 5, L), patch_mask (B, 5, L)  (:151-155)
    loss = NormSoftmaxLoss(sim_matrix(text, video)) + 0.1 * BCELoss(sum)(region_sim, patch_mask) / rows   (:157-167)

`region_loss` below is that second term as one liboat launch (oat_bce_sum)."""
import time

import torch
import torch.distributed as dist

from .. import functional as OF
from ..model.model import sim_matrix
from .trainer_dist import Multi_Trainer_dist as _Base


def region_loss(region_sim, patch_mask, weight=0.1):
    """0.1 * nn.BCELoss(reduction='sum')(region_sim.view(-1, L), patch_mask.view(-1, L)) / rows
    (trainer/trainer_region_mem.py:161-167)."""
    L = region_sim.size(-1)
    rows = region_sim.numel() // L
    return OF.bce_sum(region_sim.reshape(-1, L), patch_mask.reshape(-1, L).float(), weight / rows)


class Multi_Trainer_dist(_Base):
    def _to_device(self, data):
        data = super()._to_device(data)
        data['text_region_embedding'] = data['text_region_embedding'].to(self.device)
        data['patch_masks'] = data['patch_masks'].to(self.device)
        return data

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
                patch_mask = data['patch_masks'].squeeze(1).float()
                self.optimizer.zero_grad()
                with torch.set_grad_enabled(True):
                    text_embeds, video_embeds, region_sim = self.model(data, aug=True)
                    video_embeds = self.allgather(video_embeds, self.n_gpu, self.args)
                    text_embeds = self.allgather(text_embeds, self.n_gpu, self.args)
                    region_sim = self.allgather(region_sim, self.n_gpu, self.args)
                    patch_mask = self.allgather(patch_mask, self.n_gpu, self.args)
                    output = sim_matrix(text_embeds, video_embeds)
                    loss = self.loss(output)
                    r_loss = region_loss(region_sim, patch_mask)
                    loss = loss + r_loss
                loss.backward()
                self.optimizer.step()
                value = loss.detach().item()
                total_loss[dl_idx] += value
                if self.writer is not None and self.args.rank == 0:
                    self.writer.log_scalar(f'loss_train_{dl_idx}', value)
                if batch_idx % self.log_step == 0 and self.args.local_rank == 0:
                    self.logger.debug('Train Epoch: {} dl{} [{}/{}] Loss: {:.6f} (t2v {:.6f} region {:.6f}, {:.2f}s)'.format(
                        epoch, dl_idx, batch_idx, self.len_epoch, value, value - r_loss.item(), r_loss.item(),
                        time.time() - begin))
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
        """trainer_region_mem.py:205-300: retrieval metrics on the gathered embeddings; the per-batch loss adds
        BCE(sum)(region_sim, patch_masks) / batch (:260)."""
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
                    text_embed, vid_embed, region_sim = model(data)
                    if world > 1 and dist.is_initialized():
                        t_all = torch.empty((world * text_embed.shape[0], text_embed.shape[1]), device=self.device)
                        v_all = torch.empty_like(t_all)
                        dist.all_gather_into_tensor(t_all, text_embed.contiguous())
                        dist.all_gather_into_tensor(v_all, vid_embed.contiguous())
                    else:
                        t_all, v_all = text_embed, vid_embed
                    text_arr[dl_idx].append(t_all.cpu())
                    vid_arr[dl_idx].append(v_all.cpu())
                    loss = self.loss(sim_matrix(t_all, v_all))
                    masks = data['patch_masks'].squeeze(1).float()
                    loss = loss + OF.bce_sum(region_sim, masks, 1.0 / region_sim.size(0))
                    total_val_loss[dl_idx] += loss.item()
        nested_metrics = {i: {} for i in range(n_dl)}
        for dl_idx in range(n_dl):
            sims = sim_matrix(torch.cat(text_arr[dl_idx]).to(self.device),
                              torch.cat(vid_arr[dl_idx]).to(self.device)).detach()
            for metric in self.metrics:
                nested_metrics[dl_idx][metric.__name__] = {k: float(v) for k, v in metric(sims).items()}
        log = {f'val_loss_{i}': total_val_loss[i] / max(1, len(self.valid_data_loader[i])) for i in range(n_dl)}
        log['nested_val_metrics'] = nested_metrics
        return log
