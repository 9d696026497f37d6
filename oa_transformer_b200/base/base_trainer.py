"""Multi_BaseTrainer_dist with the behaviour of OATrans/base/base_trainer.py:7-244 that the hot path touches: device
placement, DistributedDataParallel wrap when more than one GPU is configured (find_unused_parameters=True, :23), the
epoch loop with monitor / best tracking, and checkpoints with the reference's dictionary layout
({'arch','epoch','state_dict','optimizer','monitor_best','config'}, :163-186) and resume (:188-244)."""
import torch
from numpy import inf

from ..utils import state_dict_data_parallel_fix


class Multi_BaseTrainer_dist:
    def __init__(self, args, model, loss, metrics, optimizer, config, writer=None, init_val=False):
        self.config = config
        self.logger = config.get_logger('trainer', config['trainer']['verbosity'])
        self.args = args
        self.device, device_ids = self._prepare_device(config['n_gpu'])
        self.model = model.to(self.device)
        self.model.device = self.device
        if len(device_ids) > 1 and torch.distributed.is_available() and torch.distributed.is_initialized():
            self.model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[args.local_rank],
                                                                   find_unused_parameters=True)
        self.loss = loss.to(self.device)
        self.metrics = metrics
        self.optimizer = optimizer
        cfg = config['trainer']
        self.epochs = cfg['epochs']
        self.save_period = cfg['save_period']
        self.monitor = cfg.get('monitor', 'off')
        self.init_val = cfg.get('init_val', True)
        if self.monitor == 'off':
            self.mnt_mode, self.mnt_best = 'off', 0
        else:
            self.mnt_mode, self.mnt_metric = self.monitor.split()
            assert self.mnt_mode in ['min', 'max']
            self.mnt_best = inf if self.mnt_mode == 'min' else -inf
            self.early_stop = cfg.get('early_stop', inf)
        self.start_epoch = 1
        self.checkpoint_dir = config.save_dir
        self.writer = writer
        if config.resume is not None:
            self._resume_checkpoint(config.resume)

    def _train_epoch(self, epoch):
        raise NotImplementedError

    def _valid_epoch(self, epoch):
        raise NotImplementedError

    def train(self):
        if self.init_val:
            self._valid_epoch(-1)
        for epoch in range(self.start_epoch, self.epochs + 1):
            result = self._train_epoch(epoch)
            log = {'epoch': epoch}
            for key, value in result.items():
                if key == 'nested_val_metrics':
                    for subkey, subval in value.items():
                        for subsubkey, subsubval in subval.items():
                            for k4, v4 in subsubval.items():
                                log[f"val_{subkey}_{subsubkey}_{k4}"] = v4
                else:
                    log[key] = value
            if self.args.rank == 0:
                for key, value in log.items():
                    self.logger.info('    {:15s}: {}'.format(str(key), value))
            best = False
            if self.mnt_mode != 'off' and self.args.rank == 0 and self.mnt_metric not in log:
                # base_trainer.py:117-121: a monitor that names no logged scalar switches monitoring off, loudly
                self.logger.warning("Warning: Metric '{}' is not found. Model performance monitoring is disabled."
                                    .format(self.mnt_metric))
                self.mnt_mode = 'off'
            if self.mnt_mode != 'off' and self.mnt_metric in log:
                improved = (self.mnt_mode == 'min' and log[self.mnt_metric] <= self.mnt_best) or \
                           (self.mnt_mode == 'max' and log[self.mnt_metric] >= self.mnt_best)
                if improved:
                    self.mnt_best = log[self.mnt_metric]
                    best = True
            if self.args.rank == 0 and (epoch % self.save_period == 0 or best):
                self._save_checkpoint(epoch, save_best=best)

    def _prepare_device(self, n_gpu_use):
        n_gpu = torch.cuda.device_count()
        if n_gpu_use > 0 and n_gpu == 0:
            raise RuntimeError("no CUDA device: the liboat kernels are the only implementation of this path")
        local = getattr(self.args, "local_rank", 0)
        device = torch.device('cuda:%d' % local)
        return device, list(range(n_gpu_use))

    def _save_checkpoint(self, epoch, save_best=False):
        model = self.model.module if hasattr(self.model, "module") else self.model
        state = {'arch': type(model).__name__, 'epoch': epoch, 'state_dict': self.model.state_dict(),
                 'optimizer': self.optimizer.state_dict(), 'monitor_best': self.mnt_best, 'config': self.config.config}
        filename = str(self.checkpoint_dir / 'checkpoint-epoch{}.pth'.format(epoch))
        torch.save(state, filename)
        self.logger.info("Saving checkpoint: {} ...".format(filename))
        if save_best:
            torch.save(state, str(self.checkpoint_dir / 'model_best.pth'))

    def _resume_checkpoint(self, resume_path):
        # weights_only=False: reference-produced checkpoints pickle the config dictionary next to the tensors
        # (base_trainer.py:163-175); checkpoints are trusted local files here, exactly as in the reference
        checkpoint = torch.load(str(resume_path), map_location="cpu", weights_only=False)
        self.start_epoch = checkpoint['epoch'] + 1
        self.mnt_best = checkpoint['monitor_best']
        state_dict = state_dict_data_parallel_fix(checkpoint['state_dict'], self.model.state_dict())
        self.model.load_state_dict(state_dict)
        if checkpoint['config']['optimizer']['type'] == self.config['optimizer']['type']:
            self.optimizer.load_state_dict(checkpoint['optimizer'])
        self.logger.info("Checkpoint loaded. Resume training from epoch {}".format(self.start_epoch))
