"""Entry point with the CLI and wiring of OATrans/train_dist_multi.py (launched one process per GPU by torchrun /
torch.distributed.launch: reads MASTER_ADDR / MASTER_PORT / WORLD_SIZE / RANK / LOCAL_RANK from the environment,
:127-132), resolving the three module handles the way the reference's sibling scripts do
(train_dist_multi_global_local.py:4-7): model.loss, model.metric and the arch module.

    python -m oa_transformer_b200.train_dist_multi -c oa_transformer_b200/configs/pt/cc3m_webvid/norm.json
"""
import argparse
import collections
import os

import torch

from . import data_loader as module_data
from . import model as module_arch
from . import model as module_loss
from . import model as module_metric
from .parse_config_dist_multi import ConfigParser
from .trainer.trainer_dist import Multi_Trainer_dist
from .utils import replace_nested_dict_item


def init_dataloaders(config, module):
    if "type" in config["data_loader"] and "args" in config["data_loader"]:
        train = [config.initialize("data_loader", module)]
        config['data_loader']['args'] = replace_nested_dict_item(config['data_loader']['args'], 'split', 'val')
        valid = [config.initialize("data_loader", module)]
    elif isinstance(config["data_loader"], list):
        n = len(config['data_loader'])
        train = [config.initialize('data_loader', module, index=i) for i in range(n)]
        for dl_cfg in config['data_loader']:
            dl_cfg['args'] = replace_nested_dict_item(dl_cfg['args'], 'split', 'val')
        valid = [config.initialize('data_loader', module, index=i) for i in range(n)]
    else:
        raise ValueError("Check data_loader config, not correct format.")
    return train, valid


def run(config, args, arch_module=None, trainer_cls=None):
    """arch_module / trainer_cls: the variant entry points swap them (train_dist_region_mem.py)."""
    module_arch_ = arch_module or module_arch
    trainer_cls = trainer_cls or Multi_Trainer_dist
    logger = config.get_logger('train')
    os.environ['TOKENIZERS_PARALLELISM'] = "false"
    torch.cuda.set_device(args.local_rank)
    if args.world_size > 1:
        torch.distributed.init_process_group(backend='nccl', init_method='tcp://{}:{}'.format(
            args.master_address, args.master_port), rank=args.rank, world_size=args.world_size)
    tokenizer = None
    text_model = config['arch']['args']['text_params']['model']
    if os.path.isdir(text_model):
        import transformers
        tokenizer = transformers.AutoTokenizer.from_pretrained(text_model)
    data_loader, valid_data_loader = init_dataloaders(config, module_data)
    model = config.initialize('arch', module_arch_)
    if args.local_rank == 0:
        logger.info(model)
    loss = config.initialize(name="loss", module=module_loss)
    metrics = [getattr(module_metric, met) for met in config['metrics']]
    trainable = [p for p in model.parameters() if p.requires_grad]
    import transformers
    # `transformers.AdamW` (what the reference's configs name) is gone in transformers 5.x: the fused liboat AdamW keeps
    # its semantics; any other optimizer type resolves in transformers / torch.optim as before
    from oa_transformer_b200 import optim as oat_optim
    opt_module = oat_optim if hasattr(oat_optim, config['optimizer']['type']) else (
        transformers if hasattr(transformers, config['optimizer']['type']) else torch.optim)
    optimizer = config.initialize('optimizer', opt_module, trainable)
    trainer = trainer_cls(args, model, loss, metrics, optimizer, config=config, data_loader=data_loader,
                                 valid_data_loader=valid_data_loader, tokenizer=tokenizer,
                                 max_samples_per_epoch=config['trainer']['max_samples_per_epoch'])
    trainer.train()
    if args.world_size > 1:
        torch.distributed.destroy_process_group()


def main(argv=None, arch_module=None, trainer_cls=None):
    ap = argparse.ArgumentParser(description='OA-Transformer dual-encoder training (B200 path)')
    ap.add_argument('-c', '--config', default=None, type=str)
    ap.add_argument('-r', '--resume', default=None, type=str)
    ap.add_argument('-d', '--device', default=None, type=str)
    ap.add_argument('-o', '--observe', action='store_true')
    ap.add_argument('--launcher', choices=['none', 'pytorch'], default='none')
    ap.add_argument('-k', '--local_rank', type=int, default=int(os.environ.get('LOCAL_RANK', 0)))
    ap.add_argument('--master_address', default=os.environ.get('MASTER_ADDR', '127.0.0.1'))
    ap.add_argument('--master_port', type=int, default=int(os.environ.get('MASTER_PORT', 9999)))
    ap.add_argument('--world_size', type=int, default=int(os.environ.get('WORLD_SIZE', 1)))
    ap.add_argument('--rank', type=int, default=int(os.environ.get('RANK', 0)))
    ap.add_argument('--learning_rate1', type=float, default=2e-4)
    ap.add_argument('--schedule', default=[60, 80], nargs='*', type=int)
    CustomArgs = collections.namedtuple('CustomArgs', 'flags type target')
    options = [CustomArgs(['--lr', '--learning_rate'], type=float, target=('optimizer', 'args', 'lr')),
               CustomArgs(['--bs', '--batch_size'], type=int, target=('data_loader', 'args', 'batch_size'))]
    if argv is not None:
        import sys
        sys.argv = [sys.argv[0]] + list(argv)
    config = ConfigParser(ap, options)
    run(config, config.args, arch_module=arch_module, trainer_cls=trainer_cls)


if __name__ == '__main__':
    main()
