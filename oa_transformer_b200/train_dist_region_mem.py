"""Entry point of the region-sensitive variant, wired like OATrans/train_dist_region_mem.py:4-9
(`import model.oa_model_region_mem as module_arch`, `from trainer.trainer_region_mem import Multi_Trainer_dist`):

    python -m oa_transformer_b200.train_dist_region_mem -c oa_transformer_b200/configs/pt/cc3m_webvid/synthetic-region-mem.json
"""
from .model import oa_model_region_mem as module_arch
from .train_dist_multi import main as _main
from .trainer.trainer_region_mem import Multi_Trainer_dist


def main(argv=None):
    _main(argv, arch_module=module_arch, trainer_cls=Multi_Trainer_dist)


if __name__ == '__main__':
    main()
