"""`OATrans` import overlay: put THIS directory's parent (oa_transformer_b200/overlay) in front of the reference on
PYTHONPATH and the reference's UNMODIFIED entry scripts resolve the hot path to liboat:

    cd /path/to/OA-Transformer/OATrans
    PYTHONPATH=/path/to/repo:/path/to/repo/oa_transformer_b200/overlay python train_dist_multi.py -c configs/pt/cc3m_webvid/norm.json

What the scripts import and what they get (reference file:line -> module here):
    train_dist_multi.py:4   from OATrans.data_loader import data_loader as module_data   -> oa_transformer_b200.data_loader.data_loader
    train_dist_multi.py:5   from OATrans import model as module_loss / module_metric / module_arch
                                                                                         -> oa_transformer_b200.model
                            (the reference's own OATrans/model/__init__.py is empty, so `FrozenInTime` does not even
                             resolve there - SURVEY.md fact 4)
    train_dist_multi.py:7-10,15  import utils.visualizer / utils.util / utils.param_forzen,
                                 from parse_config_dist_multi import ConfigParser,
                                 from trainer.trainer_dist import Multi_Trainer_dist       -> the mirrors of the same names
    trainer/trainer_dist.py:3-5  from OATrans.base / OATrans.utils / OATrans.model.model  -> idem
    train_dist_region_mem.py:4-9 model.oa_model_region_mem, trainer.trainer_region_mem   -> idem
    train_dist_multi.py:11-12    sacred / neptunecontrib (experiment logging services, outside the path): import-only
                                 stand-ins are registered when the packages are absent.
    train_dist_multi.py:66       config.initialize('optimizer', transformers, ...) with type "AdamW": transformers 5.x
                                 dropped AdamW; `transformers.AdamW` is pointed at the fused liboat AdamW (same update).
The cwd-relative names (utils, trainer, model, base, logger, data_loader, parse_config_dist_multi) are registered in
sys.modules when this package is imported, which happens on the scripts' first OATrans import - before Python would
look them up in the reference's own directories."""
import importlib
import importlib.machinery
import sys
import types

import oa_transformer_b200 as _pkg

_SUBPACKAGES = ("model", "trainer", "base", "utils", "logger", "data_loader", "optim")
_SUBMODULES = ("parse_config_dist_multi", "model.model", "model.loss", "model.metric", "model.oa_model",
               "model.video_transformer", "model.oa_model_region_mem", "model.oa_video_transformer_region",
               "trainer.trainer_dist", "trainer.trainer_region_mem", "trainer.trainer_global_local",
               "utils.util", "utils.visualizer", "utils.param_forzen", "data_loader.data_loader", "base.base_trainer",
               "base.base_model")


def install():
    me = sys.modules[__name__]
    for name in _SUBPACKAGES + _SUBMODULES:
        mod = importlib.import_module("oa_transformer_b200." + name)
        for alias in ("OATrans." + name, name):                 # package-absolute and cwd-relative spellings
            sys.modules[alias] = mod
        if "." not in name:
            setattr(me, name, mod)
    for name, attrs in (("sacred", {"Experiment": _Experiment}),
                        ("neptunecontrib", {}), ("neptunecontrib.monitoring", {}),
                        ("neptunecontrib.monitoring.sacred", {"NeptuneObserver": _Observer})):
        try:
            importlib.import_module(name)
        except Exception:
            m = types.ModuleType(name)
            m.__spec__ = importlib.machinery.ModuleSpec(name, None)
            m.__dict__.update(attrs)
            sys.modules[name] = m
    import transformers
    if not hasattr(transformers, "AdamW"):
        from oa_transformer_b200.optim import AdamW
        transformers.AdamW = AdamW
    return _pkg


class _Experiment:
    """Import-only stand-in for sacred.Experiment: `@ex.main` returns the function, `ex.run()` calls it."""

    def __init__(self, *a, **k):
        self.observers, self._main = [], None

    def main(self, fn):
        self._main = fn
        return fn

    def add_config(self, *a, **k):
        pass

    def log_scalar(self, *a, **k):
        pass

    def run(self):
        return self._main()


class _Observer:
    def __init__(self, *a, **k):
        pass


install()
