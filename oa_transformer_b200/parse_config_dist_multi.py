"""ConfigParser with the contract of OATrans/parse_config_dist_multi.py:13-150: JSON config + CLI overrides, run
directories, and the reflection factory `initialize(name, module, *args, index=None, **kwargs)` that instantiates
`getattr(module, cfg[name]['type'])(*args, **cfg[name]['args'])`, injecting constructor parameters that are missing
from the sub-dict but present at the top level of the config (and `args` for the classes that take the CLI namespace).
Same assertions: a config (-c) or a resume path is required; kwargs may not overwrite config entries."""
import inspect
import logging
import os
from datetime import datetime
from functools import reduce
from operator import getitem
from pathlib import Path

from .logger import setup_logging
from .utils import read_json, write_json

_TAKES_CLI_ARGS = ("FrozenInTime", "MultiDistTextObjectVideoDataLoader", "TextObjectVideoDataLoader")


class ConfigParser:
    def __init__(self, args, options='', timestamp=True, test=False):
        for opt in options:
            args.add_argument(*opt.flags, default=None, type=opt.type)
        args = args.parse_args()
        self.args = args
        if getattr(args, "device", None):
            os.environ["CUDA_VISIBLE_DEVICES"] = args.device
        if args.resume is None:
            assert args.config is not None, \
                "Configuration file need to be specified. Add '-c config.json', for example."
            self.cfg_fname = Path(args.config)
            config = read_json(self.cfg_fname)
            self.resume = None
        else:
            self.resume = Path(args.resume)
            config = read_json(self.resume.parent / 'config.json')
            if args.config is not None:
                config.update(read_json(Path(args.config)))
        self._config = _update_config(config, options, args)

        save_dir = Path(self.config['trainer']['save_dir'])
        stamp = datetime.now().strftime(r'%m%d_%H%M%S') if timestamp else ''
        name = self.config['name']
        self._save_dir = save_dir / 'models' / name / stamp
        self._web_log_dir = save_dir / 'web' / name / stamp
        self._log_dir = save_dir / 'log' / name / stamp
        self.log_levels = {0: logging.WARNING, 1: logging.INFO, 2: logging.DEBUG}
        if not test:
            self.save_dir.mkdir(parents=True, exist_ok=True)
            self.log_dir.mkdir(parents=True, exist_ok=True)
            write_json(self.config, self.save_dir / 'config.json')
            setup_logging(self.log_dir)

    def initialize(self, name, module, *args, index=None, **kwargs):
        if index is None:
            module_name = self[name]['type']
            module_args = dict(self[name]['args'])
            assert all(k not in module_args for k in kwargs), 'Overwriting kwargs given in config file is not allowed'
            module_args.update(kwargs)
        else:
            module_name = self[name][index]['type']
            module_args = dict(self[name][index]['args'])
        cls = getattr(module, module_name)
        for param in inspect.signature(cls.__init__).parameters.keys():
            if param not in module_args and param in self.config:
                module_args[param] = self[param]
            if module_name in _TAKES_CLI_ARGS and param == 'args':
                module_args[param] = self.args
        return cls(*args, **module_args)

    def __getitem__(self, name):
        return self.config[name]

    def get_logger(self, name, verbosity=2):
        assert verbosity in self.log_levels, \
            'verbosity option {} is invalid. Valid options are {}.'.format(verbosity, self.log_levels.keys())
        logger = logging.getLogger(name)
        logger.setLevel(self.log_levels[verbosity])
        return logger

    @property
    def config(self):
        return self._config

    @property
    def save_dir(self):
        return self._save_dir

    @property
    def log_dir(self):
        return self._log_dir


def _update_config(config, options, args):
    for opt in options:
        value = getattr(args, _get_opt_name(opt.flags))
        if value is not None:
            _set_by_path(config, opt.target, value)
    return config


def _get_opt_name(flags):
    for flg in flags:
        if flg.startswith('--'):
            return flg.replace('--', '')
    return flags[0].replace('--', '')


def _set_by_path(tree, keys, value):
    reduce(getitem, keys[:-1], tree)[keys[-1]] = value
