"""Helpers of OATrans/utils/util.py that the plugin surface needs: JSON io, infinite loader loop, DP-prefix fix."""
import json
from collections import OrderedDict
from itertools import repeat
from pathlib import Path



def read_json(fname):
    with Path(fname).open("rt") as handle:
        return json.load(handle, object_hook=OrderedDict)


def write_json(content, fname):
    with Path(fname).open("wt") as handle:
        json.dump(content, handle, indent=4, sort_keys=False)


def inf_loop(data_loader):
    """Endless iteration over a loader (utils/util.py)."""
    for loader in repeat(data_loader):
        yield from loader


def replace_nested_dict_item(obj, key, replace_value):
    for k, v in obj.items():
        if isinstance(v, dict):
            obj[k] = replace_nested_dict_item(v, key, replace_value)
    if key in obj:
        obj[key] = replace_value
    return obj


def state_dict_data_parallel_fix(load_state_dict, curr_state_dict):
    """utils/util.py:24-50: reconcile the 'module.' prefix between a checkpoint and the current model."""
    load_keys, curr_keys = list(load_state_dict.keys()), list(curr_state_dict.keys())
    if not curr_keys[0].startswith('module.') and load_keys[0].startswith('module.'):
        return {k[7:]: v for k, v in load_state_dict.items()}
    if curr_keys[0].startswith('module.') and not load_keys[0].startswith('module.'):
        return {'module.' + k: v for k, v in load_state_dict.items()}
    return load_state_dict
