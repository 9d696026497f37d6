"""`from utils.util import replace_nested_dict_item` (train_dist_multi.py:8) and friends."""
from . import *  # noqa: F401,F403
from . import inf_loop, read_json, replace_nested_dict_item, state_dict_data_parallel_fix, write_json  # noqa: F401
