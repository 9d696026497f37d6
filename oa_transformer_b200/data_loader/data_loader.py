"""`from OATrans.data_loader import data_loader as module_data` (train_dist_multi.py:4): the loader classes by the names
the configs use."""
from . import (DevicePrefetcher, MultiDistTextObjectVideoDataLoader, SyntheticTextObjectVideoDataLoader,  # noqa: F401
               TextObjectVideoDataLoader)
