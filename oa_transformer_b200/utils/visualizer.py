"""`import utils.visualizer as module_vis` (train_dist_multi.py:7): the HTML ranking visualiser is outside the hot path;
every shipped config sets visualizer.type = "" so nothing is ever constructed from this module."""
