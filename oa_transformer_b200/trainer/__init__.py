from .trainer_dist import AllGather_multi, Multi_Trainer_dist  # noqa: F401
