from .base_model import BaseModel  # noqa: F401
from .base_trainer import Multi_BaseTrainer_dist  # noqa: F401
