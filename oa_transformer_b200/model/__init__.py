"""Plugin surface of OATrans/model/* for the dual-encoder hot path.

`config.initialize('arch', module_arch)` / `('loss', module_loss)` / metric lookups resolve against this package
(the reference's own OATrans/model/__init__.py is empty, which is why its train_dist_multi.py cannot resolve
FrozenInTime - SURVEY.md fact 4; the sibling entry scripts import model.loss / model.metric / model.oa_model)."""
from .loss import NormSoftmaxLoss  # noqa: F401
from .metric import t2v_metrics, v2t_metrics  # noqa: F401
from .model import sim_matrix  # noqa: F401
from .oa_model import FrozenInTime  # noqa: F401
from .video_transformer import SpaceTimeTransformer  # noqa: F401
