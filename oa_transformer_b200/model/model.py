"""`from OATrans.model.model import sim_matrix` is what the reference trainer imports (trainer/trainer_dist.py:5);
the function keeps that name and signature (model/model.py:164-172) and runs on liboat."""
from ..functional import sim_matrix as _sim_matrix


def sim_matrix(a, b, eps=1e-8):
    """Cosine similarity of every row of a (text) with every row of b (video); norms are clamped at eps."""
    return _sim_matrix(a, b, eps)


from .oa_model import FrozenInTime  # noqa: E402,F401  (the reference module also defines the dual encoder)
