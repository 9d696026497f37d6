"""NormSoftmaxLoss with the reference's constructor and call contract (OATrans/model/loss.py:7-25), computed by
liboat (row/column log-sum-exp, diagonal means and dL/dsims in two launches)."""
import torch.nn as nn

from ..functional import norm_softmax_loss


class NormSoftmaxLoss(nn.Module):
    def __init__(self, temperature=0.05):
        super().__init__()
        self.temperature = temperature

    def forward(self, x):
        """x: square similarity matrix in [-1, 1] (rows = text, columns = video) on a CUDA device."""
        return norm_softmax_loss(x, self.temperature)
