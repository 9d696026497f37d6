"""BaseModel of OATrans/base/base_model.py: an nn.Module whose str() reports the trainable parameter count."""
import numpy as np
import torch.nn as nn


class BaseModel(nn.Module):
    def forward(self, *inputs):
        raise NotImplementedError

    def __str__(self):
        n = sum(int(np.prod(p.size())) for p in self.parameters() if p.requires_grad)
        return super().__str__() + "\nTrainable parameters: {}".format(n)
