"""utils/param_forzen.py:4-10: linear evaluation - only the two projections stay trainable."""


def forzen_param(model):
    for name, param in model.named_parameters():
        param.requires_grad = ('vid_proj' in name or 'txt_proj' in name)
    return True
