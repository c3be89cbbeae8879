"""Host-side helpers of the step loop (pixelssl/nn/func.py): ramp-up, tuple slicing, model
creation.  ``create_model`` is where the reference's ``DataParallel(model).cuda()`` becomes
"one process per GPU": parameters are packed into one flat HBM arena (so SGD+EMA is a single
kernel and the gradient all-reduce is a single NCCL call on a contiguous buffer)."""
import math

import torch

from ..utils import logger
from .arena import EngineParallel


def sigmoid_rampup(current, rampup_length):
    """exp(-5 (1 - clip(current, 0, L)/L)^2); 1.0 when L == 0  (nn/func.py:12-20)."""
    if rampup_length == 0:
        return 1.0
    current = min(max(float(current), 0.0), float(rampup_length))
    phase = 1.0 - current / rampup_length
    return float(math.exp(-5.0 * phase * phase))


def split_tensor_tuple(ttuple, start, end, reduce_dim=False):
    """Slice every tensor of the tuple along the batch dim (nn/func.py:24-51)."""
    if reduce_dim:
        assert end - start == 1
    if reduce_dim and end - start == 1:
        return tuple(t[start, ...] for t in ttuple)
    return tuple(t[start:end, ...] for t in ttuple)


def create_model(mclass, mname, **kwargs):
    """nn/func.py:54-62.  Returns an ``EngineParallel`` whose ``.module`` is the task model, so
    ``state_dict()`` keys keep the reference's ``module.`` prefix (checkpoint compatible)."""
    model = mclass(**kwargs)
    model = EngineParallel(model).cuda()
    logger.log_info('  ' + '=' * 76 + '\n  {0} parameters \n{1}'.format(mname, model_str(model)))
    return model


def model_str(module):
    row = '  {name:<40} {shape:>20} = {total_size:>12,d}'
    lines = ['  ' + '-' * 76]
    params = list(module.named_parameters())
    for name, p in params:
        lines.append(row.format(name=name, shape=' * '.join(str(s) for s in p.size()), total_size=p.numel()))
    lines.append('  ' + '-' * 76)
    lines.append(row.format(name='all parameters', shape='sum of above',
                            total_size=sum(int(p.numel()) for _, p in params)))
    lines.append('  ' + '=' * 76)
    lines.append('')
    return '\n'.join(lines)


def pytorch_support(required_version='1.0.0', info_str=''):
    return True
