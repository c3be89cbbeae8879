"""Criterion plugin: CommonSSEGCriterion (task/sseg/criterion.py:18-38) on the fused CE kernel."""
import torch.nn as nn

from ... import ops
from ...utils import logger


def add_parser_arguments(parser):
    pass


def sseg_criterion():
    return CommonSSEGCriterion


class CommonSSEGCriterion(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self.ignore_index = args.ignore_index

    def forward(self, pred, gt, inp, mean_upstream=None):
        """-> per-sample loss Tensor[n].  ``mean_upstream`` (engine-only, optional): 1/n when the
        caller's next op is ``torch.mean`` feeding the final loss directly, which lets the
        gradient be written by the forward kernel."""
        if len(pred) != 1 or len(gt) != 1 or len(inp) != 1:
            logger.log_err('DeepLab criterion for semantic segmentation requires\t=>\t'
                           'len(pred) == 1 \t len(gt) == 1 \t len(inp) == 1\n')
        return ops.cross_entropy2d(pred[0], gt[0], self.ignore_index, upstream_const=mean_upstream)
