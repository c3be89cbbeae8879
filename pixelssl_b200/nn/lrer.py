"""LR schedulers (pixelssl/nn/lrer.py).  ``polynomiallr`` is the one the sseg scripts use (per iteration, host scalar
math); the four per-epoch wrappers of torch's own schedulers are here so that every ``lrers`` entry a PixelSSL script
can name resolves.  All of them only edit ``optimizer.param_groups[i]['lr']``, which the fused arena optimiser
steps read at launch time."""
import math

import torch
from torch.optim.lr_scheduler import _LRScheduler

from ..utils import cmd

EPOCH_LRERS = ['steplr', 'multisteplr', 'exponentiallr', 'cosineannealinglr']
ITER_LRERS = ['polynomiallr']
VALID_LRER = EPOCH_LRERS + ITER_LRERS


def add_parser_arguments(parser):
    parser.add_argument('--last-epoch', type=int, default=-1, metavar='')
    parser.add_argument('--step-size', type=int, default=-1, metavar='')
    parser.add_argument('--milestones', type=cmd.str2intlist, default=[], metavar='')
    parser.add_argument('--gamma', type=float, default=-1, metavar='')
    parser.add_argument('--T-max', type=int, default=-1, metavar='')
    parser.add_argument('--eta-min', type=float, default=-1, metavar='')
    parser.add_argument('--power', type=float, default=-1, metavar='')


class PolynomialLR(_LRScheduler):
    """lr_g = base_g * (1 - cur_iter / max_iters) ** power, advanced once per iteration
    (lrer.py:143-179).  As in the reference, ``_LRScheduler.__init__`` performs one ``step()``,
    so the first optimiser step already runs with cur_iter == 1."""

    def __init__(self, optimizer, epochs, iters_per_epoch, power=0.9, last_epoch=-1):
        self.epochs, self.iters_per_epoch = epochs, iters_per_epoch
        self.max_iters = epochs * iters_per_epoch
        self.cur_iter = 0
        self.power = power
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
        factor = (1 - float(self.cur_iter) / self.max_iters) ** self.power
        return [base * factor for base in self.base_lrs]

    def step(self, epoch=None):
        if epoch is None:
            self.cur_iter += 1
            self.last_epoch = math.floor(self.cur_iter / self.iters_per_epoch)
        elif epoch != 0:
            self.last_epoch = epoch
            assert self.last_epoch <= self.epochs
            self.cur_iter = self.last_epoch * self.iters_per_epoch
        for group, lr in zip(self.optimizer.param_groups, self.get_lr()):
            group['lr'] = lr


def polynomiallr(args):
    args.power = 0.9 if args.power == -1 else args.power

    def polynomiallr_wrapper(optimizer):
        return PolynomialLR(optimizer, epochs=args.epochs, iters_per_epoch=args.iters_per_epoch,
                            power=args.power, last_epoch=args.last_epoch)
    return polynomiallr_wrapper


def _epoch_lrer(scheduler_cls, resolve):
    """Export function for a torch per-epoch scheduler: ``resolve(args)`` replaces the parser's -1 / [] placeholders
    by the reference's per-scheduler defaults (lrer.py:51-119) and returns the scheduler's keyword arguments."""
    def export(args):
        kwargs = resolve(args)

        def wrapper(optimizer):
            return scheduler_cls(optimizer, last_epoch=args.last_epoch, **kwargs)
        return wrapper
    return export


def _pick(args, name, default, unset=-1):
    if getattr(args, name) == unset:
        setattr(args, name, default)
    return getattr(args, name)


steplr = _epoch_lrer(torch.optim.lr_scheduler.StepLR, lambda a: {
    'step_size': _pick(a, 'step_size', a.epochs), 'gamma': _pick(a, 'gamma', 0.1)})
multisteplr = _epoch_lrer(torch.optim.lr_scheduler.MultiStepLR, lambda a: {
    'milestones': _pick(a, 'milestones', list(range(1, a.epochs)), unset=[]), 'gamma': _pick(a, 'gamma', 0.1)})
exponentiallr = _epoch_lrer(torch.optim.lr_scheduler.ExponentialLR, lambda a: {'gamma': _pick(a, 'gamma', 0.1)})
cosineannealinglr = _epoch_lrer(torch.optim.lr_scheduler.CosineAnnealingLR, lambda a: {
    'T_max': _pick(a, 'T_max', a.epochs), 'eta_min': _pick(a, 'eta_min', 0)})
