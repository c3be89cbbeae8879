"""String-named optimiser factories with the reference's flags (pixelssl/nn/optimizer.py:19-75).
Only ``sgd`` is on the sseg hot path; ``adam`` is provided for the GCT/AdvSSL auxiliary nets."""
import torch.optim as optim

from ..utils import cmd

VALID_OPTIMIZER = ['sgd', 'adam']


def add_parser_arguments(parser):
    parser.add_argument('--lr', type=float, default=-1, metavar='')
    parser.add_argument('--dampening', type=float, default=-1, metavar='')
    parser.add_argument('--nesterov', type=cmd.str2bool, default=False, metavar='')
    parser.add_argument('--weight-decay', type=float, default=-1, metavar='')
    parser.add_argument('--momentum', type=float, default=-1, metavar='')
    parser.add_argument('--alpha', type=float, default=-1, metavar='')
    parser.add_argument('--centered', type=cmd.str2bool, default=False, metavar='')
    parser.add_argument('--eps', type=float, default=-1, metavar='')
    parser.add_argument('--beta1', type=float, default=-1, metavar='')
    parser.add_argument('--beta2', type=float, default=-1, metavar='')
    parser.add_argument('--amsgrad', type=cmd.str2bool, default=False, metavar='')


def _default(value, fallback):
    return fallback if value == -1 else value


def sgd(args):
    args.lr = _default(args.lr, 0.01)
    args.weight_decay = _default(args.weight_decay, 0)
    args.momentum = _default(args.momentum, 0)
    args.dampening = _default(args.dampening, 0)

    def sgd_wrapper(param_groups):
        return optim.SGD(param_groups, lr=args.lr, momentum=args.momentum, dampening=args.dampening,
                         weight_decay=args.weight_decay, nesterov=bool(args.nesterov))
    return sgd_wrapper


def adam(args):
    args.lr = _default(args.lr, 0.001)
    args.beta1 = _default(args.beta1, 0.9)
    args.beta2 = _default(args.beta2, 0.999)
    args.eps = _default(args.eps, 1e-08)
    args.weight_decay = _default(args.weight_decay, 0.0)

    def adam_wrapper(param_groups):
        return optim.Adam(param_groups, lr=args.lr, betas=(args.beta1, args.beta2), eps=args.eps,
                          weight_decay=args.weight_decay)
    return adam_wrapper
