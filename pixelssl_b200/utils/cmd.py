"""Config-dict -> argv -> argparse, the reference's flag system (pixelssl/utils/cmd.py:10-61)."""
import re

from . import logger


def parse_args(parser, args_dict):
    argv = []
    for key, value in args_dict.items():
        flag = '-' + key if len(key) == 1 else '--' + re.sub(r'_', '-', key)
        argv += [flag, str(value)]
    return parser.parse_args(argv)


def str2bool(v):
    if v.lower() in ('yes', 'true', 't', 'y', '1'):
        return True
    if v.lower() in ('no', 'false', 'f', 'n', '0'):
        return False
    logger.log_err('str2bool requires a boolean value, but got {0}\n'.format(v))


def str2intlist(v):
    return [int(i.strip()) for i in re.sub(r'[\[\]()]', '', v).split(',')]


def str2floatlist(v):
    return [float(i.strip()) for i in re.sub(r'[\[\]()]', '', v).split(',')]
