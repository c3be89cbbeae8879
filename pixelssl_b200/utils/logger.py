"""Logging + meters with the reference's user-visible behaviour (pixelssl/utils/logger.py:37-131):
same 'PixelSSL' logger, same banners, ``log_err`` terminates the process, meters keep device
tensors and only synchronise when formatted."""
import logging
import sys

logging.basicConfig(level=logging.INFO, format='%(message)s')
logger = logging.getLogger('PixelSSL')
_BAR = 78


def _join(message):
    return ''.join(message) if isinstance(message, (list, tuple)) else message


def log_info(message):
    logger.info(_join(message))


def _banner(tag, body):
    side = (_BAR - len(tag) - 2) // 2
    head = '=' * side + ' ' + tag + ' ' + '=' * (_BAR - side - len(tag) - 2)
    return '\n%s\n%s%s\n' % (head, body, '=' * _BAR)


def log_warn(message):
    logger.warning(_banner('WARN', _join(message)))


def log_err(message):
    """Error convention of the plugin boundary (utils/logger.py:58-67): print and exit."""
    logger.error(_banner('ERROR', _join(message)))
    sys.exit()


class AvgMeter:
    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum = self.sum + val * n
        self.count += n
        self.avg = self.sum / self.count

    def __format__(self, spec):
        return '{0:{2}} ({1:{2}})'.format(float(self.val), float(self.avg), spec)


class AvgMeterSet:
    def __init__(self):
        self.meters = {}

    def __getitem__(self, key):
        return self.meters[key]

    def keys(self):
        return self.meters.keys()

    def has_key(self, key):
        return key in self.meters

    def update(self, name, value, n=1):
        self.meters.setdefault(name, AvgMeter()).update(value, n)

    def reset(self, name=None):
        if name is None:
            for m in self.meters.values():
                m.reset()
        elif name in self.meters:
            self.meters[name].reset()
        else:
            log_err('Unknown key value for AvgMeterSet: {0}\n'.format(name))

    def values(self, postfix=''):
        return {k + postfix: m.val for k, m in self.meters.items()}

    def averages(self, postfix='/avg'):
        return {k + postfix: m.avg for k, m in self.meters.items()}

    def sums(self, postfix='/sum'):
        return {k + postfix: m.sum for k, m in self.meters.items()}

    def counts(self, postfix='/count'):
        return {k + postfix: m.count for k, m in self.meters.items()}
