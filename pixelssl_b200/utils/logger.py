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


class _HostMeter:
    """val / avg of a meter as host floats (same format as AvgMeter)."""

    def __init__(self, val, avg):
        self.val, self.avg = val, avg

    def __format__(self, spec):
        return '{0:{2}} ({1:{2}})'.format(float(self.val), float(self.avg), spec)


_PINNED_RING, _PINNED_NEXT = [], [0]


def _pinned_slot(n, ring=8, width=256):
    """A view of one of a few long-lived pinned host buffers (allocating pinned memory per step is slow and can
    synchronise); a slot is reused ``ring`` snapshots later, long after its copy has been consumed."""
    import torch
    if n > width:
        return torch.empty(n, dtype=torch.float32, pin_memory=True)
    if not _PINNED_RING:
        _PINNED_RING.extend(torch.empty(width, dtype=torch.float32, pin_memory=True) for _ in range(ring))
    buf = _PINNED_RING[_PINNED_NEXT[0] % ring]
    _PINNED_NEXT[0] += 1
    return buf[:n]


class MeterSnapshot:
    """The current val / avg of every meter; device scalars are gathered into one pinned host buffer with an
    asynchronous copy (one small launch + one D2H), awaited only when the snapshot is first read."""

    def __init__(self, meters):
        import torch
        self._host, self._slots, self._event, self._buf = {}, [], None, None
        dev = []
        for k, m in meters.items():
            for field in ('val', 'avg'):
                v = getattr(m, field)
                if torch.is_tensor(v) and v.is_cuda:
                    self._slots.append((k, field))
                    dev.append(v.detach().reshape(-1)[:1].float())
                else:
                    self._host[(k, field)] = float(v)
        if dev:
            self._buf = _pinned_slot(len(dev))
            self._buf.copy_(torch.cat(dev), non_blocking=True)
            self._event = torch.cuda.Event()
            self._event.record()

    def _resolve(self):
        if self._event is not None:
            self._event.synchronize()
            for (k, field), v in zip(self._slots, self._buf.tolist()):
                self._host[(k, field)] = v
            self._event = None

    def __getitem__(self, key):
        self._resolve()
        return _HostMeter(self._host[(key, 'val')], self._host[(key, 'avg')])


class AvgMeterSet:
    def __init__(self):
        self.meters = {}

    def snapshot(self):
        return MeterSnapshot(self.meters)

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
