"""The algorithm contract of pixelssl/ssl_algorithm/ssl_base.py:19-159: module-level
``add_parser_arguments(parser)``, an export function named like the module, and an object with
``build / train / validate / save_checkpoint / load_checkpoint`` plus ``NAME`` and
``SUPPORTED_TASK_TYPES``.  TaskProxy (task_template/proxy.py:134-159,433-441) only ever touches
these."""
from ..utils import logger


def add_parser_arguments(parser):
    pass


class _SSLBase:
    NAME = 'ssl_base'
    SUPPORTED_TASK_TYPES = []

    def __init__(self, args):
        self.args = args
        self.task_func = None
        self.meters = logger.AvgMeterSet()
        self.models, self.optimizers, self.lrers, self.criterions = {}, {}, {}, {}

    def build(self, model_funcs, optimizer_funcs, lrer_funcs, criterion_funcs, task_func):
        self._build(model_funcs, optimizer_funcs, lrer_funcs, criterion_funcs, task_func)

    def train(self, data_loader, epoch):
        self._train(data_loader, epoch)

    def validate(self, data_loader, epoch):
        self._validate(data_loader, epoch)

    def save_checkpoint(self, epoch):
        self._save_checkpoint(epoch)

    def load_checkpoint(self):
        return self._load_checkpoint()

    def _build(self, model_funcs, optimizer_funcs, lrer_funcs, criterion_funcs, task_func):
        raise NotImplementedError

    def _train(self, data_loader, epoch):
        raise NotImplementedError

    def _validate(self, data_loader, epoch):
        raise NotImplementedError

    def _save_checkpoint(self, epoch):
        raise NotImplementedError

    def _load_checkpoint(self):
        raise NotImplementedError


def to_device(tensors, non_blocking=True):
    """``Variable(i).cuda()`` of every ``_batch_prehandle`` (ssl_mt.py:337-357): host -> HBM copy on
    the current stream (asynchronous when the loader pinned the batch)."""
    return tuple(t.cuda(non_blocking=non_blocking) for t in tensors)


def check_single_model_dicts(name, model_dict, optimizer_dict, lrer_dict, criterion_dict):
    if not len(model_dict) == len(optimizer_dict) == len(lrer_dict) == len(criterion_dict) == 1:
        logger.log_err('The len(element_dict) of {0} should be 1\n'.format(name.upper()))
    elif list(model_dict.keys())[0] != 'model':
        logger.log_err('In {0}, the key of element_dict should be \'model\',\n'
                       'but \'{1}\' is given\n'.format(name.upper(), model_dict.keys()))
