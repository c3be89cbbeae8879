"""Small helpers shared by the algorithm modules."""
from . import logger


def dict_value(dictionary, name, default=None, err=False):
    """Look ``name`` up in ``dictionary``.  A missing key (or no dictionary at all) yields ``default`` - or, when
    ``err`` is set, the reference's error convention: banner + exit (pixelssl/utils/tool.py:4-17)."""
    problem = None
    if dictionary is None:
        problem = 'The given dictionary is None\n'
    elif name not in dictionary:
        problem = 'Cannot find key: {0}'.format(name)
    else:
        return dictionary[name]
    if err:
        logger.log_err(problem)
    return default
