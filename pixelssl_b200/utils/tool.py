from . import logger


def dict_value(dictionary, name, default=None, err=False):
    """pixelssl/utils/tool.py:4-17."""
    if dictionary is None:
        if err:
            logger.log_err('The given dictionary is None\n')
        return default
    if name in dictionary:
        return dictionary[name]
    if err:
        logger.log_err('Cannot find key: {0}'.format(name))
    return default
