from .logger import log_info, log_warn, log_err, AvgMeter, AvgMeterSet
from .tool import dict_value
from .cmd import str2bool, str2intlist, parse_args

REGRESSION = 'regression'            # pixelssl/utils/constant.py
CLASSIFICATION = 'classification'
