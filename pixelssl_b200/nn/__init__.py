from . import func, optimizer, lrer
from .modules import Conv2d, BatchNorm2d, SynchronizedBatchNorm2d
from .arena import ParamArena, EngineParallel
