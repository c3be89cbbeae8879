"""One step of a bench workload between cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ...
   python tools/profile_step.py [precision] [config: mt|cutmix|gct|cct]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import random
import numpy as np
import torch
import bench
from pixelssl_b200 import runner, ops

prec = sys.argv[1] if len(sys.argv) > 1 else 'f16x3'
config = sys.argv[2] if len(sys.argv) > 2 else 'mt'
ops.set_conv_precision(prec)
make_cfg, lbs, ubs, size, _ = bench.CONFIGS[config]
torch.manual_seed(0); random.seed(0); np.random.seed(0)
a = runner.build_args(make_cfg(), iters_per_epoch=662)
import logging
logging.getLogger('PixelSSL').setLevel(logging.ERROR)
alg = runner.build_algorithm(a)
img, lab = bench.synthetic_host_batches(1, 0, False, lbs, ubs, size)[0]
batch = [((img.cuda(),), (lab.cuda(),))]
for i in range(3):
    alg._train(batch, i)
torch.cuda.synchronize()
torch.cuda.profiler.start()
alg._train(batch, 3)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('profiled one step, launches (ours):', ops.launch_count())
