"""Host-side cost of one step: enqueue time (no synchronisation) against device time, and a cProfile of one step.
    python tools/host_profile.py [precision] [config]"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cProfile
import pstats
import random
import numpy as np
import torch
import bench
from pixelssl_b200 import runner, ops
import logging

prec = sys.argv[1] if len(sys.argv) > 1 else 'f16x3'
config = sys.argv[2] if len(sys.argv) > 2 else 'mt'
ops.set_conv_precision(prec)
make_cfg, lbs, ubs, size, _ = bench.CONFIGS[config]
torch.manual_seed(0); random.seed(0); np.random.seed(0)
logging.getLogger('PixelSSL').setLevel(logging.ERROR)
alg = runner.build_algorithm(runner.build_args(make_cfg(), iters_per_epoch=662))
img, lab = bench.synthetic_host_batches(1, 0, False, lbs, ubs, size)[0]
batch = [((img.cuda(),), (lab.cuda(),))]
for i in range(4):
    alg._train(batch, i)
torch.cuda.synchronize()
import gc
host, dev = [], []
for i in range(10):
    if i == 6:
        gc.disable()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    alg._train(batch, 4 + i)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    host.append((t1 - t0) * 1e3); dev.append((t2 - t0) * 1e3)
gc.enable()
print('(gc disabled for the last 4)')
print('host enqueue ms/step:', ['%.1f' % h for h in host], ' total ms/step:', ['%.1f' % d for d in dev])
pr = cProfile.Profile()
pr.enable()
alg._train(batch, 20)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(28)
