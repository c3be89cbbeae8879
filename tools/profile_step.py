"""One MT step of the bench workload between cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ..."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pixelssl_b200 import runner, ops

prec = sys.argv[1] if len(sys.argv) > 1 else 'tf32'
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ops.set_conv_precision(prec)
cfg = bench.mt_config()
cfg['batch_size'], cfg['unlabeled_batch_size'] = bs, bs // 2
a = runner.build_args(cfg, iters_per_epoch=662)
import logging
logging.getLogger('PixelSSL').setLevel(logging.ERROR)
alg = runner.build_algorithm(a)
alg.s_model.train(); alg.t_model.train()
bench.LBS = bench.UBS = bs // 2
img, lab = bench.synthetic_host_batches(1, 0, pin=False)[0]
img, lab = img.cuda(), lab.cuda()
for i in range(3):
    alg.train_step((img,), (lab,), i, 1986)
torch.cuda.synchronize()
torch.cuda.profiler.start()
alg.train_step((img,), (lab,), 3, 1986)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('profiled one step, launches (ours):', ops.launch_count())
