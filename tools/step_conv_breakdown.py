"""Per-shape device time of the convolution launches inside one MT step (CUDA events around every launch):
    python tools/step_conv_breakdown.py [f16x3]"""
import os
import sys
import collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pixelssl_b200 import runner, ops
import logging
logging.getLogger('PixelSSL').setLevel(logging.ERROR)
prec = sys.argv[1] if len(sys.argv) > 1 else 'f16x3'
ops.set_conv_precision(prec)
alg = runner.build_algorithm(runner.build_args(bench.mt_config(), iters_per_epoch=662))
img, lab = bench.synthetic_host_batches(1, 0, False)[0]
img, lab = img.cuda(), lab.cuda()
for i in range(3):
    alg.train_step((img,), (lab,), i, 1986)
names = ['pxl_conv_h16_launch', 'pxl_conv_wgrad_h16_launch', 'pxl_conv_tc_launch_ex', 'pxl_conv_wgrad_tc_launch']
for n in names:
    ops.kernel_timer_start(n)
alg.train_step((img,), (lab,), 3, 1986)
agg = collections.OrderedDict()
for n in names:
    for ms, meta in ops.kernel_timer_stop(n, with_meta=True):
        fl, desc = meta
        a = agg.setdefault(desc, [0, 0.0, 0.0])
        a[0] += 1; a[1] += ms; a[2] += fl
tot = sum(a[1] for a in agg.values())
print('precision %s: %d conv launches, %.2f ms' % (prec, sum(a[0] for a in agg.values()), tot))
for desc, (cnt, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%7.3f ms %5.1f%%  x%-3d %7.1f us/launch %7.1f TF/s  %s' % (ms, 100 * ms / tot, cnt, ms / cnt * 1e3, fl / ms / 1e9, desc))
