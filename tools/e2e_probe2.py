"""Per-step host and device timestamps of algorithm.train() on pinned batches (bimodal epochs: 43 vs 61 ms/step)."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pixelssl_b200 import runner, ops
import logging
logging.getLogger('PixelSSL').setLevel(logging.ERROR)

ops.set_conv_precision('f16x3')
host = bench.synthetic_host_batches(8, 0, True)
a = runner.build_args(bench.mt_config(), iters_per_epoch=662)
a.log_freq = 1
alg = runner.build_algorithm(a)
loader = [((b[0],), (b[1],)) for b in host]
rec = []
orig = alg.train_step


def wrapped(*args, **kw):
    e = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e.record()
    orig(*args, **kw)
    rec.append((t0, time.perf_counter(), e))


alg.train_step = wrapped
import gc
for ep in range(10):
    if ep == 5:
        gc.collect(); gc.disable(); print('gc disabled from here')
    g0 = [st['collections'] for st in gc.get_stats()]
    rec.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    alg.train(loader, ep)
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t0) * 1e3 / len(loader)
    starts = [(r[0] - t0) * 1e3 for r in rec]
    enq = [(r[1] - r[0]) * 1e3 for r in rec]
    gaps = [(rec[i + 1][0] - rec[i][1]) * 1e3 for i in range(len(rec) - 1)]
    dev = [rec[i].__getitem__(2).elapsed_time(rec[i + 1][2]) for i in range(len(rec) - 1)]
    g1 = [st['collections'] for st in gc.get_stats()]
    print('gc collections gen0/1/2 during the epoch:', [b - a_ for a_, b in zip(g0, g1)], 'first step starts at %.0f ms, last ends at %.0f of %.0f ms' % (starts[0], (rec[-1][1] - t0) * 1e3, tot * len(loader)))
    print('epoch %d: %.1f ms/step | host enqueue %s | host between steps %s | device step-to-step %s' % (
        ep, tot, ' '.join('%.0f' % v for v in enq), ' '.join('%.0f' % v for v in gaps), ' '.join('%.0f' % v for v in dev)))
