"""Why is the end-to-end arm slower in the first process on a fresh box?  Times (a) the first and second H2D copy of
every pinned batch, (b) three epochs of algorithm.train() over the same pinned batches.
    python tools/e2e_probe.py"""
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
t0 = time.perf_counter()
host = bench.synthetic_host_batches(8, 0, True)
print('8 pinned batches created in %.2f s' % (time.perf_counter() - t0))
dst = [torch.empty_like(t, device='cuda') for t in host[0]]
for rnd in range(2):
    ts = []
    for img, lab in host:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dst[0].copy_(img, non_blocking=True); dst[1].copy_(lab, non_blocking=True)
        e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    print('H2D pass %d (67 MB per batch) ms:' % rnd, ['%.2f' % t for t in ts])
a = runner.build_args(bench.mt_config(), iters_per_epoch=662)
a.log_freq = 1
alg = runner.build_algorithm(a)
loader = [((b[0],), (b[1],)) for b in host]
for ep in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    alg.train(loader, ep)
    torch.cuda.synchronize()
    print('epoch %d: %.2f ms/step' % (ep, (time.perf_counter() - t0) * 1e3 / len(loader)))
alg.args.log_freq = 10 ** 9
dev = [((b[0].cuda(),), (b[1].cuda(),)) for b in host]
for ep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    alg._train(dev, 4 + ep)
    torch.cuda.synchronize()
    print('device-resident epoch %d: %.2f ms/step' % (ep, (time.perf_counter() - t0) * 1e3 / len(dev)))
