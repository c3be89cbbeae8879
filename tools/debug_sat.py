"""Which fp16-pair producer saturates on the bench workload?  python tools/debug_sat.py [size] [bs]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pixelssl_b200 import runner, ops

size = int(sys.argv[1]) if len(sys.argv) > 1 else 513
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ops.set_conv_precision('f16x3')
cfg = bench.mt_config()
cfg['batch_size'], cfg['unlabeled_batch_size'] = bs, bs // 2
a = runner.build_args(cfg, iters_per_epoch=662)
import logging
logging.getLogger('PixelSSL').setLevel(logging.ERROR)
alg = runner.build_algorithm(a)
alg.s_model.train(); alg.t_model.train()
img, lab = bench.synthetic_host_batches(1, 0, False, bs // 2, bs // 2, size)[0]
img, lab = img.cuda(), lab.cuda()

# wrap the BN backward launch to find the layers whose dx pair clips
orig_call = ops.call
state = {'n': 0}


def spy(name, *args):
    rc = orig_call(name, *args)
    if name == 'pxl_bn_bwd_dx_h16':
        before = state.get('last', (0, 0, 0, 0))
        now = ops.h16_status_sites()
        if now[3] != before[3]:
            rows, C = args[11], args[12]
            slot = args[19]
            sl = torch.empty(4, device='cuda')
            import ctypes
            torch.cuda.synchronize()
            buf = (ctypes.c_float * 4)()
            ctypes.cdll.LoadLibrary('libcudart.so').cudaMemcpy(buf, slot, 16, 2)
            amax = ctypes.c_uint.from_buffer(ctypes.c_float(buf[2])).value
            import struct
            amaxf = struct.unpack('f', struct.pack('I', amax))[0]
            print('  dx clip: rows %d C %d  +%d threads  s=%g  absmax(dz)=%g relu=%d' % (rows, C, now[3] - before[3], buf[0], amaxf, args[8]))
        state['last'] = now
    return rc


ops.call = spy
for i in range(2):
    alg.train_step((img,), (lab,), i, 1986)
    torch.cuda.synchronize()
    print('step', i, 'sites (split fixed, split dyn, bn apply, bn dx):', ops.h16_status_sites(),
          's_task_loss', float(alg.meters['s_task_loss'].val))
