"""torchrun --nproc-per-node 2 tools/debug_sat2.py : find the fp16-pair split call that clips under data parallelism."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
rank, local = int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if int(os.environ.get('WORLD_SIZE', 1)) > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
import bench
from pixelssl_b200 import runner, ops
import logging
logging.getLogger('PixelSSL').setLevel(logging.ERROR)
ops.set_conv_precision(os.environ.get('PXL_DEBUG_PREC', 'f16x3'))
a = runner.build_args(bench.mt_config(), iters_per_epoch=662)
alg = runner.build_algorithm(a)
host = bench.synthetic_host_batches(4, rank, True)
orig = ops.call
last = [ops.h16_status_sites()]


def spy(name, *args):
    rc = orig(name, *args)
    if name in ('pxl_h16_split', 'pxl_bn_apply_h16', 'pxl_bn_finalize_apply_h16', 'pxl_bn_bwd_dx_h16'):
        now = ops.h16_status_sites()
        if now != last[0] and rank == 0:
            if name == 'pxl_h16_split':
                n, scale = args[3], args[4]
                src = torch.empty(0)
                print('  CLIP in %s: n=%d scale=%g dyn=%s  +%s' % (name, n, scale, bool(args[5].value), [b - c for b, c in zip(now, last[0])]), flush=True)
            else:
                print('  CLIP in %s +%s' % (name, [b - c for b, c in zip(now, last[0])]), flush=True)
        last[0] = now
    return rc


ops.call = spy
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
    img, lab = host[i % 4]
    alg.train_step((img,), (lab,), i, 1986)
    alg.s_lrer.step()
    torch.cuda.synchronize()
    if rank == 0:
        d = alg.s_model.arena.data
        print('step %d loss %.4f cons %.5f  |param|max %.3g  grad max %.3g sites %s' % (
            i, float(alg.meters['s_task_loss'].val), float(alg.meters['cons_loss'].val), float(d.abs().max()),
            float(alg.s_model.arena.grad.abs().max()), ops.h16_status_sites()), flush=True)
if dist.is_initialized():
    dist.barrier()
    dist.destroy_process_group()
