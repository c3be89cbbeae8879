"""A few launches of one convolution shape for `ncu --set full -k regex:<kernel> -s <skip> -c 1`.
    python tools/prof_conv.py f16x3 fwd|fwdstats|wgrad <shape index of tools/bench_conv.SHAPES>"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pixelssl_b200 import ops
from tools.bench_conv import SHAPES, taps_of, CL

prec = ops.PRECISION[sys.argv[1]]
what = sys.argv[2]
name, N, H, W, Cin, Cout, k, dil, _ = SHAPES[int(sys.argv[3])]
taps = taps_of(k, dil)
nt = k * k
x = torch.randn(N, Cin, H, W, device='cuda').contiguous(memory_format=CL)
w = torch.randn(Cout * nt * Cin, device='cuda') * 0.05
out = torch.empty(N, Cout, H, W, device='cuda').contiguous(memory_format=CL)
dy = torch.randn(N, Cout, H, W, device='cuda').contiguous(memory_format=CL) * 1e-6
dw = torch.zeros(Cout * nt * Cin, device='cuda')
if prec >= 3:
    x = ops.h16_split(x, ops.H16_ACT_SCALE, prec == 3)
    w = ops.h16_split(w, ops.H16_W_SCALE, prec == 3)
    dy = ops.h16_split(dy, None, prec == 3)
for _ in range(5):
    if what in ('fwd', 'fwdstats'):
        st = torch.zeros(2 * Cout, dtype=torch.float64, device='cuda') if what == 'fwdstats' else None
        ops.conv_raw(x, w, None, taps, N, H, W, Cin, H, W, Cout, Cout, 1, 1, out=out, precision=prec, bn_stats=st)
    else:
        ops.conv_wgrad_raw(x, dy, dw, taps, N, H, W, Cin, H, W, Cout, Cout, 1, 1, precision=prec)
torch.cuda.synchronize()
print(name, 'done, tc status', ops.conv_tc_status())
