"""Device timing of pxl_bilinear_bwd on the DeepLab-v2 head shape (33x33x21(32) <- 513x513, 16 images)."""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pixelssl_b200._lib import call

CL = torch.channels_last
go = torch.randn(16, 21, 513, 513, device='cuda')
gin = torch.zeros(16, 32, 33, 33, device='cuda').contiguous(memory_format=CL)
flush = torch.empty(256 * 1024 * 1024 // 4, device='cuda')
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: ctypes.c_void_p(t.data_ptr())
ts = []
for i in range(15):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    call('pxl_bilinear_bwd', P(go), P(gin), 16, 21, 33, 33, 513, 513, 1, 1, 32, st)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts = sorted(ts[5:])
nbytes = go.numel() * 4
print('bilinear_bwd 16x21x513x513 -> 33x33: median %.3f ms, %.0f GB/s' % (ts[len(ts) // 2], nbytes / ts[len(ts) // 2] / 1e6))
