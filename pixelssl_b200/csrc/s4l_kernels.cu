// S4L (pixelssl/ssl_algorithm/ssl_s4l.py): the rotated copies of a batch.
// SSLS4L._batch_prehandle (ssl_s4l.py:296-350) doubles every input / ground-truth tensor [bs,C,H,W]: rows 0..bs-1 keep
// the samples, row bs+i holds sample i rotated by angle_i in {1,2,3} quarter turns (_rotate_tensor, :352-360):
//   1: transpose(1,2).flip(2)   out[c,i,j] = in[c, W-1-j, i]
//   2: flip(2).flip(1)          out[c,i,j] = in[c, H-1-i, W-1-j]
//   3: transpose(1,2).flip(1)   out[c,i,j] = in[c, j, H-1-i]
// (angles 1 and 3 need H == W, like the reference's in-place assignment).  One launch for the whole batch instead of
// 2*bs slice assignments; pure data movement (HBM-bound, 8 B/element).
#include "common.cuh"

__global__ void __launch_bounds__(256)
s4l_rotate_batch_kernel(const float* __restrict__ in, float* __restrict__ out, const int* __restrict__ angles,
                        int bs, int C, int H, int W) {
    const int64_t plane = (int64_t)H * W, per = (int64_t)C * plane, total = 2 * (int64_t)bs * per;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / per);
        if (n < bs) { out[i] = __ldg(in + i); continue; }
        const int s = n - bs;
        const int64_t r = i - (int64_t)n * per;
        const int c = (int)(r / plane);
        const int y = (int)((r - (int64_t)c * plane) / W), x = (int)(r % W);
        const int a = __ldg(angles + s);
        int sy = y, sx = x;
        if (a == 1) { sy = W - 1 - x; sx = y; }
        else if (a == 2) { sy = H - 1 - y; sx = W - 1 - x; }
        else if (a == 3) { sy = x; sx = H - 1 - y; }
        out[i] = __ldg(in + ((int64_t)s * C + c) * plane + (int64_t)sy * W + sx);
    }
}

// out [2*bs,C,H,W] from in [bs,C,H,W]; angles: DEVICE int32 [bs] with values 0..3 (0 = plain copy)
extern "C" int pxl_s4l_rotate_batch(const float* in, float* out, const int* angles_dev, int bs, int C, int H, int W,
                                    int any_quarter_turn, void* stream) {
    if (!in || !out || !angles_dev || bs <= 0 || C <= 0 || H <= 0 || W <= 0) return PXL_ERR_BAD_ARG;
    if (any_quarter_turn && H != W) return PXL_ERR_BAD_ARG;       // 90 / 270 degrees need square maps
    const int64_t total = 2 * (int64_t)bs * C * H * W;
    int blocks = (int)(pxl_cdiv(total, 256 * 4) < PXL_NUM_SMS * 16 ? pxl_cdiv(total, 256 * 4) : PXL_NUM_SMS * 16);
    s4l_rotate_batch_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(in, out, angles_dev, bs, C, H, W);
    PXL_CHECK_LAUNCH();
    return 0;
}
