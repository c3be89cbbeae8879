// Bilinear resize forward / backward (F.interpolate(mode='bilinear'), both align_corners flavours)
//   deeplab_v2.py:32, _pspnet.py:99-100,127, ssl_gct.py:580, ssl_adv.py:488, ssl_cct.py:482
// Output (and grad_out) are planar [n, C, H, W]; the small input is planar or NHWC (ldc stride).
// Forward is write-bound: 4*C B per output pixel.  The source-index arithmetic mirrors ATen's
// area_pixel_compute_source_index so results agree to fp32 round-off.
#include "common.cuh"

__device__ __forceinline__ float src_index(float scale, int dst, bool align_corners) {
    if (align_corners) return scale * (float)dst;
    float s = scale * ((float)dst + 0.5f) - 0.5f;
    return s < 0.f ? 0.f : s;
}

static inline float resize_scale(int in, int out, int align_corners) {
    if (align_corners) return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
    return (float)in / (float)out;
}

#define BL_MAXC 32
template <bool NHWC>
__global__ void __launch_bounds__(256)
bilinear_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int h, int w, int H, int W,
                    float sh, float sw, bool ac, int ldc) {
    const int b = blockIdx.z;
    const int y = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const float fy = src_index(sh, y, ac), fx = src_index(sw, x, ac);
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly1 = fy - (float)y0, lx1 = fx - (float)x0;
    const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    const int64_t HW = (int64_t)H * W;
    float* op = out + (int64_t)b * C * HW + (int64_t)y * W + x;
    if (NHWC) {
        const float* p00 = in + ((int64_t)(b * h + y0) * w + x0) * ldc;
        const float* p01 = in + ((int64_t)(b * h + y0) * w + x1) * ldc;
        const float* p10 = in + ((int64_t)(b * h + y1) * w + x0) * ldc;
        const float* p11 = in + ((int64_t)(b * h + y1) * w + x1) * ldc;
        for (int c = 0; c < C; ++c)
            op[(int64_t)c * HW] = ly0 * (lx0 * __ldg(p00 + c) + lx1 * __ldg(p01 + c)) +
                                  ly1 * (lx0 * __ldg(p10 + c) + lx1 * __ldg(p11 + c));
    } else {
        const int64_t hw = (int64_t)h * w;
        const float* base = in + (int64_t)b * C * hw;
        for (int c = 0; c < C; ++c) {
            const float* pc = base + (int64_t)c * hw;
            op[(int64_t)c * HW] = ly0 * (lx0 * __ldg(pc + y0 * w + x0) + lx1 * __ldg(pc + y0 * w + x1)) +
                                  ly1 * (lx0 * __ldg(pc + y1 * w + x0) + lx1 * __ldg(pc + y1 * w + x1));
        }
    }
}

extern "C" int pxl_bilinear_fwd(const float* in, float* out, int n, int C, int h, int w, int H, int W,
                                int align_corners, int in_nhwc, int ldc, void* stream) {
    if (!in || !out || n <= 0 || C <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return PXL_ERR_BAD_ARG;
    if (H > 65535 || n > 65535) return PXL_ERR_UNSUPPORTED;
    const float sh = resize_scale(h, H, align_corners), sw = resize_scale(w, W, align_corners);
    dim3 grid((unsigned)pxl_cdiv(W, 256), (unsigned)H, (unsigned)n);
    cudaStream_t st = (cudaStream_t)stream;
    if (in_nhwc) bilinear_fwd_kernel<true><<<grid, 256, 0, st>>>(in, out, C, h, w, H, W, sh, sw, align_corners != 0, ldc);
    else bilinear_fwd_kernel<false><<<grid, 256, 0, st>>>(in, out, C, h, w, H, W, sh, sw, align_corners != 0, C);
    PXL_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------
// backward: grad_in[b,c,i,j] = sum over the output pixels whose 2x2 support includes (i,j).
// One CTA per (low-res row i, channel c, sample b).  Its warps stream the contributing
// high-res rows (coalesced along x), weight by the y-coefficient and scatter along x into a
// shared-memory row accumulator with shared atomics; each high-res row is read by the (at
// most) two low-res rows it touches.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void y_support(int i, int h, int H, float sh, bool ac, int& ylo, int& yhi) {
    // conservative bounds on {y : y0(y) == i or y1(y) == i}; refined per row in the loop
    if (sh <= 0.f) { ylo = 0; yhi = H - 1; return; }
    float lo, hi;
    if (ac) { lo = ((float)i - 1.f) / sh; hi = ((float)i + 1.f) / sh; }
    else { lo = ((float)i - 0.5f) / sh - 0.5f; hi = ((float)i + 1.5f) / sh - 0.5f; }
    ylo = (int)floorf(lo) - 1; yhi = (int)ceilf(hi) + 1;
    if (i == 0) ylo = 0;           // align_corners=False clamps negative sources to row 0
    if (ylo < 0) ylo = 0;
    if (yhi > H - 1) yhi = H - 1;
}

template <bool NHWC>
__global__ void __launch_bounds__(128)
bilinear_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin, int C, int h, int w, int H, int W,
                    float sh, float sw, bool ac, int ldc) {
    extern __shared__ float acc[];   // [w]
    const int i = blockIdx.x, c = blockIdx.y, b = blockIdx.z;
    for (int j = threadIdx.x; j < w; j += blockDim.x) acc[j] = 0.f;
    __syncthreads();
    int ylo, yhi;
    y_support(i, h, H, sh, ac, ylo, yhi);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const float* gp = gout + ((int64_t)b * C + c) * (int64_t)H * W;
    for (int y = ylo + warp; y <= yhi; y += nwarps) {
        const float fy = src_index(sh, y, ac);
        const int y0 = (int)fy;
        const int y1 = y0 + (y0 < h - 1 ? 1 : 0);
        const float ly1 = fy - (float)y0;
        float wy = 0.f;
        if (y0 == i) wy += 1.f - ly1;
        if (y1 == i) wy += ly1;
        if (wy == 0.f && y0 != i && y1 != i) continue;
        const float* row = gp + (int64_t)y * W;
        for (int x = lane; x < W; x += 32) {
            const float g = __ldg(row + x) * wy;
            const float fx = src_index(sw, x, ac);
            const int x0 = (int)fx;
            const int x1 = x0 + (x0 < w - 1 ? 1 : 0);
            const float lx1 = fx - (float)x0;
            atomicAdd(acc + x0, g * (1.f - lx1));
            atomicAdd(acc + x1, g * lx1);
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < w; j += blockDim.x) {
        if (NHWC) gin[((int64_t)(b * h + i) * w + j) * ldc + c] = acc[j];
        else gin[(((int64_t)b * C + c) * h + i) * w + j] = acc[j];
    }
}

extern "C" int pxl_bilinear_bwd(const float* grad_out, float* grad_in, int n, int C, int h, int w, int H, int W,
                                int align_corners, int in_nhwc, int ldc, void* stream) {
    if (!grad_out || !grad_in || n <= 0 || C <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return PXL_ERR_BAD_ARG;
    if (C > 65535 || n > 65535 || (size_t)w * sizeof(float) > 48 * 1024) return PXL_ERR_UNSUPPORTED;
    const float sh = resize_scale(h, H, align_corners), sw = resize_scale(w, W, align_corners);
    dim3 grid((unsigned)h, (unsigned)C, (unsigned)n);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)w * sizeof(float);
    if (in_nhwc) bilinear_bwd_kernel<true><<<grid, 128, smem, st>>>(grad_out, grad_in, C, h, w, H, W, sh, sw, align_corners != 0, ldc);
    else bilinear_bwd_kernel<false><<<grid, 128, smem, st>>>(grad_out, grad_in, C, h, w, H, W, sh, sw, align_corners != 0, C);
    PXL_CHECK_LAUNCH();
    return 0;
}
