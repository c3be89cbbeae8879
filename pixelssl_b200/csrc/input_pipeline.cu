// Training / validation input pipeline of the sseg task on the GPU (task/sseg/data.py:90-123, 142-292):
//   8-bit HWC image (+ 8-bit label map)  ->  [resize short edge]  ->  [zero pad]  ->  crop  ->  [h-flip]
//   ->  (x/255 - mean)/std in the reference's float order  ->  float32 CHW image, float32 label map
// in ONE launch per sample, bit for bit what PIL + numpy produce in the reference:
//   * Image.resize(BILINEAR) on 8-bit images is Pillow's separable antialiased resampling in 22-bit fixed point with an
//     8-bit intermediate image (libImaging/Resample.c): horizontal pass first, each pass
//     clip8((2^21 + sum src*k) >> 22).  The per-output-sample windows and fixed-point weights are computed on the host
//     (double arithmetic, exactly Pillow's precompute_coeffs / normalize_coeffs_8bpc) and passed as tables.  The kernel
//     evaluates only the output crop: for an output pixel it recomputes the horizontal pass for the few source rows
//     its vertical window touches (no intermediate image in HBM).
//   * Image.resize(NEAREST) for the label is a gather through host-computed index tables (ImagingScaleAffine's
//     accumulated source coordinate).
//   * Normalize (data.py:153-161): float32 x/255, then float64 (x - mean) and (x / std) each rounded to float32, which
//     is what numpy's in-place ops on a float32 array with float64 constants do.
// HBM-bound and tiny (a 513x513 crop is 3.2 MB out); one thread per output pixel, coalesced CHW stores.
#include "common.cuh"

struct PrehandleParams {
    int H, W;            // source image
    int ow, oh;          // resized size (== W, H when no_resize)
    int no_resize;
    int crop_w, crop_h;  // output size
    int x1, y1;          // crop origin in the resized (and padded) image
    int flip;            // horizontal flip of the crop
    int kmax_x, kmax_y;  // row strides of the weight tables
    float label_fill;    // label value in the padded area
    float label_const;   // used when lab == NULL (unlabeled sample: the reference returns image[0]*0 - 1)
    double mean[3], stdv[3];
};

__device__ __forceinline__ int clip8(long long v) { return v < 0 ? 0 : (v > 255 ? 255 : (int)v); }

__global__ void __launch_bounds__(256)
prehandle_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ lab,
                 const int* __restrict__ xb, const int* __restrict__ xk,      // [ow][2] (first, count), [ow][kmax_x]
                 const int* __restrict__ yb, const int* __restrict__ yk,      // [oh][2], [oh][kmax_y]
                 const int* __restrict__ lx, const int* __restrict__ ly,      // label gather tables [ow], [oh]
                 PrehandleParams p, float* __restrict__ out_img, float* __restrict__ out_lab) {
    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy = blockIdx.y;
    if (ox >= p.crop_w) return;
    const int rx = p.x1 + (p.flip ? (p.crop_w - 1 - ox) : ox);
    const int ry = p.y1 + oy;
    const bool inside = rx < p.ow && ry < p.oh;          // else: the zero padding ImageOps.expand added right / below
    int v[3] = {0, 0, 0};
    float lv = p.label_fill;
    if (inside) {
        if (p.no_resize) {
            const uint8_t* s = img + ((int64_t)ry * p.W + rx) * 3;
            v[0] = s[0]; v[1] = s[1]; v[2] = s[2];
        } else {
            const int x0 = __ldg(xb + 2 * rx), xn = __ldg(xb + 2 * rx + 1);
            const int y0 = __ldg(yb + 2 * ry), yn = __ldg(yb + 2 * ry + 1);
            const int* kx = xk + (int64_t)rx * p.kmax_x;
            const int* ky = yk + (int64_t)ry * p.kmax_y;
            long long acc[3] = {1ll << 21, 1ll << 21, 1ll << 21};
            for (int j = 0; j < yn; ++j) {
                // horizontal pass of source row y0 + j at output column rx (8-bit intermediate, like Pillow)
                const uint8_t* row = img + ((int64_t)(y0 + j) * p.W + x0) * 3;
                long long h[3] = {1ll << 21, 1ll << 21, 1ll << 21};
                if (p.ow == p.W) {                        // Pillow skips a pass whose size does not change
                    h[0] = (long long)img[((int64_t)(y0 + j) * p.W + rx) * 3 + 0] << 22;
                    h[1] = (long long)img[((int64_t)(y0 + j) * p.W + rx) * 3 + 1] << 22;
                    h[2] = (long long)img[((int64_t)(y0 + j) * p.W + rx) * 3 + 2] << 22;
                } else {
                    for (int i = 0; i < xn; ++i) {
                        const long long k = __ldg(kx + i);
                        h[0] += row[3 * i + 0] * k; h[1] += row[3 * i + 1] * k; h[2] += row[3 * i + 2] * k;
                    }
                }
                const long long w = __ldg(ky + j);
                acc[0] += clip8(h[0] >> 22) * w; acc[1] += clip8(h[1] >> 22) * w; acc[2] += clip8(h[2] >> 22) * w;
            }
            if (p.oh == p.H) {
                // no vertical pass: the (single-row) window is the identity
                const uint8_t* row = img + ((int64_t)ry * p.W + x0) * 3;
                long long h[3] = {1ll << 21, 1ll << 21, 1ll << 21};
                for (int i = 0; i < xn; ++i) {
                    const long long k = __ldg(kx + i);
                    h[0] += row[3 * i + 0] * k; h[1] += row[3 * i + 1] * k; h[2] += row[3 * i + 2] * k;
                }
                v[0] = clip8(h[0] >> 22); v[1] = clip8(h[1] >> 22); v[2] = clip8(h[2] >> 22);
            } else {
                v[0] = clip8(acc[0] >> 22); v[1] = clip8(acc[1] >> 22); v[2] = clip8(acc[2] >> 22);
            }
        }
        if (lab) lv = p.no_resize ? (float)lab[(int64_t)ry * p.W + rx] : (float)lab[(int64_t)__ldg(ly + ry) * p.W + __ldg(lx + rx)];
    }
    if (!lab) lv = p.label_const;
    const int64_t plane = (int64_t)p.crop_w * p.crop_h;
    const int64_t o = (int64_t)oy * p.crop_w + ox;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float a = __fdiv_rn((float)v[c], 255.0f);
        const float b = __double2float_rn((double)a - p.mean[c]);
        out_img[c * plane + o] = __double2float_rn((double)b / p.stdv[c]);
    }
    out_lab[o] = lv;
}

// Tables are DEVICE int32 arrays built by the host side (pixelssl_b200/task/sseg/gpu_input.py):
//   xb [ow][2] / yb [oh][2]: first source index and tap count of every resized column / row; xk [ow][kmax_x] /
//   yk [oh][kmax_y]: 22-bit fixed-point weights; lx [ow] / ly [oh]: NEAREST source indices for the label.
// no_resize != 0 (validation without rescaling): tables may be NULL, ow == W and oh == H.
// lab NULL: unlabeled sample, the label output is the constant label_const (data.py:105).
extern "C" int pxl_input_prehandle(const uint8_t* img_hwc, const uint8_t* lab_hw, int H, int W, int ow, int oh, int no_resize,
                                   const int* xb, const int* xk, int kmax_x, const int* yb, const int* yk, int kmax_y,
                                   const int* lx, const int* ly, int x1, int y1, int crop_w, int crop_h, int flip,
                                   float label_fill, float label_const, const double* mean3_host, const double* std3_host,
                                   float* out_img_chw, float* out_lab_hw, void* stream) {
    if (!img_hwc || !out_img_chw || !out_lab_hw || !mean3_host || !std3_host) return PXL_ERR_BAD_ARG;
    if (H <= 0 || W <= 0 || ow <= 0 || oh <= 0 || crop_w <= 0 || crop_h <= 0 || x1 < 0 || y1 < 0) return PXL_ERR_BAD_ARG;
    if (!no_resize && (!xb || !xk || !yb || !yk || kmax_x <= 0 || kmax_y <= 0)) return PXL_ERR_BAD_ARG;
    if (no_resize && (ow != W || oh != H)) return PXL_ERR_BAD_ARG;
    if (lab_hw && !no_resize && (!lx || !ly)) return PXL_ERR_BAD_ARG;
    PrehandleParams p;
    p.H = H; p.W = W; p.ow = ow; p.oh = oh; p.no_resize = no_resize; p.crop_w = crop_w; p.crop_h = crop_h;
    p.x1 = x1; p.y1 = y1; p.flip = flip; p.kmax_x = kmax_x; p.kmax_y = kmax_y;
    p.label_fill = label_fill; p.label_const = label_const;
    for (int c = 0; c < 3; ++c) { p.mean[c] = mean3_host[c]; p.stdv[c] = std3_host[c]; }
    dim3 grid((unsigned)pxl_cdiv(crop_w, 256), (unsigned)crop_h);
    prehandle_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img_hwc, lab_hw, xb, xk, yb, yk, lx, ly, p, out_img_chw, out_lab_hw);
    PXL_CHECK_LAUNCH();
    return 0;
}
