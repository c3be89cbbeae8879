// tcgen05 (5th-gen tensor core) convolution kernels.  Placeholder entry points until the
// TMA + tcgen05 implicit-GEMM lands: they report PXL_ERR_UNSUPPORTED so callers can select the
// fp32 path explicitly; nothing is silently computed elsewhere.
#include "common.cuh"

extern "C" int pxl_conv_tc_impl(const pxl_conv_geom*, const int*, const float*, const float*, const float*, float*, void*) {
    return PXL_ERR_UNSUPPORTED;
}
extern "C" int pxl_conv_wgrad_tc_impl(const pxl_conv_geom*, const int*, const float*, const float*, float*, void*) {
    return PXL_ERR_UNSUPPORTED;
}
