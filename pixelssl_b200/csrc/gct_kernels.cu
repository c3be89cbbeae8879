// GCT / AdvSSL tail kernels (separable reflect-pad Gaussian blur, 3x3 dilate, min-max normalise,
// masked BCE-with-logits).  Filled in as those rows of SURVEY.md section 8 are built.
#include "common.cuh"
