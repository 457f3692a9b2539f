// tcgen05 implicit-GEMM 3x3 convolution (placeholder until the kernel lands: reports "unsupported" so the
// direct kernel runs).
#include "mzb_resnet_model.h"

bool mzb_conv_tc_supported(const ConvParams&, int, int, int) { return false; }
int mzb_conv_tc_launch(int, int, int, const ConvParams&, const __nv_bfloat16*, const float*, const __nv_bfloat16*, int,
                       __nv_bfloat16*, cudaStream_t) {
  mzb_set_error("tensor-core convolution not available");
  return MZB_EUNSUPPORTED;
}
