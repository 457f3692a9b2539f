// Residual model handle + the tensor-core convolution hooks shared by mzb_resnet.cu / mzb_conv_tc.cu.
#pragma once
#include <algorithm>
#include <vector>

#include "mzb_resnet.cuh"

struct mzb_resnet_model {
  mzb_resnet_config cfg;
  int Cobs, H, W, A, S, full, blocks, C, downsample, precision;
  int Hl, Wl;                      // latent (hidden-state) height / width
  int n_tensors;
  ConvParams rep_conv, dyn_conv, ds_conv1, ds_conv2;
  std::vector<Block> rep_blocks, dyn_blocks, pred_blocks, ds1, ds2, ds3;
  HeadParams reward, value, policy;
  // bf16 DownSample stem on the tensor cores: resblocks1 with channels zero-padded C/2 -> stem_cp1 (multiple of 16)
  float* pv_w = nullptr;           // value and policy 1x1 weights concatenated [r_value + r_policy][C] (fused projection)
  int* t16_counters = nullptr;     // k_recurrent16: [0] next image, [1] warps that left; reset by the last warp of a launch
  int stem_tc = 0, stem_cp1 = 0;
  std::vector<Block> ds1_tc;
  // DownSample.conv2 (C/2 -> C, stride 2) as six 16 x 16 MMA taps on pixel-pair rows (mzb_stem16.cu): bf16 [6][C][16], or NULL
  __nv_bfloat16* ds_conv2_mma = nullptr;
  std::vector<void*> allocs;
};

// tcgen05 implicit-GEMM 3x3 convolution on NHWC bf16 (mzb_conv_tc.cu)
bool mzb_conv_tc_enabled();
bool mzb_conv_tc_supported(const ConvParams& cp, int H, int W, int cin_stride);
struct TcProj { const float* w; float* out; int r; };   // fused 1x1 head projection: w [r][C_out] fp32, out [B][r][H*W] fp32
int mzb_conv_tc_launch(int B, int H, int W, const ConvParams& cp, const __nv_bfloat16* x, const float* plane,
                       const __nv_bfloat16* residual, int relu, __nv_bfloat16* y, cudaStream_t stream, int zero_pads = 0,
                       const TcProj* proj = nullptr);

// head mlp on mma.sync (mzb_head_mma.cu); w_host[l] = fc weights [in][out] fp32 on the host
bool mzb_head_mma_pack(mzb_resnet_model* m, const HeadParams& hp, const std::vector<std::vector<float>>& w_host,
                       const std::vector<std::vector<float>>& b_host, void** opaque);
void mzb_head_mma_free(void* opaque);
int mzb_head_mma_launch(void* opaque, const float* proj, long long proj_stride, int proj_off, int B, int S, int mode,
                        const uint8_t* legal, float* logits, float* scalar, float* priors, cudaStream_t stream);
// several heads of one inference as ONE launch (grid.y = head); mode 0: support -> scalar, 1: legal-action softmax
struct MmaHeadCall {
  void* opaque; const float* proj; long long proj_stride; int proj_off; int mode;
  const uint8_t* legal; float* logits; float* scalar; float* priors;
};
int mzb_head_mma_launch_n(int n, const MmaHeadCall* calls, int B, int S, cudaStream_t stream);

// whole recurrent inference of a 16-channel network as one kernel (mzb_tower16.cu); projections feed the head kernels
bool mzb_tower16_supported(const mzb_resnet_model* m, int in_layout, int out_layout);
int mzb_tower16_recurrent(mzb_resnet_model* m, int B, const void* state_in, int in_layout, long long in_row_stride,
                          const int* in_slot, long long slot_stride, const int* action, void* state_out, int out_layout,
                          long long out_row_stride, long long out_off, float* proj_r, float* proj_vp, cudaStream_t s);

// a stem stage's residual tower with the image resident in shared memory (mzb_stem16.cu), in place on a padded buffer
bool mzb_stem16_supported(const std::vector<Block>& blocks, int H, int W, int C);
// s2_w != NULL: the tower is the pixel-pair stage (H lines of W pair rows); its output is not written back - the stride-2
// convolution that follows (s2_w [6][16][16] bf16, s2_shift [16]) runs on the resident image and y2 receives the
// H/2 x W padded image of the next stage (pad rows and the trailing halo written as zeros).
int mzb_stem16_tower(const std::vector<Block>& blocks, int B, int H, int W, __nv_bfloat16* x, cudaStream_t s,
                     const __nv_bfloat16* s2_w = nullptr, const float* s2_shift = nullptr, __nv_bfloat16* y2 = nullptr);
