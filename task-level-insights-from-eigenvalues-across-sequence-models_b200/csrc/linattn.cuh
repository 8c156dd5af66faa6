// linattn.cuh -- parameters shared by the causal linear-attention layer kernels (k1_attn.cu: recurrent forms; k9_linattn_mma.cu: chunked tensor-core form)
#pragma once
#include "common.cuh"

namespace eigb200 {

struct LinAttnParams {
  const float* q; const float* k; const float* v; int64_t ld; const float* gate;
  int phi_elu, normalise; float kscale;
  float* out; int64_t ldo; int64_t T; int H, d, dv;
  // fused depthwise causal conv + SiLU in front of phi (chunked kernel only): weights (C, kconv), bias (C); conv_ch_* = conv channel of the matrix's
  // first column (head 0), < 0: that matrix is taken as it is
  const float* conv_w; const float* conv_b; int kconv; int conv_ch_q, conv_ch_k, conv_ch_v;
};

// chunked tensor-core form (d = dv = 64, 16-byte aligned rows); returns false when the shape is not its own
bool linattn_mma_supported(const LinAttnParams& p);
cudaError_t linattn_mma_launch(const LinAttnParams& p, int64_t B, cudaStream_t stream);

}  // namespace eigb200
