// spconv_internal.cuh -- what the sparse-conv translation units share (not part of the C ABI).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace bevf {

// fused epilogue of every gather-GEMM kernel: bias, folded BatchNorm (x * scale + shift), residual (fp32 or bf16), ReLU
struct SpconvEpilogue {
  const float *bias;
  const float *scale;
  const float *shift;
  const float *residual;
  const __nv_bfloat16 *residual_bf16;
  int relu;
};

// register-gather kernel for the narrow layers (spconv_rg.cu).  Its weight image (fragment order) follows the UMMA image in
// the packed-weight buffer of bevf_spconv_pack_weight_bf16 whenever spconv_rg_supported(cin_pad, cout).
bool spconv_rg_supported(int cin_pad, int cout);
int spconv_rg_pack_weight(const float *w_okc, void *w_rg, int kv, int cin, int cin_pad, int cout, cudaStream_t st);
int spconv_rg_launch(int cin_pad, int cout, const void *feats_bf16, const void *w_rg, const int *pair_fwd, int ld, int n_out,
                     const int *n_out_dev, int kv, const SpconvEpilogue &ep, float *out_f32, void *out_bf16, cudaStream_t st);

}  // namespace bevf
