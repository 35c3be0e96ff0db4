// Split-bf16 tcgen05 convolution entry points (conv_tc.cu).
#pragma once
#include <cuda_bf16.h>
#include "hn_common.cuh"

namespace hn {

bool conv_tc_supported(const ConvDesc& d, const Act& in, const Act& out);
// fp32 halo-NHWC in/out convenience wrapper (unit tests): splits, runs the plane kernel, merges.
int conv_tc(const ConvDesc& d, const Act& in, const Act& out, const float* residual, cudaStream_t st);
// The real thing: operands and result are bf16 hi/lo plane pairs (hi plane first, lo plane at +numel).
int conv_tc_planes(const ConvDesc& d, const __nv_bfloat16* wq, const Act& in, const __nv_bfloat16* in_planes,
                   const Act& out, __nv_bfloat16* out_planes, float* out_f32, const __nv_bfloat16* res_planes,
                   cudaStream_t st);
int split_planes(const float* in, __nv_bfloat16* out, size_t n, cudaStream_t st);
int merge_planes(const __nv_bfloat16* in, float* out, size_t n, cudaStream_t st);
int pack_weight_tc(const float* w_oihw, __nv_bfloat16* out, int Cout, int Cin, int kh, int kw, cudaStream_t st);

}  // namespace hn
