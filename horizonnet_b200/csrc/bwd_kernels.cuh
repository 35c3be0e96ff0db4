// Backward-pass building blocks (bwd_kernels.cu), fp32 halo-NHWC.  See the .cu for what each one computes.
#pragma once
#include "hn_common.cuh"

namespace hn {

int fill_f32(float* p, size_t n, float v, cudaStream_t st);
int transpose_f32(const float* in, float* out, int rows, int cols, cudaStream_t st);           // out[c][r] = in[r][c]

// dW[Cout][kh][kw][Cin] (zeroed here, then accumulated); in = the conv's forward input, dz = d(raw conv output)
int conv_wgrad_f32(const ConvDesc& d, const Act& in, const Act& dz, float* dw_ohwi, cudaStream_t st);
int ohwi_to_oihw(const float* in, float* out, int Cout, int Cin, int kh, int kw, cudaStream_t st);
// flipped / transposed weights in conv_f32's [K][N] packing, for conv_dgrad_f32
int pack_dgrad_weight(const float* w_oihw, float* out, int Cout, int Cin, int kh, int kw, cudaStream_t st);
// din (+)= conv_transpose(dz, W).  dz: halo 1 with circular halo columns; dilate_scratch: din.numel()/Cin*Cout floats
// (strided convs only); ones / zeros: >= Cin floats of 1 / 0.
int conv_dgrad_f32(const ConvDesc& d, const float* wd_packed, const Act& dz, const Act& din, bool accumulate,
                   float* dilate_scratch, const float* ones, const float* zeros, cudaStream_t st);

// bn[4*C]: scale, shift, mean of z, invstd.  train: from the batch sums (bn_batch_stats) + running-stat update;
// frozen: from the running statistics.
int bn_finalize_full(const double* sums, long long count, const float* gamma, const float* beta, const float* bias,
                     float* running_mean, float* running_var, double factor, bool train, float* bn, int C, cudaStream_t st);
int bn_apply_fwd(const Act& z, const float* bn, const float* res, bool relu, const Act& y, cudaStream_t st);
// dz (with halo columns), dres += relu-masked dy, parameter gradients; sums: 2*C doubles of scratch
int bn_bwd(const Act& dy, const Act& y, const Act& z, const float* bn, bool train, bool relu, double* sums, const Act& dz,
           float* dres, float* dgamma, float* dbeta, float* dbias, cudaStream_t st);

int maxpool_bwd(const Act& x, const Act& dp, float* dx, cudaStream_t st);                      // dx accumulated (atomics)
int ghc_to_sequence_bwd(const float* dseq, const Act dghc[4], cudaStream_t st);                // accumulated (atomics)
int head_bwd(const float* dbon, const float* dcor, const float* rnn, const float* w, float* drnn, float* dw, float* db, int T,
             int B, cudaStream_t st);
int col_sum(const float* x, size_t rows, int cols, float* out, cudaStream_t st);

int lstm_gather(const float* hout, const float* xp, float* hprev, float* xpd, int T, int B, cudaStream_t st);
int lstm_cell_scan(float* gates, float* cell, int T, int B, cudaStream_t st);
int lstm_bwd_steps(const float* dout, const float* gates, const float* cell, const float* whh_f, const float* whh_b,
                   float* dgates, float* dc, int T, int B, cudaStream_t st);
int stem_input_nhwc(const float* x, int in_channels, float* out, int B, cudaStream_t st);

}  // namespace hn
