// conv_common.cuh -- internal interface between the conv entry points and their two implementations.
#pragma once
#include "common.cuh"

namespace mvd {
// CUDA-core tiles (conv_generic.cu)
int generic_fprop(const mvd_conv3d_args* a, cudaStream_t st);
int generic_dgrad(const mvd_conv3d_args* a, cudaStream_t st);
int generic_wgrad(const mvd_conv3d_args* a, cudaStream_t st);
// tcgen05 implicit GEMM (conv_tc.cu); *_supported() says whether the shape is covered
bool tc_fprop_supported(const mvd_conv3d_args* a);
bool tc_dgrad_supported(const mvd_conv3d_args* a);
bool tc_wgrad_supported(const mvd_conv3d_args* a);
int tc_fprop(const mvd_conv3d_args* a, cudaStream_t st);
int tc_dgrad(const mvd_conv3d_args* a, cudaStream_t st);
bool tc_dgrad_fuses_norm_bwd(const mvd_conv3d_args* a);   // a->norm_bwd goes into the dgrad epilogue (halo kernels)
int tc_wgrad(const mvd_conv3d_args* a, cudaStream_t st);
size_t tc_wgrad_workspace_bytes(const mvd_conv3d_args* a);   // fp32 scratch [taps][Cin][Cout]
size_t tc_splitk_workspace_bytes(const mvd_conv3d_args* a, int pass);   // fprop (0) / dgrad (1): split-K partials, 0 = not split
int tc_wgrad_begin(const mvd_conv3d_args* a, cudaStream_t st, float** scratch);   // validates + zeroes the workspace
int tc_wgrad_finish(const mvd_conv3d_args* a, cudaStream_t st);                   // scratch -> dw[co][ci][tap] (+ dbias)
// sub-pixel data gradient of the stride-2 3x3x3 conv with 32 input channels (conv_tc_subpixel.cu)
bool tc_subpixel_dgrad_supported(const mvd_conv3d_args* a);
int tc_subpixel_dgrad(const mvd_conv3d_args* a, cudaStream_t st);
// halo-plane forward of the stride-2 3x3x3 conv with 32 input channels (conv_tc_halo.cu)
bool tc_halo_s2_fprop_supported(const mvd_conv3d_args* a);
int tc_halo_s2_fprop(const mvd_conv3d_args* a, cudaStream_t st);
// sliding-window halo variant for 3x3x3 / stride 1 (conv_tc_wgrad_halo.cu)
bool tc_wgrad_halo_supported(const mvd_conv3d_args* a);
int tc_wgrad_halo_splits(const mvd_conv3d_args* a);     // CTAs per (ci, co) set = slices of the deterministic reduction
bool wgrad_deterministic();
void set_wgrad_deterministic(int on);
int tc_wgrad_halo(const mvd_conv3d_args* a, cudaStream_t st);
}  // namespace mvd
