// placeholder until the tcgen05 kernels land
#include "conv_common.cuh"
namespace mvd {
bool tc_fprop_supported(const mvd_conv3d_args*) { return false; }
bool tc_dgrad_supported(const mvd_conv3d_args*) { return false; }
bool tc_wgrad_supported(const mvd_conv3d_args*) { return false; }
int tc_fprop(const mvd_conv3d_args*, cudaStream_t) { return MVD_ERR_UNSUPPORTED; }
int tc_dgrad(const mvd_conv3d_args*, cudaStream_t) { return MVD_ERR_UNSUPPORTED; }
int tc_wgrad(const mvd_conv3d_args*, cudaStream_t) { return MVD_ERR_UNSUPPORTED; }
size_t tc_wgrad_workspace_bytes(const mvd_conv3d_args*) { return 0; }
}
extern "C" int mvd_tc_selftest(float*, int, mvd_stream_t) { return MVD_ERR_UNSUPPORTED; }
