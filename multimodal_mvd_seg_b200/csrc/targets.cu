// targets.cu -- deep-supervision target production on the GPU.
//
// Replaces DownsampleSegForDSTransform2.__call__ (nnunetv2/training/data_augmentation/custom_transforms/
// deep_supervision_donwsampling.py:27-55), which runs batchgenerators' resize_segmentation(order = 0) per (sample,
// channel) on the CPU workers: nearest-neighbour resampling with the pixel-centre convention of skimage.transform.resize
// / scipy.ndimage.zoom(grid_mode=True):   src = floor((o + 0.5) * I / O)   per axis, clipped to I - 1
// (integer form ((2 o + 1) * I) / (2 * O); for the factor-2 pyramids of 3d_fullres this is src = 2 o + 1).
// All scales of one batch are produced by ONE launch from the full-resolution segmentation already in HBM, so a step
// uploads the full-resolution target only (the coarser scales are 1/8 + 1/64 + ... = 14 % of the target bytes).
#include "common.cuh"

namespace mvd {

constexpr int kMaxDsScales = 8;
struct DsParams {
  const float* src;
  int BC, Di, Hi, Wi;
  int n;
  float* dst[kMaxDsScales];
  int Do[kMaxDsScales], Ho[kMaxDsScales], Wo[kMaxDsScales];
  long long begin[kMaxDsScales + 1];   // running voxel count over the scales (per (b, c) plane stack it is x BC)
};

__global__ void __launch_bounds__(256) ds_targets_kernel(const __grid_constant__ DsParams P) {
  const long long total = P.begin[P.n];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int s = 0;
    while (s + 1 < P.n && i >= P.begin[s + 1]) ++s;
    long long j = i - P.begin[s];
    const int Wo = P.Wo[s], Ho = P.Ho[s], Do = P.Do[s];
    const int ow = (int)(j % Wo); j /= Wo;
    const int oh = (int)(j % Ho); j /= Ho;
    const int od = (int)(j % Do);
    const long long bc = j / Do;
    int iw = (int)(((2LL * ow + 1) * P.Wi) / (2LL * Wo)), ih = (int)(((2LL * oh + 1) * P.Hi) / (2LL * Ho)),
        id = (int)(((2LL * od + 1) * P.Di) / (2LL * Do));
    iw = min(iw, P.Wi - 1); ih = min(ih, P.Hi - 1); id = min(id, P.Di - 1);
    P.dst[s][i - P.begin[s]] = __ldg(P.src + ((bc * P.Di + id) * P.Hi + ih) * (long long)P.Wi + iw);
  }
}

}  // namespace mvd

using namespace mvd;

extern "C" int mvd_downsample_seg_nearest(const float* seg, int BC, int Di, int Hi, int Wi, int n_scales,
                                          float* const* dst, const int* out_dhw, mvd_stream_t stream) {
  MVD_REQUIRE(seg && dst && out_dhw && BC > 0 && Di > 0 && Hi > 0 && Wi > 0, "downsample_seg_nearest: bad arguments");
  MVD_REQUIRE(n_scales >= 1 && n_scales <= kMaxDsScales, "downsample_seg_nearest: 1..%d scales per call", kMaxDsScales);
  DsParams P;
  P.src = seg; P.BC = BC; P.Di = Di; P.Hi = Hi; P.Wi = Wi; P.n = n_scales;
  P.begin[0] = 0;
  for (int s = 0; s < n_scales; ++s) {
    MVD_REQUIRE(dst[s] && out_dhw[3 * s] > 0 && out_dhw[3 * s + 1] > 0 && out_dhw[3 * s + 2] > 0,
                "downsample_seg_nearest: scale %d has a null output or an empty extent", s);
    P.dst[s] = dst[s];
    P.Do[s] = out_dhw[3 * s]; P.Ho[s] = out_dhw[3 * s + 1]; P.Wo[s] = out_dhw[3 * s + 2];
    P.begin[s + 1] = P.begin[s] + (long long)BC * P.Do[s] * P.Ho[s] * P.Wo[s];
  }
  const int grid = grid_for(P.begin[n_scales], 256 * 4, num_sms() * 8);
  ds_targets_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
  MVD_LAUNCH_CHECK("downsample_seg_nearest");
  return MVD_OK;
}
