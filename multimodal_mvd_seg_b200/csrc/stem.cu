// stem.cu -- the network's first convolution (Cin = 1 or 2 modalities -> 32 features, 3x3x3, stride 1).
// K = 27*Cin is too thin for an implicit GEMM over taps, so the stem is run as an explicit one:
//   im2col : X_col[v][j] = x[v + tap(j)][ci(j)], j = tap*Cin + ci, zero-padded to Kpad (32 or 64) columns -- one streaming
//            pass, 16-byte stores (HBM-bound: V*Kpad*2 B written, the 2*Cin-byte voxels read through L1/L2)
//   fprop  : the tcgen05 kernel with a single tap over X_col (a [V x Kpad] x [Kpad x 32] GEMM)
//   wgrad  : the tcgen05 wgrad kernel over the same X_col (kept from the forward), no dgrad (the input is data).
#include "common.cuh"

namespace mvd {

__global__ void __launch_bounds__(256) im2col_small_kernel(const bf16* __restrict__ x, int ldx, int B, int D, int H,
                                                           int W, int Cin, int kd, int kh, int kw, int pd, int ph,
                                                           int pw, bf16* __restrict__ out, int Kpad) {
  const int groups = Kpad >> 3;
  const long long V = (long long)B * D * H * W;
  const long long total = V * groups;
  const int Kreal = kd * kh * kw * Cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long v = i / groups;
    const int g = (int)(i - v * groups);
    long long t = v;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H); t /= H;
    const int d = (int)(t % D);
    const int b = (int)(t / D);
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = g * 8 + e;
      float val = 0.f;
      if (j < Kreal) {
        const int tap = j / Cin, ci = j - tap * Cin;
        const int tw = tap % kw, th = (tap / kw) % kh, td = tap / (kw * kh);
        const int z = d + td - pd, yy = h + th - ph, xx = w + tw - pw;
        if (z >= 0 && z < D && yy >= 0 && yy < H && xx >= 0 && xx < W)
          val = bf2f(x[((((long long)b * D + z) * H + yy) * W + xx) * ldx + ci]);
      }
      f[e] = val;
    }
    *reinterpret_cast<bf16x8*>(out + v * Kpad + g * 8) = pack8(f);
  }
}

}  // namespace mvd

using namespace mvd;

extern "C" int mvd_im2col_small(const void* x, int ldx, int B, int D, int H, int W, int Cin, int kd, int kh, int kw,
                                int pd, int ph, int pw, void* out, int Kpad, mvd_stream_t stream) {
  MVD_REQUIRE(x && out && B > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && ldx >= Cin, "im2col_small: bad arguments");
  MVD_REQUIRE(Kpad % 8 == 0 && Kpad >= kd * kh * kw * Cin && ((uintptr_t)out & 15) == 0,
              "im2col_small: Kpad must be a multiple of 8 covering taps*Cin, out 16-byte aligned");
  const long long total = (long long)B * D * H * W * (Kpad / 8);
  int grid = grid_for(total, 256, num_sms() * 16);
  im2col_small_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, B, D, H, W, Cin, kd, kh, kw, pd, ph,
                                                              pw, (bf16*)out, Kpad);
  MVD_LAUNCH_CHECK("im2col_small");
  return MVD_OK;
}
