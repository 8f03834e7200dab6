// stem.cu -- the network's first convolution (Cin = 1 or 2 modalities -> 32 features, 3x3x3, stride 1).
// K = 27*Cin is too thin for an implicit GEMM over taps, so the stem is run as an explicit one:
//   im2col : X_col[v][j] = x[v + tap(j)][ci(j)], j = tap*Cin + ci, zero-padded to Kpad (32 or 64) columns -- one streaming
//            pass, 16-byte stores (HBM-bound: V*Kpad*2 B written, the 2*Cin-byte voxels read through L1/L2)
//   fprop  : the tcgen05 kernel with a single tap over X_col (a [V x Kpad] x [Kpad x 32] GEMM)
//   wgrad  : the tcgen05 wgrad kernel over the same X_col (kept from the forward), no dgrad (the input is data).
#include "common.cuh"

namespace mvd {

// block = 256 threads = (256 / G) consecutive w positions of one (b, d, h) line x G = KPAD/8 column groups; thread
// (w, g) writes one 16-byte group, so a voxel's KPAD*2-byte row is written by G adjacent threads (coalesced) and no
// per-thread integer division is needed (tap decode is compile-time for the 3x3x3 stem).
template <int CIN, int KPAD>
__global__ void __launch_bounds__(256) im2col_small_kernel(const bf16* __restrict__ x, int ldx, int D, int H, int W,
                                                           bf16* __restrict__ out) {
  constexpr int G = KPAD / 8;
  constexpr int WPB = 256 / G;
  constexpr int KREAL = 27 * CIN;
  const int g = threadIdx.x % G;
  const int w = blockIdx.x * WPB + threadIdx.x / G;
  int line = blockIdx.y;                 // (b*D + d)*H + h
  const int h = line % H; line /= H;
  const int d = line % D;
  const int b = line / D;
  if (w >= W) return;
  float f[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int j = g * 8 + e;
    float val = 0.f;
    if (j < KREAL) {
      const int tap = j / CIN, ci = j % CIN;
      const int tw = tap % 3, th = (tap / 3) % 3, td = tap / 9;
      const int z = d + td - 1, yy = h + th - 1, xx = w + tw - 1;
      if (z >= 0 && z < D && yy >= 0 && yy < H && xx >= 0 && xx < W)
        val = bf2f(x[((((long long)b * D + z) * H + yy) * W + xx) * ldx + ci]);
    }
    f[e] = val;
  }
  const long long v = (((long long)b * D + d) * H + h) * W + w;
  *reinterpret_cast<bf16x8*>(out + v * KPAD + g * 8) = pack8(f);
}

}  // namespace mvd

using namespace mvd;

extern "C" int mvd_im2col_small(const void* x, int ldx, int B, int D, int H, int W, int Cin, int kd, int kh, int kw,
                                int pd, int ph, int pw, void* out, int Kpad, mvd_stream_t stream) {
  MVD_REQUIRE(x && out && B > 0 && D > 0 && H > 0 && W > 0 && Cin > 0 && ldx >= Cin, "im2col_small: bad arguments");
  MVD_REQUIRE(Kpad % 8 == 0 && Kpad >= kd * kh * kw * Cin && ((uintptr_t)out & 15) == 0,
              "im2col_small: Kpad must be a multiple of 8 covering taps*Cin, out 16-byte aligned");
  MVD_REQUIRE(kd == 3 && kh == 3 && kw == 3 && pd == 1 && ph == 1 && pw == 1 && (Cin == 1 || Cin == 2) &&
                  Kpad == (Cin == 1 ? 32 : 64), "im2col_small: built for the 3x3x3 pad-1 stem with 1 or 2 input channels");
  const int wpb = 256 / (Kpad / 8);
  dim3 grid((W + wpb - 1) / wpb, (unsigned)((long long)B * D * H));
  if (Cin == 1)
    im2col_small_kernel<1, 32><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, D, H, W, (bf16*)out);
  else
    im2col_small_kernel<2, 64><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, D, H, W, (bf16*)out);
  MVD_LAUNCH_CHECK("im2col_small");
  return MVD_OK;
}
