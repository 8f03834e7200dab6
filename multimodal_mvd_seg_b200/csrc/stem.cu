// stem.cu -- the network's first convolution (Cin = 1 or 2 modalities -> 32 features, 3x3x3, stride 1).
// K = 27*Cin is too thin for an implicit GEMM over taps, so the stem is run as an explicit one:
//   im2col : X_col[v][j] = x[v + tap(j)][ci(j)], j = tap*Cin + ci, zero-padded to Kpad (32 or 64) columns -- one streaming
//            pass, 16-byte stores (HBM-bound: V*Kpad*2 B written, the 2*Cin-byte voxels staged through shared memory)
//   fprop  : the tcgen05 kernel with a single tap over X_col (a [V x Kpad] x [Kpad x 32] GEMM)
//   wgrad  : the tcgen05 wgrad kernel over the same X_col (kept from the forward), no dgrad (the input is data).
#include "common.cuh"

namespace mvd {

// block = one (b, d, h) line.  The 3 x 3 neighbouring input lines (zero outside the volume) are staged in shared
// memory -- whole lines with 16-byte loads when the input is dense (ldx == CIN), no per-element index arithmetic --
// then thread (w, g) assembles one 16-byte group of the voxel's KPAD-column row from shared memory (tap decode is
// compile-time) and a voxel's row is written by G adjacent threads.  Bound by the X_col write.
// Shared layout: line r = td*3 + th holds voxels [-PADV, W + PADV) at s_in[r*LP + (PADV + w)*CIN + ci]; PADV = 8 / CIN
// voxels = 16 bytes keeps the staged line 16-byte aligned.
template <int CIN, int KPAD>
__global__ void __launch_bounds__(256) im2col_small_kernel(const bf16* __restrict__ x, int ldx, int D, int H, int W,
                                                           bf16* __restrict__ out, int vec) {
  constexpr int G = KPAD / 8;
  constexpr int WPB = 256 / G;
  constexpr int KREAL = 27 * CIN;
  constexpr int PADV = 8 / CIN;
  extern __shared__ __align__(16) bf16 s_in[];
  const int LP = (W + 2 * PADV) * CIN;    // elements per staged line (multiple of 8 when W*CIN is)
  int line = blockIdx.x;                  // (b*D + d)*H + h
  const int h = line % H; line /= H;
  const int d = line % D;
  const int b = line / D;
  if (vec) {
    const int vpl = LP / 8;               // 16-byte vectors per staged line: [pad][W*CIN/8 data][pad]
    for (int i = threadIdx.x; i < 9 * vpl; i += 256) {
      const int r = i / vpl, j = i - r * vpl;
      const int z = d + r / 3 - 1, yy = h + r % 3 - 1;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (j >= 1 && j < vpl - 1 && z >= 0 && z < D && yy >= 0 && yy < H)
        v = *reinterpret_cast<const uint4*>(x + (((long long)b * D + z) * H + yy) * (long long)W * CIN + (j - 1) * 8);
      *reinterpret_cast<uint4*>(s_in + r * LP + j * 8) = v;
    }
  } else {
    for (int i = threadIdx.x; i < 9 * LP; i += 256) {
      const int r = i / LP, e = i - r * LP;
      const int ci = e % CIN, xx = e / CIN - PADV;
      const int z = d + r / 3 - 1, yy = h + r % 3 - 1;
      bf16 v = f2bf(0.f);
      if (z >= 0 && z < D && yy >= 0 && yy < H && xx >= 0 && xx < W)
        v = x[((((long long)b * D + z) * H + yy) * W + xx) * ldx + ci];
      s_in[i] = v;
    }
  }
  __syncthreads();
  const int g = threadIdx.x % G;
  const long long v0 = (((long long)b * D + d) * H + h) * W;
  for (int w = threadIdx.x / G; w < W; w += WPB) {
    bf16x8 o;
    bf16* oe = reinterpret_cast<bf16*>(&o);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = g * 8 + e;
      bf16 val = f2bf(0.f);
      if (j < KREAL) {
        const int tap = j / CIN, ci = j % CIN;
        const int tw = tap % 3, r = tap / 3;   // r = td*3 + th
        val = s_in[r * LP + (PADV + w + tw - 1) * CIN + ci];
      }
      oe[e] = val;
    }
    stg16(out + (v0 + w) * KPAD + g * 8, o);
  }
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
  MVD_REQUIRE((long long)B * D * H < (1LL << 31) && W <= 2048, "im2col_small: volume too large");
  const unsigned grid = (unsigned)((long long)B * D * H);
  const size_t smem = (size_t)9 * (W + 2 * (8 / Cin)) * Cin * sizeof(bf16);
  // whole-line 16-byte staging needs a dense input whose lines are multiples of 16 bytes
  const int vec = (ldx == Cin && (W * Cin) % 8 == 0 && ((uintptr_t)x & 15) == 0) ? 1 : 0;
  if (Cin == 1)
    im2col_small_kernel<1, 32><<<grid, 256, smem, (cudaStream_t)stream>>>((const bf16*)x, ldx, D, H, W, (bf16*)out, vec);
  else
    im2col_small_kernel<2, 64><<<grid, 256, smem, (cudaStream_t)stream>>>((const bf16*)x, ldx, D, H, W, (bf16*)out, vec);
  MVD_LAUNCH_CHECK("im2col_small");
  return MVD_OK;
}
