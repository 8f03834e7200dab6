// softskel.cu -- soft erosion / dilation / skeleton levels of training/loss/soft_skeleton.py:6-37 on fp32 volumes
// [B][D][H][W], forward and backward with PyTorch's tie rules:
//   * -max_pool3d(-x) along one axis -> gradient to the FIRST minimum in scan order (out-of-volume taps ignored)
//   * max_pool3d 3x3x3               -> gradient to the FIRST maximum in (d,h,w) scan order
//   * torch.min(a,b)                 -> 0.5 / 0.5 on ties
//   * relu'(0) = 0
// Round-1 form: one streaming stencil kernel per level (neighbour taps served by L1/L2; 4 B/voxel of HBM per read
// or written volume).  The backward scatters with fp32 atomics.
#include <string.h>
#include "common.cuh"

namespace mvd {

struct Vol {
  int B, D, H, W;
  __device__ __forceinline__ long long n() const { return (long long)B * D * H * W; }
};

__device__ __forceinline__ void decode(long long i, const Vol& s, int& b, int& d, int& h, int& w) {
  w = (int)(i % s.W);
  long long t = i / s.W;
  h = (int)(t % s.H);
  t /= s.H;
  d = (int)(t % s.D);
  b = (int)(t / s.D);
}

// min over the 3-tap window along one axis; returns value, writes index offset (-1,0,1) of the first minimum
__device__ __forceinline__ float min3_first(const float* __restrict__ p, long long stride, int pos, int len, int& off) {
  float best = 0.f;
  bool have = false;
  off = 0;
#pragma unroll
  for (int k = -1; k <= 1; ++k) {
    int q = pos + k;
    if (q < 0 || q >= len) continue;
    float v = p[(long long)k * stride];
    if (!have || v < best) {
      best = v;
      off = k;
      have = true;
    }
  }
  return best;
}

__global__ void __launch_bounds__(256) erode_kernel(const float* __restrict__ in, float* __restrict__ out, Vol s) {
  const long long N = s.n();
  const long long sH = s.W, sD = (long long)s.H * s.W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    int b, d, h, w, o;
    decode(i, s, b, d, h, w);
    const float* p = in + i;
    float p1 = min3_first(p, sD, d, s.D, o);
    float p2 = min3_first(p, sH, h, s.H, o);
    float p3 = min3_first(p, 1, w, s.W, o);
    out[i] = fminf(fminf(p1, p2), p3);
  }
}

__global__ void __launch_bounds__(256) erode_bwd_kernel(const float* __restrict__ in, const float* __restrict__ gout,
                                                        float* __restrict__ gin, Vol s) {
  const long long N = s.n();
  const long long sH = s.W, sD = (long long)s.H * s.W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    const float g = gout[i];
    if (g == 0.f) continue;
    int b, d, h, w, o1, o2, o3;
    decode(i, s, b, d, h, w);
    const float* p = in + i;
    float p1 = min3_first(p, sD, d, s.D, o1);
    float p2 = min3_first(p, sH, h, s.H, o2);
    float p3 = min3_first(p, 1, w, s.W, o3);
    float m12 = fminf(p1, p2);
    float w12 = (m12 < p3) ? 1.f : ((m12 == p3) ? 0.5f : 0.f);
    float w3 = 1.f - w12;
    float w1 = w12 * ((p1 < p2) ? 1.f : ((p1 == p2) ? 0.5f : 0.f));
    float w2 = w12 - w1;
    if (w1 != 0.f) atomicAdd(gin + i + (long long)o1 * sD, g * w1);
    if (w2 != 0.f) atomicAdd(gin + i + (long long)o2 * sH, g * w2);
    if (w3 != 0.f) atomicAdd(gin + i + (long long)o3, g * w3);
  }
}

// 3x3x3 max; returns value and linear offset of the first maximum in (d,h,w) scan order
__device__ __forceinline__ float max27_first(const float* __restrict__ p, const Vol& s, int d, int h, int w,
                                             long long& off) {
  const long long sH = s.W, sD = (long long)s.H * s.W;
  float best = 0.f;
  bool have = false;
  off = 0;
#pragma unroll
  for (int kd = -1; kd <= 1; ++kd) {
    if (d + kd < 0 || d + kd >= s.D) continue;
#pragma unroll
    for (int kh = -1; kh <= 1; ++kh) {
      if (h + kh < 0 || h + kh >= s.H) continue;
#pragma unroll
      for (int kw = -1; kw <= 1; ++kw) {
        if (w + kw < 0 || w + kw >= s.W) continue;
        long long o = kd * sD + kh * sH + kw;
        float v = p[o];
        if (!have || v > best) {
          best = v;
          off = o;
          have = true;
        }
      }
    }
  }
  return best;
}

__global__ void __launch_bounds__(256) dilate_kernel(const float* __restrict__ in, float* __restrict__ out, Vol s) {
  const long long N = s.n();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    int b, d, h, w;
    long long o;
    decode(i, s, b, d, h, w);
    out[i] = max27_first(in + i, s, d, h, w, o);
  }
}

__global__ void __launch_bounds__(256) dilate_bwd_kernel(const float* __restrict__ in, const float* __restrict__ gout,
                                                         float gscale, float* __restrict__ gin, Vol s) {
  const long long N = s.n();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    const float g = gout[i] * gscale;
    if (g == 0.f) continue;
    int b, d, h, w;
    long long o;
    decode(i, s, b, d, h, w);
    max27_first(in + i, s, d, h, w, o);
    atomicAdd(gin + i + o, g);
  }
}

__global__ void __launch_bounds__(256) skel_update_kernel(const float* __restrict__ Ej, const float* __restrict__ Ej1,
                                                          const float* __restrict__ skel_in,
                                                          float* __restrict__ delta_out, float* __restrict__ skel_out,
                                                          int first, Vol s) {
  const long long N = s.n();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    int b, d, h, w;
    long long o;
    decode(i, s, b, d, h, w);
    float opened = max27_first(Ej1 + i, s, d, h, w, o);
    float delta = fmaxf(Ej[i] - opened, 0.f);
    if (delta_out) delta_out[i] = delta;
    float sk;
    if (first) sk = delta;
    else {
      // no FMA contraction: PyTorch rounds skel*delta before the subtraction (soft_skeleton.py:36)
      float prev = skel_in[i];
      sk = __fadd_rn(prev, fmaxf(__fsub_rn(delta, __fmul_rn(prev, delta)), 0.f));
    }
    skel_out[i] = sk;
  }
}

__global__ void __launch_bounds__(256) skel_chain_bwd_kernel(const float* __restrict__ delta,
                                                             const float* __restrict__ skel,
                                                             const float* __restrict__ g_skel,
                                                             float* __restrict__ g_delta, int L, long long N) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    float G = g_skel[i];
    for (int j = L - 1; j >= 1; --j) {
      float dl = delta[(long long)j * N + i];
      float sk = skel[(long long)(j - 1) * N + i];
      bool m = __fsub_rn(dl, __fmul_rn(sk, dl)) > 0.f;
      g_delta[(long long)j * N + i] = m ? G * (1.f - sk) : 0.f;
      G = m ? G * (1.f - dl) : G;
    }
    g_delta[i] = G;
  }
}

__global__ void __launch_bounds__(256) skel_level_bwd_kernel(const float* __restrict__ Ej1,
                                                             const float* __restrict__ delta_j,
                                                             const float* __restrict__ g_delta_j,
                                                             float* __restrict__ gEj, float* __restrict__ gEj1, Vol s) {
  const long long N = s.n();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    if (!(delta_j[i] > 0.f)) continue;
    const float g = g_delta_j[i];
    if (g == 0.f) continue;
    int b, d, h, w;
    long long o;
    decode(i, s, b, d, h, w);
    atomicAdd(gEj + i, g);
    max27_first(Ej1 + i, s, d, h, w, o);
    atomicAdd(gEj1 + i + o, -g);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Fused forward: up to kSkelMaxLevels levels of soft_skel (soft_skeleton.py:29-37) per launch, all iterations ON CHIP.
//
// A CTA owns a 32 x 16 x TZ tile of the volume.  It stages E_j0 for the tile plus a halo of R = n + 1 voxels in shared
// memory (cells outside the volume hold +inf: min-pooling ignores them; the 3x3x3 max-pool skips them explicitly),
// then per level l:   B = erode(A)  (7-point cross minimum, every interior cell of the box)
//                     opened = max27(B); delta = relu(A - opened); skel update          (tile cells only)
//                     swap(A, B)
// The halo shrinks by one valid layer per level, so after n levels the tile itself is still exact.  The running
// skeleton of the tile stays in shared memory.  HBM traffic of a pass: E_j0 (+ skel_in) read once, skel_out (+ the
// per-level E / delta / skel stacks the backward needs, when asked for) written once -- instead of 7 volume passes per
// level in the level-by-level form (mvd_soft_erode + mvd_skel_update).  Arithmetic is identical (min / max / the same
// rounded adds and multiplies), so results are bit-identical to the unfused kernels and to the reference.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kSkelMaxLevels = 4, kSkelTX = 32, kSkelTY = 16, kSkelThreads = 1024;

struct SkelPass {
  const float* E_in;                  // E_j0
  const float* skel_in;               // skeleton after level j0 - 1, NULL when j0 == 0
  float* E_next[kSkelMaxLevels];      // E_{j0 + l + 1} or NULL
  float* delta[kSkelMaxLevels];       // delta_{j0 + l} or NULL
  float* skel[kSkelMaxLevels];        // skeleton after level j0 + l or NULL (the last one is required)
  int n, TZ;
  Vol s;
  int tiles_x, tiles_y, tiles_z;
};

__global__ void __launch_bounds__(kSkelThreads) skel_fused_kernel(const __grid_constant__ SkelPass P) {
  extern __shared__ float sk_smem[];
  const int R = P.n + 1;
  const int SX = kSkelTX + 2 * R, SY = kSkelTY + 2 * R, SZ = P.TZ + 2 * R;
  const int SXY = SX * SY, SN = SXY * SZ;
  float* A = sk_smem;
  float* Bf = sk_smem + SN;
  float* SK = sk_smem + 2 * SN;           // running skeleton of the tile [TZ][TY][TX]
  const float INF = __int_as_float(0x7f800000);
  int t = blockIdx.x;
  const int tx = t % P.tiles_x; t /= P.tiles_x;
  const int ty = t % P.tiles_y; t /= P.tiles_y;
  const int tz = t % P.tiles_z;
  const int b = t / P.tiles_z;
  const int x0 = tx * kSkelTX, y0 = ty * kSkelTY, z0 = tz * P.TZ;
  const int W = P.s.W, H = P.s.H, D = P.s.D;
  const long long vol = (long long)D * H * W;
  const float* src = P.E_in + (long long)b * vol;
  // box cells inside the volume: [lx, hx) x [ly, hy) x [lz, hz) in box coordinates
  const int lx = max(0, R - x0), hx = min(SX, W - x0 + R);
  const int ly = max(0, R - y0), hy = min(SY, H - y0 + R);
  const int lz = max(0, R - z0), hz = min(SZ, D - z0 + R);
  const bool touches = (lx > 0) || (ly > 0) || (lz > 0) || (hx < SX) || (hy < SY) || (hz < SZ);
  // stage the box row by row (one warp per x row: no per-element index arithmetic)
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    // four rows per trip, all loads issued before the first store (one row per trip left a single load in flight per lane)
    for (int row0 = warp; row0 < SY * SZ; row0 += 4 * nwarps) {
      float v[4][2];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int row = row0 + u * nwarps;
        v[u][0] = v[u][1] = INF;
        if (row < SY * SZ) {
          const int sy = row % SY, sz = row / SY;
          const bool row_in = sy >= ly && sy < hy && sz >= lz && sz < hz;
          const float* grow = src + ((long long)(z0 - R + sz) * H + (y0 - R + sy)) * W + (x0 - R);
          if (row_in && lane >= lx && lane < hx) v[u][0] = grow[lane];
          if (row_in && lane + 32 < SX && lane + 32 >= lx && lane + 32 < hx) v[u][1] = grow[lane + 32];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int row = row0 + u * nwarps;
        if (row < SY * SZ) {
          float* arow = A + row * SX;
          if (lane < SX) arow[lane] = v[u][0];
          if (lane + 32 < SX) arow[lane + 32] = v[u][1];
        }
      }
    }
  }
  const int TN = kSkelTX * kSkelTY * P.TZ;
  if (P.skel_in) {
    const float* sk_src = P.skel_in + (long long)b * vol;
    for (int i = threadIdx.x; i < TN; i += kSkelThreads) {
      const int ix = i % kSkelTX, r = i / kSkelTX;
      const int iy = r % kSkelTY, iz = r / kSkelTY;
      const int x = x0 + ix, y = y0 + iy, z = z0 + iz;
      SK[i] = (x < W && y < H && z < D) ? sk_src[((long long)z * H + y) * W + x] : 0.f;
    }
  }
  __syncthreads();
  for (int l = 0; l < P.n; ++l) {
    // ---- B = erode(A) on the cells that can still be exact at this level: box shrunk by l + 1 layers.
    // Each thread walks columns along z with a 3-deep register window (5 shared loads per cell instead of 7).
    {
      const int lo = l + 1;
      const int wx = SX - 2 * lo, wy = SY - 2 * lo, z_lo = lo, z_hi = SZ - lo;   // [z_lo, z_hi)
      for (int p = threadIdx.x; p < wx * wy; p += kSkelThreads) {
        const int sx = lo + p % wx, sy = lo + p / wx;
        const bool col_in = sx >= lx && sx < hx && sy >= ly && sy < hy;
        int c = z_lo * SXY + sy * SX + sx;
        float a_prev = A[c - SXY], a_cur = A[c];
        for (int sz = z_lo; sz < z_hi; ++sz, c += SXY) {
          const float a_next = A[c + SXY];
          float m = fminf(fminf(a_cur, fminf(a_prev, a_next)),
                          fminf(fminf(A[c - 1], A[c + 1]), fminf(A[c - SX], A[c + SX])));
          if (touches && !(col_in && sz >= lz && sz < hz)) m = INF;
          Bf[c] = m;
          a_prev = a_cur;
          a_cur = a_next;
        }
      }
    }
    __syncthreads();
    // ---- tile cells: opened = max27(B) as a sliding maximum along z of in-plane 3x3 maxima (9 loads per cell),
    // delta, skeleton
    const bool first = (P.skel_in == nullptr) && (l == 0);
    float* gE = P.E_next[l] ? P.E_next[l] + (long long)b * vol : nullptr;
    float* gD = P.delta[l] ? P.delta[l] + (long long)b * vol : nullptr;
    float* gS = P.skel[l] ? P.skel[l] + (long long)b * vol : nullptr;
    for (int p = threadIdx.x; p < kSkelTX * kSkelTY; p += kSkelThreads) {
      const int ix = p % kSkelTX, iy = p / kSkelTX;
      const int x = x0 + ix, y = y0 + iy;
      if (x >= W || y >= H) continue;
      int c = (R - 1) * SXY + (iy + R) * SX + (ix + R);      // plane z = -1 of the tile
      auto plane_max = [&](int cc) -> float {
        float m = -INF;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            float v = Bf[cc + dy * SX + dx];
            if (touches && v == INF) v = -INF;
            m = fmaxf(m, v);
          }
        return m;
      };
      float m_prev = plane_max(c), m_cur = plane_max(c + SXY);
      c += SXY;                                              // plane z = 0
      const int nz = min(P.TZ, D - z0);
      for (int iz = 0; iz < nz; ++iz, c += SXY) {
        const float m_next = plane_max(c + SXY);
        const float opened = fmaxf(m_cur, fmaxf(m_prev, m_next));
        m_prev = m_cur;
        m_cur = m_next;
        const float delta = fmaxf(A[c] - opened, 0.f);
        const int i = (iz * kSkelTY + iy) * kSkelTX + ix;
        float sk;
        if (first) sk = delta;
        else {
          // no FMA contraction: PyTorch rounds skel*delta before the subtraction (soft_skeleton.py:36)
          const float prev = SK[i];
          sk = __fadd_rn(prev, fmaxf(__fsub_rn(delta, __fmul_rn(prev, delta)), 0.f));
        }
        SK[i] = sk;
        const long long g = ((long long)(z0 + iz) * H + y) * W + x;
        if (gE) gE[g] = Bf[c];
        if (gD) gD[g] = delta;
        if (gS) gS[g] = sk;
      }
    }
    __syncthreads();
    float* tmp = A; A = Bf; Bf = tmp;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Fused BACKWARD: up to kSkelBwdLevels consecutive soft_skel levels j = a+n-1 ... a per launch, all scatter-adds in
// SHARED memory (no global atomics, no per-level volume passes), delta_j recomputed from the E volumes (no delta stack).
//
// Per level j the backward of   delta_j = relu(E_j - dilate(E_{j+1})),   E_{j+1} = erode(E_j),
//                               skel_j = skel_{j-1} + relu(delta_j - skel_{j-1} * delta_j)      (soft_skeleton.py:29-37)
// is  (chain)   m = delta_j - skel_{j-1}*delta_j > 0;  g_delta = m ? G (1 - skel_{j-1}) : 0;  G <- m ? G (1 - delta_j) : G
//               (j = 0: g_delta = G),   gd = delta_j > 0 ? g_delta : 0
//     (A)       gE_j[v] += gd;   gE_{j+1}[argmax27 E_{j+1} around v] -= gd        (first maximum in scan order)
//     (B)       gE_j[first minima of the three axis windows of E_j around w] += gE_{j+1}[w] * (1, .5/.5 on ties)
// A CTA owns a 32 x 16 x 8 tile.  For the tile's gE_a to be complete, gE_{a+l} must be complete on the tile grown by l
// voxels, gd_{a+l} is needed on the tile grown by l + 2 and the chain state G on the tile grown by n + 1: three gradient
// boxes (G, gE upper, gE lower) with halo n + 1 plus one E box with halo n + 2 live in shared memory; the E box is
// refilled per phase (E_{j+1} for A, E_j for B).  Between launches only G and the upper gE travel through HBM.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kSkelBwdLevels = 2, kBwdTX = 32, kBwdTY = 16, kBwdTZ = 8, kBwdThreads = 1024;

struct SkelBwdPass {
  const float* E[kSkelBwdLevels + 1];        // E_a ... E_{a+n}
  const float* skel_prev[kSkelBwdLevels];    // skeleton after level a+l-1 (l = 0..n-1); [0] is NULL when a == 0
  const float* G_in;                         // d/d skel entering level a+n-1
  const float* gE_top_in;                    // partial gE_{a+n} from the launch above, NULL = zero
  float* gE_out;                             // gE_a (final gradient of the input when a == 0)
  float* G_out;                              // chain state below level a (NULL when a == 0)
  int n, first_is_level0;
  Vol s;
  int tiles_x, tiles_y, tiles_z;
};

struct Box {   // a halo'd box around the tile: cell (x, y, z) in tile coordinates -halo .. T+halo-1
  int halo, nx, ny, nz;
  __device__ __forceinline__ int idx(int x, int y, int z) const { return ((z + halo) * ny + (y + halo)) * nx + (x + halo); }
  __device__ __forceinline__ int cells() const { return nx * ny * nz; }
};

__global__ void __launch_bounds__(kBwdThreads) skel_bwd_fused_kernel(const __grid_constant__ SkelBwdPass P) {
  extern __shared__ float sb_smem[];
  const int n = P.n, R = n + 1;
  Box gb{R, kBwdTX + 2 * R, kBwdTY + 2 * R, kBwdTZ + 2 * R};
  Box eb{R + 1, kBwdTX + 2 * R + 2, kBwdTY + 2 * R + 2, kBwdTZ + 2 * R + 2};
  float* G = sb_smem;
  float* gUp = G + gb.cells();
  float* gLo = gUp + gb.cells();
  float* Eb = gLo + gb.cells();
  const float INF = __int_as_float(0x7f800000);
  int t = blockIdx.x;
  const int tx = t % P.tiles_x; t /= P.tiles_x;
  const int ty = t % P.tiles_y; t /= P.tiles_y;
  const int tz = t % P.tiles_z;
  const int b = t / P.tiles_z;
  const int x0 = tx * kBwdTX, y0 = ty * kBwdTY, z0 = tz * kBwdTZ;
  const int W = P.s.W, H = P.s.H, D = P.s.D;
  const long long vol = (long long)D * H * W;
  const long long boff = (long long)b * vol;
  // (ncu: the first version spent most of its 627 M warp instructions on per-cell index arithmetic -- div / mod by the
  // run-time box extents, 64-bit offsets, six-way bounds tests.  Now: multiply-high division, unsigned bounds tests,
  // 32-bit offsets inside the sample.)
  auto inside = [&](int x, int y, int z) {
    return (unsigned)(x0 + x) < (unsigned)W && (unsigned)(y0 + y) < (unsigned)H && (unsigned)(z0 + z) < (unsigned)D;
  };
  auto gidx = [&](int x, int y, int z) { return boff + (long long)(((z0 + z) * H + (y0 + y)) * W + (x0 + x)); };
  // c -> (x, y, z) of an ex x ey x ez box: exact for c < 2^16 and extents <= 64 (magic = ceil(2^32 / d))
  struct Dec { unsigned ex, ey, mx, my; };
  auto make_dec = [](int ex, int ey) {
    Dec d;
    d.ex = (unsigned)ex; d.ey = (unsigned)ey;
    d.mx = (unsigned)((0x100000000ull + ex - 1) / (unsigned)ex);
    d.my = (unsigned)((0x100000000ull + ey - 1) / (unsigned)ey);
    return d;
  };
  auto decode = [](const Dec& d, int c, int h, int& x, int& y, int& z) {
    const unsigned r = __umulhi((unsigned)c, d.mx);
    x = (int)((unsigned)c - r * d.ex) - h;
    const unsigned q = __umulhi(r, d.my);
    y = (int)(r - q * d.ey) - h;
    z = (int)q - h;
  };
  // fill a box region (tile grown by `h`) from a global volume, `oob` outside the volume (src == nullptr: all `oob`)
  // (four independent global loads per thread in flight: with one load per trip the 512-thread CTA -- the only one its
  // 200 KB of boxes leave room for on the SM -- spent most of the kernel waiting for single loads: 1.04 ms per launch)
  auto load_box = [&](float* dst, const Box& bx, const float* src, int h, float oob) {
    const int ex = kBwdTX + 2 * h, ey = kBwdTY + 2 * h, ez = kBwdTZ + 2 * h;
    const int total = ex * ey * ez;
    const Dec dd = make_dec(ex, ey);
    for (int c0 = threadIdx.x; c0 < total; c0 += 4 * kBwdThreads) {
      float v[4];
      int di[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + u * kBwdThreads;
        v[u] = oob;
        di[u] = -1;
        if (c < total) {
          int x, y, z;
          decode(dd, c, h, x, y, z);
          di[u] = bx.idx(x, y, z);
          if (src && inside(x, y, z)) v[u] = __ldg(src + gidx(x, y, z));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (di[u] >= 0) dst[di[u]] = v[u];
    }
  };
  load_box(G, gb, P.G_in, R, 0.f);
  load_box(gUp, gb, P.gE_top_in, R, 0.f);
  load_box(gLo, gb, nullptr, R, 0.f);
  for (int l = n - 1; l >= 0; --l) {
    const bool level0 = (l == 0) && P.first_is_level0;
    // ---------------- phase A: chain, pointwise part, dilate routing (E box = E_{j+1} on the tile grown by l + 3)
    __syncthreads();
    load_box(Eb, eb, P.E[l + 1], l + 3, INF);
    __syncthreads();
    {
      const int h = l + 2;
      const int ex = kBwdTX + 2 * h, ey = kBwdTY + 2 * h, ez = kBwdTZ + 2 * h;
      const float* Ej = P.E[l];
      const float* skp = P.skel_prev[l];
      const Dec dd = make_dec(ex, ey);
      for (int c = threadIdx.x; c < ex * ey * ez; c += kBwdThreads) {
        int x, y, z;
        decode(dd, c, h, x, y, z);
        if (!inside(x, y, z)) continue;
        // the two global operands of the cell are requested first: they arrive while the 27-point scan runs
        const long long g = gidx(x, y, z);
        const float ej = __ldg(Ej + g);
        const float sk = level0 ? 0.f : __ldg(skp + g);
        // opened = max27(E_{j+1}) with the FIRST maximum in (z, y, x) scan order; out-of-volume cells hold +inf: skipped
        float best = 0.f;
        int bo = -1;
#pragma unroll
        for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
          for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
              const int o = eb.idx(x + dx, y + dy, z + dz);
              const float v = Eb[o];
              if (v != INF && (bo < 0 || v > best)) { best = v; bo = ((dz + 1) * 3 + (dy + 1)) * 3 + (dx + 1); }
            }
        const float delta = fmaxf(ej - best, 0.f);
        const int gi = gb.idx(x, y, z);
        float Gv = G[gi], gdel;
        if (level0) {
          gdel = Gv;
        } else {
          const bool m = __fsub_rn(delta, __fmul_rn(sk, delta)) > 0.f;
          gdel = m ? Gv * (1.f - sk) : 0.f;
          if (m) G[gi] = Gv * (1.f - delta);
        }
        if (!(delta > 0.f) || gdel == 0.f) continue;
        if (x >= -l && x < kBwdTX + l && y >= -l && y < kBwdTY + l && z >= -l && z < kBwdTZ + l) atomicAdd(&gLo[gi], gdel);
        const int dz = bo / 9 - 1, dy = (bo / 3) % 3 - 1, dx = bo % 3 - 1;
        const int ux = x + dx, uy = y + dy, uz = z + dz;
        if (ux >= -(l + 1) && ux < kBwdTX + l + 1 && uy >= -(l + 1) && uy < kBwdTY + l + 1 && uz >= -(l + 1) && uz < kBwdTZ + l + 1)
          atomicAdd(&gUp[gb.idx(ux, uy, uz)], -gdel);
      }
    }
    // ---------------- phase B: erosion routing gE_{j+1} -> gE_j (E box = E_j on the tile grown by l + 2)
    __syncthreads();
    load_box(Eb, eb, P.E[l], l + 2, INF);
    __syncthreads();
    {
      const int h = l + 1;
      const int ex = kBwdTX + 2 * h, ey = kBwdTY + 2 * h, ez = kBwdTZ + 2 * h;
      const Dec dd = make_dec(ex, ey);
      for (int c = threadIdx.x; c < ex * ey * ez; c += kBwdThreads) {
        int x, y, z;
        decode(dd, c, h, x, y, z);
        const float g = gUp[gb.idx(x, y, z)];
        if (g == 0.f || !inside(x, y, z)) continue;
        // first minimum of each axis window (out-of-volume taps hold +inf and can never win: the centre is finite)
        auto min3 = [&](int sx, int sy, int sz, int& off) {
          float best = Eb[eb.idx(x - sx, y - sy, z - sz)];
          off = -1;
          const float c0 = Eb[eb.idx(x, y, z)], c1 = Eb[eb.idx(x + sx, y + sy, z + sz)];
          if (c0 < best) { best = c0; off = 0; }
          if (c1 < best) { best = c1; off = 1; }
          return best;
        };
        int o1, o2, o3;
        const float p1 = min3(0, 0, 1, o1), p2 = min3(0, 1, 0, o2), p3 = min3(1, 0, 0, o3);
        const float m12 = fminf(p1, p2);
        const float w12 = (m12 < p3) ? 1.f : ((m12 == p3) ? 0.5f : 0.f);
        const float w3 = 1.f - w12;
        const float w1 = w12 * ((p1 < p2) ? 1.f : ((p1 == p2) ? 0.5f : 0.f));
        const float w2 = w12 - w1;
        auto put = [&](int ux, int uy, int uz, float v) {
          if (v != 0.f && ux >= -l && ux < kBwdTX + l && uy >= -l && uy < kBwdTY + l && uz >= -l && uz < kBwdTZ + l)
            atomicAdd(&gLo[gb.idx(ux, uy, uz)], v);
        };
        put(x, y, z + o1, g * w1);
        put(x, y + o2, z, g * w2);
        put(x + o3, y, z, g * w3);
      }
    }
    __syncthreads();
    // gE_j becomes the upper gradient of the next level; the other box is cleared for gE_{j-1}
    float* tmp = gUp; gUp = gLo; gLo = tmp;
    if (l > 0) load_box(gLo, gb, nullptr, R, 0.f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < kBwdTX * kBwdTY * kBwdTZ; c += kBwdThreads) {
    const int x = c % kBwdTX, r = c / kBwdTX;
    const int y = r % kBwdTY, z = r / kBwdTY;
    if (!inside(x, y, z)) continue;
    const long long g = gidx(x, y, z);
    P.gE_out[g] = gUp[gb.idx(x, y, z)];
    if (P.G_out) P.G_out[g] = G[gb.idx(x, y, z)];
  }
}

static int vol_grid(long long N) { return grid_for(N, 256 * 2, num_sms() * 16); }

}  // namespace mvd

using namespace mvd;

#define VOL_CHECK(name) MVD_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, name ": bad volume dims")

extern "C" {

int mvd_soft_erode(const float* in, float* out, int B, int D, int H, int W, mvd_stream_t stream) {
  MVD_REQUIRE(in && out && in != out, "soft_erode: bad pointers");
  VOL_CHECK("soft_erode");
  Vol s{B, D, H, W};
  erode_kernel<<<vol_grid((long long)B * D * H * W), 256, 0, (cudaStream_t)stream>>>(in, out, s);
  MVD_LAUNCH_CHECK("soft_erode");
  return MVD_OK;
}

int mvd_soft_dilate(const float* in, float* out, int B, int D, int H, int W, mvd_stream_t stream) {
  MVD_REQUIRE(in && out && in != out, "soft_dilate: bad pointers");
  VOL_CHECK("soft_dilate");
  Vol s{B, D, H, W};
  dilate_kernel<<<vol_grid((long long)B * D * H * W), 256, 0, (cudaStream_t)stream>>>(in, out, s);
  MVD_LAUNCH_CHECK("soft_dilate");
  return MVD_OK;
}

int mvd_soft_erode_bwd(const float* in, const float* gout, float* gin, int B, int D, int H, int W,
                       mvd_stream_t stream) {
  MVD_REQUIRE(in && gout && gin, "soft_erode_bwd: bad pointers");
  VOL_CHECK("soft_erode_bwd");
  Vol s{B, D, H, W};
  erode_bwd_kernel<<<vol_grid((long long)B * D * H * W), 256, 0, (cudaStream_t)stream>>>(in, gout, gin, s);
  MVD_LAUNCH_CHECK("soft_erode_bwd");
  return MVD_OK;
}

int mvd_soft_dilate_bwd(const float* in, const float* gout, float gscale, float* gin, int B, int D, int H, int W,
                        mvd_stream_t stream) {
  MVD_REQUIRE(in && gout && gin, "soft_dilate_bwd: bad pointers");
  VOL_CHECK("soft_dilate_bwd");
  Vol s{B, D, H, W};
  dilate_bwd_kernel<<<vol_grid((long long)B * D * H * W), 256, 0, (cudaStream_t)stream>>>(in, gout, gscale, gin, s);
  MVD_LAUNCH_CHECK("soft_dilate_bwd");
  return MVD_OK;
}

int mvd_skel_update(const float* Ej, const float* Ej1, const float* skel_in, float* delta_out, float* skel_out,
                    int first, int B, int D, int H, int W, mvd_stream_t stream) {
  MVD_REQUIRE(Ej && Ej1 && skel_out && (first || skel_in), "skel_update: bad pointers");
  VOL_CHECK("skel_update");
  Vol s{B, D, H, W};
  skel_update_kernel<<<vol_grid((long long)B * D * H * W), 256, 0, (cudaStream_t)stream>>>(Ej, Ej1, skel_in, delta_out,
                                                                                             skel_out, first, s);
  MVD_LAUNCH_CHECK("skel_update");
  return MVD_OK;
}

int mvd_soft_skel_fused(const float* E_in, const float* skel_in, int n_levels, float* const* E_next,
                        float* const* delta, float* const* skel, int B, int D, int H, int W, mvd_stream_t stream) {
  MVD_REQUIRE(E_in && E_next && delta && skel && n_levels >= 1 && n_levels <= kSkelMaxLevels,
              "soft_skel_fused: 1..%d levels per pass", kSkelMaxLevels);
  VOL_CHECK("soft_skel_fused");
  MVD_REQUIRE(skel[n_levels - 1] != nullptr, "soft_skel_fused: the last level's skeleton output is required");
  SkelPass P;
  memset(&P, 0, sizeof(P));
  P.E_in = E_in; P.skel_in = skel_in; P.n = n_levels;
  for (int l = 0; l < n_levels; ++l) { P.E_next[l] = E_next[l]; P.delta[l] = delta[l]; P.skel[l] = skel[l]; }
  P.s = Vol{B, D, H, W};
  const int R = n_levels + 1;
  // deepest tile whose two halo'd boxes + skeleton tile fit in ~200 KB of shared memory
  int TZ = 16;
  size_t smem = 0;
  for (; TZ >= 2; TZ -= 2) {
    const size_t sn = (size_t)(kSkelTX + 2 * R) * (kSkelTY + 2 * R) * (TZ + 2 * R);
    smem = (2 * sn + (size_t)kSkelTX * kSkelTY * TZ) * sizeof(float);
    if (smem <= 200 * 1024) break;
  }
  MVD_REQUIRE(TZ >= 2, "soft_skel_fused: tile does not fit shared memory");
  if (TZ > D) TZ = D;
  {
    const size_t sn = (size_t)(kSkelTX + 2 * R) * (kSkelTY + 2 * R) * (TZ + 2 * R);
    smem = (2 * sn + (size_t)kSkelTX * kSkelTY * TZ) * sizeof(float);
  }
  P.TZ = TZ;
  P.tiles_x = (W + kSkelTX - 1) / kSkelTX; P.tiles_y = (H + kSkelTY - 1) / kSkelTY; P.tiles_z = (D + TZ - 1) / TZ;
  const long long tiles = (long long)B * P.tiles_x * P.tiles_y * P.tiles_z;
  MVD_REQUIRE(tiles < (1LL << 31), "soft_skel_fused: too many tiles");
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    MVD_CUDA(cudaFuncSetAttribute(skel_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 204 * 1024));
    attr_smem = 204 * 1024;
  }
  skel_fused_kernel<<<(unsigned)tiles, kSkelThreads, smem, (cudaStream_t)stream>>>(P);
  MVD_LAUNCH_CHECK("soft_skel_fused");
  return MVD_OK;
}

int mvd_soft_skel_bwd_fused(const float* const* E, const float* const* skel_prev, int n_levels, int first_is_level0,
                            const float* G_in, const float* gE_top_in, float* gE_out, float* G_out, int B, int D, int H,
                            int W, mvd_stream_t stream) {
  MVD_REQUIRE(E && skel_prev && G_in && gE_out && n_levels >= 1 && n_levels <= kSkelBwdLevels,
              "soft_skel_bwd_fused: 1..%d levels per launch", kSkelBwdLevels);
  VOL_CHECK("soft_skel_bwd_fused");
  MVD_REQUIRE(first_is_level0 || G_out, "soft_skel_bwd_fused: G_out is required unless the launch ends at level 0");
  SkelBwdPass P;
  memset(&P, 0, sizeof(P));
  for (int l = 0; l <= n_levels; ++l) {
    MVD_REQUIRE(E[l] != nullptr, "soft_skel_bwd_fused: missing E volume %d", l);
    P.E[l] = E[l];
  }
  for (int l = 0; l < n_levels; ++l) {
    MVD_REQUIRE(skel_prev[l] != nullptr || (l == 0 && first_is_level0), "soft_skel_bwd_fused: missing skeleton %d", l);
    P.skel_prev[l] = skel_prev[l];
  }
  P.G_in = G_in; P.gE_top_in = gE_top_in; P.gE_out = gE_out; P.G_out = first_is_level0 ? nullptr : G_out;
  P.n = n_levels; P.first_is_level0 = first_is_level0 ? 1 : 0;
  P.s = Vol{B, D, H, W};
  P.tiles_x = (W + kBwdTX - 1) / kBwdTX; P.tiles_y = (H + kBwdTY - 1) / kBwdTY; P.tiles_z = (D + kBwdTZ - 1) / kBwdTZ;
  const long long tiles = (long long)B * P.tiles_x * P.tiles_y * P.tiles_z;
  MVD_REQUIRE(tiles < (1LL << 31), "soft_skel_bwd_fused: too many tiles");
  const int R = n_levels + 1;
  const size_t gcells = (size_t)(kBwdTX + 2 * R) * (kBwdTY + 2 * R) * (kBwdTZ + 2 * R);
  const size_t ecells = (size_t)(kBwdTX + 2 * R + 2) * (kBwdTY + 2 * R + 2) * (kBwdTZ + 2 * R + 2);
  const size_t smem = (3 * gcells + ecells) * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    MVD_CUDA(cudaFuncSetAttribute(skel_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    attr_done = true;
  }
  MVD_REQUIRE(smem <= 210 * 1024, "soft_skel_bwd_fused: boxes do not fit shared memory");
  skel_bwd_fused_kernel<<<(unsigned)tiles, kBwdThreads, smem, (cudaStream_t)stream>>>(P);
  MVD_LAUNCH_CHECK("soft_skel_bwd_fused");
  return MVD_OK;
}

int mvd_skel_chain_bwd(const float* delta, const float* skel, const float* g_skel, float* g_delta, int L, long long N,
                       mvd_stream_t stream) {
  MVD_REQUIRE(delta && skel && g_skel && g_delta && L >= 1 && N > 0, "skel_chain_bwd: bad arguments");
  skel_chain_bwd_kernel<<<vol_grid(N), 256, 0, (cudaStream_t)stream>>>(delta, skel, g_skel, g_delta, L, N);
  MVD_LAUNCH_CHECK("skel_chain_bwd");
  return MVD_OK;
}

int mvd_skel_level_bwd(const float* Ej1, const float* delta_j, const float* g_delta_j, float* gEj, float* gEj1, int B,
                       int D, int H, int W, mvd_stream_t stream) {
  MVD_REQUIRE(Ej1 && delta_j && g_delta_j && gEj && gEj1, "skel_level_bwd: bad pointers");
  VOL_CHECK("skel_level_bwd");
  Vol s{B, D, H, W};
  skel_level_bwd_kernel<<<vol_grid((long long)B * D * H * W), 256, 0, (cudaStream_t)stream>>>(Ej1, delta_j, g_delta_j,
                                                                                               gEj, gEj1, s);
  MVD_LAUNCH_CHECK("skel_level_bwd");
  return MVD_OK;
}

}  // extern "C"
