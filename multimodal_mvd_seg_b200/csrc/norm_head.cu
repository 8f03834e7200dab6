// norm_head.cu -- InstanceNorm3d + LeakyReLU of the LAST decoder block folded into its only consumer, the 1x1x1
// segmentation head (UNetDecoder.py:67-70, 104-121; get_network_from_plans.py:41-44).
//
// At the full-resolution stage the normalised activation z = lrelu(IN(y)) (2 x 128^3 x 32 bf16 = 268 MB) is read by
// nothing but the head.  The unfused path writes z, reads it back for the logits, and in backward writes the head's dz
// (another 268 MB) only for the InstanceNorm backward to read it twice.  Here z and dz never exist in HBM:
//   inorm_lrelu_head_fwd        : y -> (normalise, LeakyReLU, x W^T + b) -> logits              (reads y, writes 8 B / voxel)
//   inorm_lrelu_head_bwd_stats  : y, dlogits -> InstanceNorm backward sums  S1 = sum g', S2 = sum g' xhat  with
//                                 dz = dlogits W recomputed per voxel, plus the head's dW = sum dlogits (x) z and db
//   inorm_lrelu_head_bwd_apply  : y, dlogits -> dy (and the conv-bias gradient sum), again from dlogits directly
// Rounding points are those of the unfused path (z and dz rounded to bf16 where the reference materialises them).
// Shape: C = 32 features (4 channel groups of 8 = 4 consecutive lanes per voxel), K = 4 classes, dense logits rows.
#include "common.cuh"

namespace mvd {
namespace {

constexpr int HC = 32, HK = 4, HCG = HC / 8, kThreads = 256, kRows = kThreads / HCG, kStage = 4;

struct NormConst {      // per (sample, channel group): everything a thread needs for its 8 channels
  float sc[8], sh[8], rs[8], xm[8];
};

// scale / shift / rstd / -mean*rstd of the thread's 8 channels from the fp64 sums
__device__ __forceinline__ void load_norm(const double* __restrict__ stats, const float* __restrict__ gamma,
                                          const float* __restrict__ beta, int b, int cg, long long V, float eps,
                                          NormConst& n) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = cg * 8 + k;
    const double s1 = stats[((long long)b * HC + c) * 2 + 0], s2 = stats[((long long)b * HC + c) * 2 + 1];
    const double m = s1 / (double)V;
    double var = s2 / (double)V - m * m;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    n.sc[k] = g * rstd;
    n.sh[k] = be - (float)m * g * rstd;
    n.rs[k] = rstd;
    n.xm[k] = -(float)m * rstd;
  }
}

__device__ __forceinline__ void st_stream16(bf16* p, const bf16x8& v) {
  const uint4 u = *reinterpret_cast<const uint4*>(&v);
  asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t smem_addr, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void unpack4(uint2 u, float* g) {
  g[0] = __uint_as_float(u.x << 16); g[1] = __uint_as_float(u.x & 0xffff0000u);
  g[2] = __uint_as_float(u.y << 16); g[3] = __uint_as_float(u.y & 0xffff0000u);
}

// ---- forward: logits[v][k] = bf16( sum_c z[v][c] * bf16(W[k][c]) + bf16(b[k]) ),  z = bf16 lrelu(bf16 IN(y)) ---------
__global__ void __launch_bounds__(kThreads) inorm_lrelu_head_fwd_kernel(
    const bf16* __restrict__ y, int ldy, const double* __restrict__ stats, const float* __restrict__ gamma,
    const float* __restrict__ beta, const float* __restrict__ w, const float* __restrict__ bias,
    bf16* __restrict__ logits, long long V, float eps, float slope) {
  extern __shared__ __align__(16) uint4 ring[];
  const int b = blockIdx.y;
  const int cg = threadIdx.x % HCG, r = threadIdx.x / HCG;
  NormConst n;
  load_norm(stats, gamma, beta, b, cg, V, eps, n);
  float wr[HK][8], bz[HK];
#pragma unroll
  for (int k = 0; k < HK; ++k) {
    bz[k] = bias ? round_bf(bias[k]) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) wr[k][j] = round_bf(w[k * HC + cg * 8 + j]);
  }
  const bf16* yb = y + (long long)b * V * ldy;
  bf16* lb = logits + (long long)b * V * HK;
  const long long step = (long long)gridDim.x * kRows;
  const long long v0 = (long long)blockIdx.x * kRows + r;
  // every lane of a warp runs the same number of steps (the class sums are reduced with shuffles)
  const long long vw = (long long)blockIdx.x * kRows + (threadIdx.x & ~31) / HCG;
  const long long iters = vw < V ? (V - vw + step - 1) / step : 0;
  const uint4* mine = ring + threadIdx.x;
  const uint32_t mine_u = (uint32_t)__cvta_generic_to_shared(mine);
  auto issue = [&](long long i) {
    const long long v = v0 + i * step;
    if (i < iters && v < V) cp_async16(mine_u + (uint32_t)(i & (kStage - 1)) * (kThreads * 16), yb + v * ldy + cg * 8);
    cp_async_commit();
  };
  for (int i = 0; i < kStage - 1; ++i) issue(i);
  for (long long i = 0; i < iters; ++i) {
    issue(i + kStage - 1);
    cp_async_wait<kStage - 1>();
    const long long v = v0 + i * step;
    const bool ok = v < V;
    float f[8];
    if (ok) {
      const uint4 u = mine[(int)(i & (kStage - 1)) * kThreads];
      unpack8(*reinterpret_cast<const bf16x8*>(&u), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float t = round_bf(fmaf(f[k], n.sc[k], n.sh[k]));
        f[k] = t > 0.f ? t : slope * t;
      }
#pragma unroll
      for (int k = 0; k < 8; k += 2) round_bf2(f[k], f[k + 1]);     // z as the unfused path stores it
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = 0.f;
    }
    float acc[HK];
#pragma unroll
    for (int k = 0; k < HK; ++k) {
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) a = fmaf(f[j], wr[k][j], a);
#pragma unroll
      for (int off = HCG / 2; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
      acc[k] = a + bz[k];
    }
    if (ok && cg == 0) {
      __nv_bfloat162 a = __floats2bfloat162_rn(acc[0], acc[1]), c = __floats2bfloat162_rn(acc[2], acc[3]);
      *reinterpret_cast<uint2*>(lb + v * HK) = make_uint2(*reinterpret_cast<unsigned*>(&a), *reinterpret_cast<unsigned*>(&c));
    }
  }
  cp_async_wait<0>();
}

// ---- backward, shared sweep: per voxel row the thread gets its 8 y values and the voxel's 4 logit gradients ------------
template <typename Body>
__device__ __forceinline__ void sweep_y_dl(const bf16* __restrict__ yb, int ldy, const bf16* __restrict__ dlb,
                                           long long V, int cg, int r, uint4* ring, Body body) {
  const long long step = (long long)gridDim.x * kRows;
  const long long v0 = (long long)blockIdx.x * kRows + r;
  const long long iters = v0 < V ? (V - v0 + step - 1) / step : 0;
  const uint4* mine = ring + threadIdx.x;
  const uint2* mine_g = reinterpret_cast<const uint2*>(ring + kStage * kThreads) + threadIdx.x;
  const uint32_t mine_u = (uint32_t)__cvta_generic_to_shared(mine);
  const uint32_t mine_gu = (uint32_t)__cvta_generic_to_shared(mine_g);
  auto issue = [&](long long i) {
    if (i < iters) {
      const long long v = v0 + i * step;
      const int st = (int)(i & (kStage - 1));
      cp_async16(mine_u + (uint32_t)st * (kThreads * 16), yb + v * ldy + cg * 8);
      cp_async8(mine_gu + (uint32_t)st * (kThreads * 8), dlb + v * HK);
    }
    cp_async_commit();
  };
  for (int i = 0; i < kStage - 1; ++i) issue(i);
  for (long long i = 0; i < iters; ++i) {
    issue(i + kStage - 1);
    cp_async_wait<kStage - 1>();
    const int st = (int)(i & (kStage - 1));
    const uint4 u = mine[st * kThreads];
    float fy[8], g[HK];
    unpack8(*reinterpret_cast<const bf16x8*>(&u), fy);
    unpack4(mine_g[st * kThreads], g);
    body(fy, g, v0 + i * step);
  }
  cp_async_wait<0>();
}

// dz[c] = bf16( sum_k g[k] * bf16(W[k][c]) ): the head's data gradient as the unfused head_bwd stores it
__device__ __forceinline__ void head_dz(const float* g, const float (*wr)[8], float* dz) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < HK; ++k) a = fmaf(g[k], wr[k][j], a);
    dz[j] = a;
  }
#pragma unroll
  for (int j = 0; j < 8; j += 2) round_bf2(dz[j], dz[j + 1]);
}

__global__ void __launch_bounds__(kThreads, 2) inorm_lrelu_head_bwd_stats_kernel(
    const bf16* __restrict__ dl, const bf16* __restrict__ y, int ldy, const double* __restrict__ stats,
    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ w, long long V,
    float eps, float slope, double* __restrict__ bstats, float* __restrict__ dw, float* __restrict__ db) {
  extern __shared__ __align__(16) uint4 ring[];
  __shared__ float sacc[HK * HC + HK + 2 * HC];
  for (int i = threadIdx.x; i < HK * HC + HK + 2 * HC; i += kThreads) sacc[i] = 0.f;
  __syncthreads();
  const int b = blockIdx.y;
  const int cg = threadIdx.x % HCG, r = threadIdx.x / HCG;
  NormConst n;
  load_norm(stats, gamma, beta, b, cg, V, eps, n);
  float wr[HK][8];
#pragma unroll
  for (int k = 0; k < HK; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) wr[k][j] = round_bf(w[k * HC + cg * 8 + j]);
  float s1[8], s2[8], acc[HK][8], accb[HK];
#pragma unroll
  for (int k = 0; k < 8; ++k) s1[k] = s2[k] = 0.f;
#pragma unroll
  for (int k = 0; k < HK; ++k) {
    accb[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
  }
  sweep_y_dl(y + (long long)b * V * ldy, ldy, dl + (long long)b * V * HK, V, cg, r, ring,
             [&](const float* fy, const float* g, long long) {
               float dz[8], z[8];
               head_dz(g, wr, dz);
#pragma unroll
               for (int k = 0; k < 8; ++k) {
                 const float t = fmaf(fy[k], n.sc[k], n.sh[k]);
                 const float gp = dz[k] * (t > 0.f ? 1.f : slope);
                 const float xh = fmaf(fy[k], n.rs[k], n.xm[k]);
                 s1[k] += gp;
                 s2[k] = fmaf(gp, xh, s2[k]);
                 const float tr = round_bf(t);
                 z[k] = tr > 0.f ? tr : slope * tr;
               }
#pragma unroll
               for (int k = 0; k < 8; k += 2) round_bf2(z[k], z[k + 1]);
#pragma unroll
               for (int k = 0; k < HK; ++k) {
                 accb[k] += g[k];
#pragma unroll
                 for (int j = 0; j < 8; ++j) acc[k][j] = fmaf(g[k], z[j], acc[k][j]);
               }
             });
  // block reduction in shared memory, then one atomic per value and block
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    atomicAdd(&sacc[HK * HC + HK + cg * 8 + k], s1[k]);
    atomicAdd(&sacc[HK * HC + HK + HC + cg * 8 + k], s2[k]);
  }
#pragma unroll
  for (int k = 0; k < HK; ++k) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sacc[k * HC + cg * 8 + j], acc[k][j]);
    if (cg == 0) atomicAdd(&sacc[HK * HC + k], accb[k]);
  }
  __syncthreads();
  if (dw)
    for (int i = threadIdx.x; i < HK * HC; i += kThreads) atomicAdd(&dw[i], sacc[i]);
  if (db && threadIdx.x < HK) atomicAdd(&db[threadIdx.x], sacc[HK * HC + threadIdx.x]);
  if (threadIdx.x < HC) {
    atomicAdd(&bstats[((long long)b * HC + threadIdx.x) * 2 + 0], (double)sacc[HK * HC + HK + threadIdx.x]);
    atomicAdd(&bstats[((long long)b * HC + threadIdx.x) * 2 + 1], (double)sacc[HK * HC + HK + HC + threadIdx.x]);
  }
}

__global__ void __launch_bounds__(kThreads, 3) inorm_lrelu_head_bwd_apply_kernel(
    const bf16* __restrict__ dl, const bf16* __restrict__ y, int ldy, bf16* __restrict__ dy, int lddy,
    const double* __restrict__ stats, const double* __restrict__ bstats, const float* __restrict__ gamma,
    const float* __restrict__ beta, const float* __restrict__ w, int B, long long V, float eps, float slope,
    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dsum) {
  extern __shared__ __align__(16) uint4 ring[];
  __shared__ float sacc[HC];
  if (threadIdx.x < HC) sacc[threadIdx.x] = 0.f;
  const int b = blockIdx.y;
  if (blockIdx.x == 0 && blockIdx.y == 0 && (dgamma || dbeta) && threadIdx.x < HC) {
    double a = 0.0, d = 0.0;
    for (int bb = 0; bb < B; ++bb) {
      a += bstats[((long long)bb * HC + threadIdx.x) * 2 + 0];
      d += bstats[((long long)bb * HC + threadIdx.x) * 2 + 1];
    }
    if (dbeta) dbeta[threadIdx.x] = (float)a;
    if (dgamma) dgamma[threadIdx.x] = (float)d;
  }
  __syncthreads();
  const int cg = threadIdx.x % HCG, r = threadIdx.x / HCG;
  NormConst n;
  load_norm(stats, gamma, beta, b, cg, V, eps, n);
  float wr[HK][8], p1[8], p2[8], s[8];
#pragma unroll
  for (int k = 0; k < HK; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) wr[k][j] = round_bf(w[k * HC + cg * 8 + j]);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = cg * 8 + k;
    const float m1 = (float)(bstats[((long long)b * HC + c) * 2 + 0] / (double)V);
    const float m2 = (float)(bstats[((long long)b * HC + c) * 2 + 1] / (double)V);
    // dy = sc*(g' - m1 - xhat*m2), xhat = y*rs + xm  ==  fma(y, p1, fma(g', sc, p2))
    p1[k] = -n.sc[k] * n.rs[k] * m2;
    p2[k] = n.sc[k] * (-n.xm[k] * m2 - m1);
    s[k] = 0.f;
  }
  bf16* ob = dy + (long long)b * V * lddy;
  sweep_y_dl(y + (long long)b * V * ldy, ldy, dl + (long long)b * V * HK, V, cg, r, ring,
             [&](const float* fy, const float* g, long long v) {
               float dz[8], o[8];
               head_dz(g, wr, dz);
#pragma unroll
               for (int k = 0; k < 8; ++k) {
                 const float t = fmaf(fy[k], n.sc[k], n.sh[k]);
                 const float gp = dz[k] * (t > 0.f ? 1.f : slope);
                 o[k] = fmaf(fy[k], p1[k], fmaf(gp, n.sc[k], p2[k]));
               }
               const bf16x8 pk = pack8(o);
               st_stream16(ob + v * lddy + cg * 8, pk);
               if (dsum) {
                 unpack8(pk, o);
#pragma unroll
                 for (int k = 0; k < 8; ++k) s[k] += o[k];
               }
             });
  if (dsum) {
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&sacc[cg * 8 + k], s[k]);
    __syncthreads();
    if (threadIdx.x < HC) atomicAdd(&dsum[threadIdx.x], sacc[threadIdx.x]);
  }
}

template <typename K>
int wave_blocks(K kernel, size_t smem, int B, long long V, bool& attr) {
  if (!attr) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024) != cudaSuccess) {
      (void)cudaGetLastError();
      return 0;
    }
    attr = true;
  }
  int bps = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kernel, kThreads, smem) != cudaSuccess || bps < 1) {
    (void)cudaGetLastError();
    bps = 2;
  }
  long long nb = ((long long)num_sms() * bps) / B;
  const long long need = (V + kRows - 1) / kRows;
  if (nb > need) nb = need;
  return nb < 1 ? 1 : (int)nb;
}

bool shape_ok(const void* y, int ldy, int C, int K) {
  return C == HC && K == HK && ldy % 8 == 0 && ((uintptr_t)y & 15) == 0;
}

}  // namespace
}  // namespace mvd

using namespace mvd;

extern "C" {

int mvd_inorm_lrelu_head_supported(int C, int K) { return (C == HC && K == HK) ? 1 : 0; }

int mvd_inorm_lrelu_head_fwd(const void* y, int ldy, const double* stats, const float* gamma, const float* beta,
                             const float* w, const float* bias, void* logits, int B, long long V, int C, int K,
                             float eps, float slope, mvd_stream_t stream) {
  MVD_REQUIRE(y && stats && w && logits && B > 0 && V > 0, "inorm_lrelu_head_fwd: bad arguments");
  MVD_REQUIRE(shape_ok(y, ldy, C, K) && ((uintptr_t)logits & 7) == 0, "inorm_lrelu_head_fwd: built for C = 32, K = 4, aligned rows");
  const size_t smem = (size_t)kStage * kThreads * 16;
  static bool attr = false;
  const int nb = wave_blocks(inorm_lrelu_head_fwd_kernel, smem, B, V, attr);
  MVD_REQUIRE(nb > 0, "inorm_lrelu_head_fwd: cannot configure the kernel");
  inorm_lrelu_head_fwd_kernel<<<dim3((unsigned)nb, B), kThreads, smem, (cudaStream_t)stream>>>(
      (const bf16*)y, ldy, stats, gamma, beta, w, bias, (bf16*)logits, V, eps, slope);
  MVD_LAUNCH_CHECK("inorm_lrelu_head_fwd");
  return MVD_OK;
}

int mvd_inorm_lrelu_head_bwd_stats(const void* dlogits, const void* y, int ldy, const double* stats, const float* gamma,
                                   const float* beta, const float* w, int B, long long V, int C, int K, float eps,
                                   float slope, double* bstats, float* dw, float* dbias, mvd_stream_t stream) {
  MVD_REQUIRE(dlogits && y && stats && w && bstats && B > 0 && V > 0, "inorm_lrelu_head_bwd_stats: bad arguments");
  MVD_REQUIRE(shape_ok(y, ldy, C, K) && ((uintptr_t)dlogits & 7) == 0, "inorm_lrelu_head_bwd_stats: built for C = 32, K = 4");
  const size_t smem = (size_t)kStage * kThreads * (16 + 8);
  static bool attr = false;
  const int nb = wave_blocks(inorm_lrelu_head_bwd_stats_kernel, smem, B, V, attr);
  MVD_REQUIRE(nb > 0, "inorm_lrelu_head_bwd_stats: cannot configure the kernel");
  inorm_lrelu_head_bwd_stats_kernel<<<dim3((unsigned)nb, B), kThreads, smem, (cudaStream_t)stream>>>(
      (const bf16*)dlogits, (const bf16*)y, ldy, stats, gamma, beta, w, V, eps, slope, bstats, dw, dbias);
  MVD_LAUNCH_CHECK("inorm_lrelu_head_bwd_stats");
  return MVD_OK;
}

int mvd_inorm_lrelu_head_bwd_apply(const void* dlogits, const void* y, int ldy, void* dy, int lddy, const double* stats,
                                   const double* bstats, const float* gamma, const float* beta, const float* w, int B,
                                   long long V, int C, int K, float eps, float slope, float* dgamma, float* dbeta,
                                   float* dsum, mvd_stream_t stream) {
  MVD_REQUIRE(dlogits && y && dy && stats && bstats && w && B > 0 && V > 0, "inorm_lrelu_head_bwd_apply: bad arguments");
  MVD_REQUIRE(shape_ok(y, ldy, C, K) && lddy % 8 == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)dlogits & 7) == 0,
              "inorm_lrelu_head_bwd_apply: built for C = 32, K = 4");
  const size_t smem = (size_t)kStage * kThreads * (16 + 8);
  static bool attr = false;
  const int nb = wave_blocks(inorm_lrelu_head_bwd_apply_kernel, smem, B, V, attr);
  MVD_REQUIRE(nb > 0, "inorm_lrelu_head_bwd_apply: cannot configure the kernel");
  inorm_lrelu_head_bwd_apply_kernel<<<dim3((unsigned)nb, B), kThreads, smem, (cudaStream_t)stream>>>(
      (const bf16*)dlogits, (const bf16*)y, ldy, (bf16*)dy, lddy, stats, bstats, gamma, beta, w, B, V, eps, slope, dgamma,
      dbeta, dsum);
  MVD_LAUNCH_CHECK("inorm_lrelu_head_bwd_apply");
  return MVD_OK;
}

}  // extern "C"
