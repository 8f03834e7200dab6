// norm_act.cu -- layout conversion at the module edge, InstanceNorm3d(affine)+LeakyReLU forward/backward,
// bf16 accumulate.  All kernels are HBM-bound streaming kernels: 16-byte vector accesses along the contiguous
// channel axis, grids sized in multiples of the SM count, fp32 math, double accumulation of the per-(b,c) sums.
//
// Reference ops replaced: nn.InstanceNorm3d(eps=1e-5, affine=True) + nn.LeakyReLU(inplace=True)
// (nnunetv2/utilities/get_network_from_plans.py:41-44) and their autograd.
#include <stdlib.h>
#include "common.cuh"

namespace mvd {

// ------------------------------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------------------------------
template <int C>
__global__ void ncdhw_to_ndhwc_small_kernel(const float* __restrict__ src, long long sb, bf16* __restrict__ dst,
                                            long long V, int ld) {
  const int b = blockIdx.y;
  const float* s = src + (long long)b * sb;
  bf16* d = dst + (long long)b * V * ld;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < C; ++c) d[v * ld + c] = f2bf(__ldg(s + (long long)c * V + v));
  }
}

__global__ void ncdhw_to_ndhwc_generic_kernel(const float* __restrict__ src, long long sb, bf16* __restrict__ dst, int C,
                                              long long V, int ld) {
  const int b = blockIdx.y;
  const float* s = src + (long long)b * sb;
  bf16* d = dst + (long long)b * V * ld;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long long)gridDim.x * blockDim.x)
    for (int c = 0; c < C; ++c) d[v * ld + c] = f2bf(__ldg(s + (long long)c * V + v));
}

__global__ void ndhwc_to_ncdhw_kernel(const bf16* __restrict__ src, int ld, float* __restrict__ dst, int C,
                                      long long V) {
  const int b = blockIdx.y;
  const bf16* s = src + (long long)b * V * ld;
  float* d = dst + (long long)b * C * V;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (long long)gridDim.x * blockDim.x)
    for (int c = 0; c < C; ++c) d[(long long)c * V + v] = bf2f(s[v * ld + c]);
}

// ------------------------------------------------------------------------------------------------------------
// InstanceNorm statistics: per (b,c) sum and sum of squares over V voxels.
// Thread layout: CG = C/8 channel groups (one 16-byte vector each); a block of 256 threads covers 256/CG voxel rows
// per step.  Per-thread fp32 partials over a bounded run (<= 4096 rows), block tree in shared memory, then one double
// atomicAdd per (channel, block) -- contention is B*C addresses x gridDim.x adds, negligible.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bf16x8 ld16(const bf16* p) { return ldg16_pinned(p); }
constexpr int kStatThreads = 256;
constexpr int kStage1 = 8;   // ring depth of the one-input streaming reductions (32 KB per block)

// Sweep direction.  Every streaming kernel walks the volume in BANDS (all blocks side by side, band after band).
// Experiment (round 1): sweeping a tensor the previous kernel has just written BACK TO FRONT (rev = 1), so that the part
// still resident in the 126 MB L2 is met first, was measured in the cfg-2 step and changed nothing (bwd_apply 0.973 vs
// 0.976 ms, bwd_stats 0.818 vs 0.826 ms, fwd 0.663 vs 0.648 ms per step) -- it stays available behind MVD_SWEEP_REV=1.
__device__ __forceinline__ long long sweep_row(long long v, long long V, int rev) { return rev ? (V - 1 - v) : v; }

__global__ void __launch_bounds__(kStatThreads) inorm_stats_kernel(const bf16* __restrict__ y, int ld, long long V,
                                                                   int C, double* __restrict__ stats,
                                                                   int rev) {
  extern __shared__ __align__(16) float sm[];  // staging ring, then [rows][CG*8][2]
  const int b = blockIdx.y;
  const int CG = C >> 3;
  const int rows = kStatThreads / CG;  // voxel rows handled in parallel
  const int tid = threadIdx.x;
  const int cg = tid % CG, r = tid / CG;
  const bf16* base = y + (long long)b * V * ld;
  const long long step = (long long)gridDim.x * rows;   // band sweep: iteration i covers rows [i*step, (i+1)*step)
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  if (r < rows) {
    // cp.async-staged band sweep (private ring of kStage1 slots per thread: bytes in flight cost no registers, and the
    // assembler cannot sink an async copy next to its use as it does with batches of plain loads)
    const long long v0 = (long long)blockIdx.x * rows + r;
    const long long iters = v0 < V ? (V - v0 + step - 1) / step : 0;
    const bf16x8* ring = reinterpret_cast<const bf16x8*>(sm) + tid;
    const uint32_t ring_u = (uint32_t)__cvta_generic_to_shared(ring);
    auto issue = [&](long long i) {
      if (i < iters)
        cp_async16(ring_u + (uint32_t)(i & (kStage1 - 1)) * (kStatThreads * 16),
                   base + sweep_row(v0 + i * step, V, rev) * ld + cg * 8);
      cp_async_commit();
    };
    for (int i = 0; i < kStage1 - 1; ++i) issue(i);
    for (long long i = 0; i < iters; ++i) {
      issue(i + kStage1 - 1);
      cp_async_wait<kStage1 - 1>();
      const bf16x8 p = ring[(int)(i & (kStage1 - 1)) * kStatThreads];
      float f[8];
      unpack8(p, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s[k] += f[k];
        q[k] = fmaf(f[k], f[k], q[k]);
      }
    }
    cp_async_wait<0>();
  }
  __syncthreads();   // the reduction buffers below alias the staging ring
  // reduce across the `rows` threads that share a channel group
  float* ss = sm;                       // [rows][C]
  float* qq = sm + rows * C;            // [rows][C]
  if (r < rows) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      ss[r * C + cg * 8 + i] = s[i];
      qq[r * C + cg * 8 + i] = q[i];
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += kStatThreads) {
    double a = 0.0, d = 0.0;
    for (int rr = 0; rr < rows; ++rr) {
      a += (double)ss[rr * C + c];
      d += (double)qq[rr * C + c];
    }
    atomicAdd(&stats[((long long)b * C + c) * 2 + 0], a);
    atomicAdd(&stats[((long long)b * C + c) * 2 + 1], d);
  }
}

// ------------------------------------------------------------------------------------------------------------
// Launch shape of the read+write kernels (scripts/probes/stream_probe.cu on B200, 2 x 128^3 x 32 bf16): 4 resident blocks
// of 256 threads per SM with 4 x 16 B loads in flight per thread and streaming (evict-first) stores reach 6.08 TB/s
// (93 % of the copy rate); 6 blocks/SM with default stores -- the previous shape -- 4.74 TB/s.
__device__ __forceinline__ void st_stream(bf16* p, const bf16x8& v) {
  const uint4 u = *reinterpret_cast<const uint4*>(&v);
  asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
}

// The two-input backward kernels stage their loads through shared memory with cp.async: every thread owns a private
// ring of kStage 16-byte slots per input, so the bytes in flight cost no registers (the direct-load form needs 80+
// registers for 2 x 2 loads in flight, which caps it at 3 blocks/SM and ~49 KB in flight per SM -- 81 % of the copy
// rate in the probe; the staged form reaches 100 %).  A thread only ever reads back its own slots: no block barrier.
constexpr int kStage = 4;
constexpr size_t kRingBytes = (size_t)2 * kStage * kStatThreads * 16;   // two inputs

// forward apply:  t = bf16(gamma*(y-mean)*rstd + beta);  z = t > 0 ? t : bf16(slope*t)
// (two roundings, as the reference's bf16 InstanceNorm followed by an in-place bf16 LeakyReLU produces)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_scale_shift(const double* __restrict__ stats, const float* __restrict__ gamma,
                                                 const float* __restrict__ beta, int b, int C, long long V, float eps,
                                                 float* sc, float* sh, float* mean_out, float* rstd_out) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double s1 = stats[((long long)b * C + c) * 2 + 0];
    double s2 = stats[((long long)b * C + c) * 2 + 1];
    double m = s1 / (double)V;
    double var = s2 / (double)V - m * m;
    if (var < 0.0) var = 0.0;
    float rstd = (float)(1.0 / sqrt(var + (double)eps));
    float g = gamma ? gamma[c] : 1.f;
    float be = beta ? beta[c] : 0.f;
    sc[c] = g * rstd;
    sh[c] = be - (float)m * g * rstd;
    if (mean_out) mean_out[c] = (float)m;
    if (rstd_out) rstd_out[c] = rstd;
  }
}

__global__ void __launch_bounds__(256) inorm_lrelu_fwd_kernel(const bf16* __restrict__ y, int ldy,
                                                              bf16* __restrict__ z, int ldz,
                                                              const double* __restrict__ stats,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, long long V, int C,
                                                              float eps, float slope, int rev) {
  extern __shared__ float sm[];
  float* sc = sm;
  float* sh = sm + C;
  const int b = blockIdx.y;
  load_scale_shift(stats, gamma, beta, b, C, V, eps, sc, sh, nullptr, nullptr);
  __syncthreads();
  const int CG = C >> 3;
  const int rows = 256 / CG;
  const int cg = threadIdx.x % CG, r = threadIdx.x / CG;
  const bf16* yb = y + (long long)b * V * ldy;
  bf16* zb = z + (long long)b * V * ldz;
  if (r >= rows) return;
  float csc[8], csh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { csc[k] = sc[cg * 8 + k]; csh[k] = sh[cg * 8 + k]; }
  // plain pointer walk (one 64-bit add per vector): with the row index recomputed per load the scheduler sank every load
  // next to its use, i.e. ONE 16-byte load in flight per thread (SASS) -- the kernel sat at 75 % of the copy rate
  const long long step = (long long)gridDim.x * rows;
  long long v = (long long)blockIdx.x * rows + r;
  const bf16* yp = yb + v * ldy + cg * 8;
  bf16* zp = zb + v * ldz + cg * 8;
  const long long ys = step * ldy, zs = step * ldz;
  auto body = [&](const bf16x8& p, bf16* dst) {
    float f[8];
    unpack8(p, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float t = round_bf(fmaf(f[k], csc[k], csh[k]));
      f[k] = t > 0.f ? t : slope * t;
    }
    st_stream(dst, pack8(f));
  };
  for (; v + 3 * step < V; v += 4 * step, yp += 4 * ys, zp += 4 * zs) {
    bf16x8 p[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) p[u] = ld16(yp + u * ys);
#pragma unroll
    for (int u = 0; u < 4; ++u) body(p[u], zp + u * zs);
  }
  for (; v < V; v += step, yp += ys, zp += zs) body(ld16(yp), zp);
}

// ------------------------------------------------------------------------------------------------------------
// backward statistics: g' = dz * (pre > 0 ? 1 : slope);  S1 = sum g',  S2 = sum g' * xhat   per (b,c)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kStatThreads, 3) inorm_lrelu_bwd_stats_kernel(
    const bf16* __restrict__ dz, int lddz, const bf16* __restrict__ y, int ldy, const double* __restrict__ stats,
    const float* __restrict__ gamma, const float* __restrict__ beta, long long V, int C, float eps, float slope,
    double* __restrict__ bstats, int rev) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.y;
  const int CG = C >> 3;
  const int rows = kStatThreads / CG;
  float* sc = sm;            // gamma*rstd
  float* sh = sm + C;        // beta - mean*gamma*rstd
  float* mean = sm + 2 * C;
  float* rstd = sm + 3 * C;
  float* red = sm + 4 * C;   // [rows][C][2]
  load_scale_shift(stats, gamma, beta, b, C, V, eps, sc, sh, mean, rstd);
  __syncthreads();
  const int tid = threadIdx.x;
  const int cg = tid % CG, r = tid / CG;
  const bf16* yb = y + (long long)b * V * ldy;
  const bf16* gb = dz + (long long)b * V * lddz;
  const long long step = (long long)gridDim.x * rows;
  float s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
  if (r < rows) {
    float csc[8], csh[8], cme[8], crs[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = cg * 8 + k;
      csc[k] = sc[c]; csh[k] = sh[c]; cme[k] = mean[c]; crs[k] = rstd[c];
    }
    // sign(round_bf(t)) == sign(t) (bf16 keeps the fp32 exponent range), so the LeakyReLU mask needs no rounding;
    // xhat = y*rstd - mean*rstd is one FMA
    float cxm[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) cxm[k] = -cme[k] * crs[k];
    auto body = [&](const bf16x8& py, const bf16x8& pg) {
      float fy[8], fg[8];
      unpack8(py, fy);
      unpack8(pg, fg);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float t = fmaf(fy[k], csc[k], csh[k]);
        const float gp = fg[k] * (t > 0.f ? 1.f : slope);
        const float xh = fmaf(fy[k], crs[k], cxm[k]);
        s1[k] += gp;
        s2[k] = fmaf(gp, xh, s2[k]);
      }
    };
    // staged sweep: iteration i covers row v0 + i*step; kStage - 1 iterations are in flight ahead of the consumer
    const long long v0 = (long long)blockIdx.x * rows + r;
    const long long iters = v0 < V ? (V - v0 + step - 1) / step : 0;
    const bf16x8* ring = reinterpret_cast<const bf16x8*>(red) + tid;      // slot s of input j: ring[(j * kStage + s) * 256]
    const uint32_t ring_u = (uint32_t)__cvta_generic_to_shared(ring);
    auto issue = [&](long long i) {
      if (i < iters) {
        const long long w = sweep_row(v0 + i * step, V, rev);
        const uint32_t slot = ring_u + (uint32_t)(i & (kStage - 1)) * (kStatThreads * 16);
        cp_async16(slot, yb + w * ldy + cg * 8);
        cp_async16(slot + kStage * kStatThreads * 16, gb + w * lddz + cg * 8);
      }
      cp_async_commit();
    };
    for (int i = 0; i < kStage - 1; ++i) issue(i);
    for (long long i = 0; i < iters; ++i) {
      issue(i + kStage - 1);
      cp_async_wait<kStage - 1>();
      const int sl = (int)(i & (kStage - 1));
      const bf16x8 py = ring[sl * kStatThreads], pg = ring[(kStage + sl) * kStatThreads];
      body(py, pg);
    }
    cp_async_wait<0>();
  }
  __syncthreads();   // the reduction buffer below aliases the staging ring
  if (r < rows) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      red[(r * C + cg * 8 + k) * 2 + 0] = s1[k];
      red[(r * C + cg * 8 + k) * 2 + 1] = s2[k];
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += kStatThreads) {
    double a = 0.0, d = 0.0;
    for (int rr = 0; rr < rows; ++rr) {
      a += (double)red[(rr * C + c) * 2 + 0];
      d += (double)red[(rr * C + c) * 2 + 1];
    }
    atomicAdd(&bstats[((long long)b * C + c) * 2 + 0], a);
    atomicAdd(&bstats[((long long)b * C + c) * 2 + 1], d);
  }
}

// dy = gamma*rstd*(g' - S1/V - xhat*S2/V); optionally dsum[c] += sum_v dy[v][c] (the bias gradient of the conv in
// front of the norm -- analytically zero, numerically the rounding noise the reference also produces).
// Thread = (voxel row r, 8-channel group cg) with rows strided over the block's run, like the statistics kernels.
__global__ void __launch_bounds__(kStatThreads, 3) inorm_lrelu_bwd_apply_kernel(
    const bf16* __restrict__ dz, int lddz, const bf16* __restrict__ y, int ldy, bf16* __restrict__ dy, int lddy,
    const double* __restrict__ stats, const double* __restrict__ bstats, const float* __restrict__ gamma,
    const float* __restrict__ beta, int B, long long V, int C, float eps, float slope, float* __restrict__ dgamma,
    float* __restrict__ dbeta, float* __restrict__ dsum, int rev) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.y;
  float* sc = sm;
  float* sh = sm + C;
  float* mean = sm + 2 * C;
  float* rstd = sm + 3 * C;
  float* m1 = sm + 4 * C;  // S1/V
  float* m2 = sm + 5 * C;  // S2/V
  float* acc = sm + 6 * C; // per-channel sum of dy (block partial)
  load_scale_shift(stats, gamma, beta, b, C, V, eps, sc, sh, mean, rstd);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    m1[c] = (float)(bstats[((long long)b * C + c) * 2 + 0] / (double)V);
    m2[c] = (float)(bstats[((long long)b * C + c) * 2 + 1] / (double)V);
    acc[c] = 0.f;
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && (dgamma || dbeta)) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      double a = 0.0, d = 0.0;
      for (int bb = 0; bb < B; ++bb) {
        a += bstats[((long long)bb * C + c) * 2 + 0];
        d += bstats[((long long)bb * C + c) * 2 + 1];
      }
      if (dbeta) dbeta[c] = (float)a;
      if (dgamma) dgamma[c] = (float)d;
    }
  }
  __syncthreads();
  const int CG = C >> 3;
  const int rows = kStatThreads / CG;
  const int tid = threadIdx.x;
  const int cg = tid % CG, r = tid / CG;
  const bf16* yb = y + (long long)b * V * ldy;
  const bf16* gb = dz + (long long)b * V * lddz;
  bf16* ob = dy + (long long)b * V * lddy;
  const long long step = (long long)gridDim.x * rows;
  float s[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = 0.f;
  if (r < rows) {
    // dy = sc*(g' - m1 - xhat*m2) with xhat = (y - mean)*rstd  ==  fma(y, p1, fma(g', sc, p2)),
    //   p1 = -sc*rstd*m2,  p2 = sc*(mean*rstd*m2 - m1): two FMAs per element (fp32 reassociation only)
    float csc[8], csh[8], p1[8], p2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = cg * 8 + k;
      csc[k] = sc[c]; csh[k] = sh[c];
      const float rm2 = rstd[c] * m2[c];
      p1[k] = -csc[k] * rm2;
      p2[k] = csc[k] * (mean[c] * rm2 - m1[c]);
    }
    auto body = [&](const bf16x8& py, const bf16x8& pg, long long v) {
      float fy[8], fg[8], o[8];
      unpack8(py, fy);
      unpack8(pg, fg);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float t = fmaf(fy[k], csc[k], csh[k]);
        const float gp = fg[k] * (t > 0.f ? 1.f : slope);
        o[k] = fmaf(fy[k], p1[k], fmaf(gp, csc[k], p2[k]));
      }
      const bf16x8 pk = pack8(o);
      st_stream(ob + v * lddy + cg * 8, pk);
      if (dsum) {          // bias gradient = sum of the ROUNDED dy (what the reference's conv backward sums)
        unpack8(pk, o);
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] += o[k];
      }
    };
    const long long v0 = (long long)blockIdx.x * rows + r;
    const long long iters = v0 < V ? (V - v0 + step - 1) / step : 0;
    const bf16x8* ring = reinterpret_cast<const bf16x8*>(sm + 7 * C) + tid;
    const uint32_t ring_u = (uint32_t)__cvta_generic_to_shared(ring);
    auto issue = [&](long long i) {
      if (i < iters) {
        const long long w = sweep_row(v0 + i * step, V, rev);
        const uint32_t slot = ring_u + (uint32_t)(i & (kStage - 1)) * (kStatThreads * 16);
        cp_async16(slot, yb + w * ldy + cg * 8);
        cp_async16(slot + kStage * kStatThreads * 16, gb + w * lddz + cg * 8);
      }
      cp_async_commit();
    };
    for (int i = 0; i < kStage - 1; ++i) issue(i);
    for (long long i = 0; i < iters; ++i) {
      issue(i + kStage - 1);
      cp_async_wait<kStage - 1>();
      const int sl = (int)(i & (kStage - 1));
      const bf16x8 py = ring[sl * kStatThreads], pg = ring[(kStage + sl) * kStatThreads];
      body(py, pg, sweep_row(v0 + i * step, V, rev));
    }
    cp_async_wait<0>();
  }
  if (dsum) {
    if (r < rows) {
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(&acc[cg * 8 + k], s[k]);
    }
    __syncthreads();
    for (int c = tid; c < C; c += kStatThreads) atomicAdd(&dsum[c], acc[c]);
  }
}

// dst += src (bf16, pitched)
__global__ void __launch_bounds__(256) add_bf16_kernel(bf16* __restrict__ dst, int ldd, const bf16* __restrict__ src,
                                                       int lds, long long NV, int C) {
  const int CG = C >> 3;
  const long long nvec = NV * CG;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    long long v = i / CG;
    int cg = (int)(i - v * CG);
    bf16x8 a = ldg16(dst + v * ldd + cg * 8);
    bf16x8 s = ldg16(src + v * lds + cg * 8);
    float fa[8], fs[8];
    unpack8(a, fa);
    unpack8(s, fs);
#pragma unroll
    for (int k = 0; k < 8; ++k) fa[k] += fs[k];
    stg16(dst + v * ldd + cg * 8, pack8(fa));
  }
}

// Grid for a streaming kernel over V voxel rows of B samples: exactly ONE resident wave (SM count x blocks that fit per
// SM), every block with the same share of rows.  (ncu, round 1: 1024 blocks on 148 x 5 slots = 1.4 waves left half the
// machine idle in the second wave -> 50 % of the HBM roofline.)  Returns blocks per sample; *rpb = rows per block, a
// multiple of `gran`.
template <typename K>
static long long one_wave(K kernel, int threads, size_t smem, int B, long long V, long long gran, long long* rpb,
                          int max_bps = 0) {
  int bps = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kernel, threads, smem) != cudaSuccess || bps < 1) {
    (void)cudaGetLastError();
    bps = 2;
  }
  if (max_bps > 0 && bps > max_bps) bps = max_bps;
  long long nblk = ((long long)num_sms() * bps) / B;
  if (nblk < 1) nblk = 1;
  long long r = (V + nblk - 1) / nblk;
  r = (r + gran - 1) / gran * gran;
  if (r < gran) r = gran;
  *rpb = r;
  return (V + r - 1) / r;
}

// MVD_SWEEP_REV=1 turns the back-to-front sweeps on (experiment switch; default: every kernel sweeps front to back)
static int sweep_rev(int want) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MVD_SWEEP_REV");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on ? want : 0;
}

// dynamic shared memory above 48 KB needs an opt-in, once per kernel
template <typename K>
static cudaError_t allow_big_smem(K kernel, int idx) {
  static bool done[4] = {false, false, false, false};
  if (done[idx]) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
  if (e == cudaSuccess) done[idx] = true;
  return e;
}

static bool vec_ok(const void* p, int ld, int C) {
  return (C % 8 == 0) && (ld % 8 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15) == 0);
}

}  // namespace mvd

using namespace mvd;

extern "C" {

int mvd_ncdhw_f32_to_ndhwc_bf16(const float* src, long long src_batch_stride, void* dst, int B, int C, long long V,
                                int ld_dst,
                                mvd_stream_t stream) {
  MVD_REQUIRE(src && dst && B > 0 && C > 0 && V > 0 && ld_dst >= C, "ncdhw_to_ndhwc: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(grid_for(V, 256, num_sms() * 8), B);
  bf16* d = (bf16*)dst;
  const long long sb = src_batch_stride > 0 ? src_batch_stride : (long long)C * V;
  if (C == 1) ncdhw_to_ndhwc_small_kernel<1><<<grid, 256, 0, st>>>(src, sb, d, V, ld_dst);
  else if (C == 2) ncdhw_to_ndhwc_small_kernel<2><<<grid, 256, 0, st>>>(src, sb, d, V, ld_dst);
  else if (C == 4) ncdhw_to_ndhwc_small_kernel<4><<<grid, 256, 0, st>>>(src, sb, d, V, ld_dst);
  else ncdhw_to_ndhwc_generic_kernel<<<grid, 256, 0, st>>>(src, sb, d, C, V, ld_dst);
  MVD_LAUNCH_CHECK("ncdhw_to_ndhwc");
  return MVD_OK;
}

int mvd_ndhwc_bf16_to_ncdhw_f32(const void* src, int ld_src, float* dst, int B, int C, long long V,
                                mvd_stream_t stream) {
  MVD_REQUIRE(src && dst && B > 0 && C > 0 && V > 0 && ld_src >= C, "ndhwc_to_ncdhw: bad arguments");
  dim3 grid(grid_for(V, 256, num_sms() * 8), B);
  ndhwc_to_ncdhw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)src, ld_src, dst, C, V);
  MVD_LAUNCH_CHECK("ndhwc_to_ncdhw");
  return MVD_OK;
}

int mvd_inorm_stats(const void* y, int ldy, int B, long long V, int C, double* stats, mvd_stream_t stream) {
  MVD_REQUIRE(y && stats && B > 0 && V > 0, "inorm_stats: bad arguments");
  MVD_REQUIRE(vec_ok(y, ldy, C) && C <= 2048 && (kStatThreads / (C / 8)) >= 1,
              "inorm_stats: need C %% 8 == 0, ld %% 8 == 0, 16B-aligned pointer, C <= 2048 (C=%d ld=%d)", C, ldy);
  const int rows = kStatThreads / (C / 8);
  size_t smem = (size_t)rows * C * 2 * sizeof(float);
  if (smem < (size_t)kStage1 * kStatThreads * 16) smem = (size_t)kStage1 * kStatThreads * 16;
  long long rpb;
  const long long nblk = one_wave(inorm_stats_kernel, kStatThreads, smem, B, V, (long long)rows * 4, &rpb);
  dim3 grid((unsigned)nblk, B);
  inorm_stats_kernel<<<grid, kStatThreads, smem, (cudaStream_t)stream>>>((const bf16*)y, ldy, V, C, stats, sweep_rev(1));
  MVD_LAUNCH_CHECK("inorm_stats");
  return MVD_OK;
}

int mvd_inorm_lrelu_fwd(const void* y, int ldy, void* z, int ldz, const double* stats, const float* gamma,
                        const float* beta, int B, long long V, int C, float eps, float slope, mvd_stream_t stream) {
  MVD_REQUIRE(y && z && stats && B > 0 && V > 0, "inorm_lrelu_fwd: bad arguments");
  MVD_REQUIRE(vec_ok(y, ldy, C) && vec_ok(z, ldz, C), "inorm_lrelu_fwd: need C %% 8 == 0 and 16B-aligned pitched rows");
  // grid-stride kernel: one resident wave; small volumes get fewer blocks (>= 4 rows per thread slot)
  long long rpb;
  const int frows = 256 / ((C / 8) > 256 ? 256 : (C / 8));
  const long long nblk = one_wave(inorm_lrelu_fwd_kernel, 256, 2 * C * sizeof(float), B, V, (long long)frows * 4, &rpb, 4);
  dim3 grid((unsigned)nblk, B);
  inorm_lrelu_fwd_kernel<<<grid, 256, 2 * C * sizeof(float), (cudaStream_t)stream>>>(
      (const bf16*)y, ldy, (bf16*)z, ldz, stats, gamma, beta, V, C, eps, slope, sweep_rev(1));
  MVD_LAUNCH_CHECK("inorm_lrelu_fwd");
  return MVD_OK;
}

int mvd_inorm_lrelu_bwd_stats(const void* dz, int lddz, const void* y, int ldy, const double* stats,
                              const float* gamma, const float* beta, int B, long long V, int C, float eps,
                              float slope, double* bstats, mvd_stream_t stream) {
  MVD_REQUIRE(dz && y && stats && bstats && B > 0 && V > 0, "inorm_lrelu_bwd_stats: bad arguments");
  MVD_REQUIRE(vec_ok(y, ldy, C) && vec_ok(dz, lddz, C) && C <= 2048, "inorm_lrelu_bwd_stats: alignment/C");
  const int rows = kStatThreads / (C / 8);
  size_t red_bytes = (size_t)rows * C * 2 * sizeof(float);
  size_t smem = (size_t)4 * C * sizeof(float) + (red_bytes > kRingBytes ? red_bytes : kRingBytes);
  MVD_CUDA(allow_big_smem(inorm_lrelu_bwd_stats_kernel, 0));
  long long rpb;
  const long long nblk = one_wave(inorm_lrelu_bwd_stats_kernel, kStatThreads, smem, B, V, (long long)rows * 2, &rpb);
  dim3 grid((unsigned)nblk, B);
  inorm_lrelu_bwd_stats_kernel<<<grid, kStatThreads, smem, (cudaStream_t)stream>>>(
      (const bf16*)dz, lddz, (const bf16*)y, ldy, stats, gamma, beta, V, C, eps, slope, bstats, sweep_rev(1));
  MVD_LAUNCH_CHECK("inorm_lrelu_bwd_stats");
  return MVD_OK;
}

int mvd_inorm_lrelu_bwd_apply(const void* dz, int lddz, const void* y, int ldy, void* dy, int lddy,
                              const double* stats, const double* bstats, const float* gamma, const float* beta,
                              int B, long long V, int C, float eps, float slope, float* dgamma, float* dbeta,
                              float* dsum, mvd_stream_t stream) {
  MVD_REQUIRE(dz && y && dy && stats && bstats && B > 0 && V > 0, "inorm_lrelu_bwd_apply: bad arguments");
  MVD_REQUIRE(vec_ok(y, ldy, C) && vec_ok(dz, lddz, C) && vec_ok(dy, lddy, C) && C <= 2048,
              "inorm_lrelu_bwd_apply: alignment/C");
  const int rows = kStatThreads / (C / 8);
  long long rpb;
  const size_t smem = (size_t)7 * C * sizeof(float) + kRingBytes;
  MVD_CUDA(allow_big_smem(inorm_lrelu_bwd_apply_kernel, 1));
  const long long nblk = one_wave(inorm_lrelu_bwd_apply_kernel, kStatThreads, smem, B, V, (long long)rows * 2, &rpb);
  dim3 grid((unsigned)nblk, B);
  inorm_lrelu_bwd_apply_kernel<<<grid, kStatThreads, smem, (cudaStream_t)stream>>>(
      (const bf16*)dz, lddz, (const bf16*)y, ldy, (bf16*)dy, lddy, stats, bstats, gamma, beta, B, V, C, eps, slope,
      dgamma, dbeta, dsum, sweep_rev(0));
  MVD_LAUNCH_CHECK("inorm_lrelu_bwd_apply");
  return MVD_OK;
}

int mvd_add_bf16(void* dst, int ldd, const void* src, int lds, long long NV, int C, mvd_stream_t stream) {
  MVD_REQUIRE(dst && src && NV > 0, "add_bf16: bad arguments");
  MVD_REQUIRE(vec_ok(dst, ldd, C) && vec_ok(src, lds, C), "add_bf16: alignment/C");
  int grid = grid_for(NV * (C / 8), 256 * 4, num_sms() * 8);
  add_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((bf16*)dst, ldd, (const bf16*)src, lds, NV, C);
  MVD_LAUNCH_CHECK("add_bf16");
  return MVD_OK;
}

}  // extern "C"
