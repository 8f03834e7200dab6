// head.cu -- the decoder's 1x1x1 segmentation heads (nn.Conv3d(C, num_classes, 1), UNetDecoder.py:67-70).
// K = num_classes <= 8 outputs per voxel: HBM-bound, CUDA cores.  One thread per voxel, the [K][C] weights
// (rounded to bf16, as autocast feeds them to the reference conv) broadcast from shared memory.
#include "common.cuh"

namespace mvd {
// dz store of the head backward; `accumulate`: add to what another consumer of the same activation already wrote there
// (the up-convolution that also reads a decoder stage output: folds autograd's two-consumer sum into this kernel)
__device__ __forceinline__ void store_dz(bf16* p, const float* o, bool accumulate) {
  if (accumulate) {
    float old[8], s[8];
    unpack8(ldg16(p), old);
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = o[j] + old[j];
    stg16(p, pack8(s));
  } else {
    stg16(p, pack8(o));
  }
}


constexpr int kMaxHeadK = 8;

template <int K>
__global__ void __launch_bounds__(256) head_fwd_kernel(const bf16* __restrict__ z, int ldz,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       bf16* __restrict__ out, int ldl, long long NV, int C) {
  extern __shared__ float sw[];  // [K][C] + [K]
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) sw[i] = round_bf(w[i]);
  for (int i = threadIdx.x; i < K; i += blockDim.x) sw[K * C + i] = bias ? round_bf(bias[i]) : 0.f;
  __syncthreads();
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < NV; v += (long long)gridDim.x * blockDim.x) {
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.f;
    const bf16* zp = z + v * ldz;
    for (int c0 = 0; c0 < C; c0 += 8) {
      bf16x8 p = ldg16(zp + c0);
      float f[8];
      unpack8(p, f);
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[k] = fmaf(f[j], sw[k * C + c0 + j], acc[k]);
    }
    bf16* op = out + v * ldl;
#pragma unroll
    for (int k = 0; k < K; ++k) op[k] = f2bf(acc[k] + sw[K * C + k]);
  }
}

// dz[v][c] = sum_k dl[v][k] * w[k][c]
template <int K>
__global__ void __launch_bounds__(256) head_bwd_data_kernel(const bf16* __restrict__ dl, int ldl,
                                                            const float* __restrict__ w, bf16* __restrict__ dz,
                                                            int lddz, long long NV, int C, bool acc_dz) {
  extern __shared__ float sw[];
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) sw[i] = round_bf(w[i]);
  __syncthreads();
  const int CG = C >> 3;
  const long long nvec = NV * CG;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    long long v = i / CG;
    int cg = (int)(i - v * CG);
    float g[K];
#pragma unroll
    for (int k = 0; k < K; ++k) g[k] = bf2f(dl[v * ldl + k]);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) a = fmaf(g[k], sw[k * C + cg * 8 + j], a);
      o[j] = a;
    }
    store_dz(dz + v * lddz + cg * 8, o, acc_dz);
  }
}

// dw[k][c] += sum_v dl[v][k] * z[v][c] ; dbias[k] += sum_v dl[v][k]
// Thread = (voxel row r, 8-channel group cg): 16-byte loads of z, K*8 fp32 accumulators, rows strided over the block's
// voxel run; partials meet in shared memory (float atomics, once per thread), one global atomicAdd per (k,c) per block.
template <int K>
__global__ void __launch_bounds__(256) head_bwd_weight_kernel(const bf16* __restrict__ dl, int ldl,
                                                              const bf16* __restrict__ z, int ldz,
                                                              float* __restrict__ dw, float* __restrict__ dbias,
                                                              long long NV, int C, long long rows_per_block) {
  extern __shared__ float sacc[];  // [K][C] + [K]
  for (int i = threadIdx.x; i < K * C + K; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  long long v0 = (long long)blockIdx.x * rows_per_block;
  long long v1 = v0 + rows_per_block;
  if (v1 > NV) v1 = NV;
  const int CG = C >> 3;
  const int rows = 256 / CG;
  const int tid = threadIdx.x;
  const int cg = tid % CG, r = tid / CG;
  float acc[K][8], accb[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    accb[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
  }
  if (r < rows) {
    for (long long v = v0 + r; v < v1; v += rows) {
      bf16x8 p = ldg16(z + v * ldz + cg * 8);
      float f[8];
      unpack8(p, f);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float g = bf2f(dl[v * ldl + k]);
        if (cg == 0) accb[k] += g;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[k][j] = fmaf(g, f[j], acc[k][j]);
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&sacc[k * C + cg * 8 + j], acc[k][j]);
      if (cg == 0) atomicAdd(&sacc[K * C + k], accb[k]);
    }
  }
  __syncthreads();
  for (int i = tid; i < K * C; i += 256) atomicAdd(&dw[i], sacc[i]);
  if (dbias && tid < K) atomicAdd(&dbias[tid], sacc[K * C + tid]);
}

// ------------------------------------------------------------------------------------------------------------------
// Fast paths for C = 8*CG with CG a power of two <= 32 (C = 32 .. 256: every head but the 320-channel 8^3 one).
// Thread = (voxel row r, 8-channel group cg): one coalesced 16-byte load of z per row, the thread's K x 8 weights live
// in registers (no shared-memory reads in the loop), 4 rows in flight per thread, logits / gradient rows moved with
// one 8-byte access when K == 4.  Grids are one resident wave.
// ------------------------------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void load_row_k(const bf16* p, bool vec, float* g) {
  if (K == 4 && vec) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    g[0] = __uint_as_float(u.x << 16); g[1] = __uint_as_float(u.x & 0xffff0000u);
    g[2] = __uint_as_float(u.y << 16); g[3] = __uint_as_float(u.y & 0xffff0000u);
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) g[k] = bf2f(p[k]);
  }
}

template <int K, int CG>
__global__ void __launch_bounds__(256) head_fwd_cg_kernel(const bf16* __restrict__ z, int ldz,
                                                          const float* __restrict__ w, const float* __restrict__ bias,
                                                          bf16* __restrict__ out, int ldl, long long NV, int vec) {
  constexpr int C = CG * 8, ROWS = 256 / CG, U = 4;
  const int cg = threadIdx.x % CG, r = threadIdx.x / CG;
  float wr[K][8], bz[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    bz[k] = bias ? round_bf(bias[k]) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) wr[k][j] = round_bf(w[k * C + cg * 8 + j]);
  }
  const long long stride = (long long)gridDim.x * ROWS;
  for (long long base = (long long)blockIdx.x * ROWS; base < NV; base += U * stride) {
    bf16x8 p[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long v = base + u * stride + r;
      ok[u] = v < NV;
      if (ok[u]) p[u] = ldg16(z + v * ldz + cg * 8);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float f[8], acc[K];
      if (ok[u]) unpack8(p[u], f);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) a = fmaf(f[j], wr[k][j], a);
#pragma unroll
        for (int off = CG / 2; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
        acc[k] = a + bz[k];
      }
      if (ok[u] && cg == 0) {
        bf16* op = out + (base + u * stride + r) * ldl;
        if (K == 4 && vec) {
          __nv_bfloat162 a = __floats2bfloat162_rn(acc[0], acc[1]), b = __floats2bfloat162_rn(acc[2], acc[3]);
          *reinterpret_cast<uint2*>(op) = make_uint2(*reinterpret_cast<unsigned*>(&a), *reinterpret_cast<unsigned*>(&b));
        } else {
#pragma unroll
          for (int k = 0; k < K; ++k) op[k] = f2bf(acc[k]);
        }
      }
    }
  }
}

// dz[v][c] = sum_k dl[v][k] * w[k][c]  and  dw[k][c] += sum_v dl[v][k] * z[v][c], dbias[k] += sum_v dl[v][k]
// in ONE pass over the rows (dz may be null).
template <int K, int CG>
__global__ void __launch_bounds__(256) head_bwd_cg_kernel(const bf16* __restrict__ dl, int ldl,
                                                          const bf16* __restrict__ z, int ldz,
                                                          const float* __restrict__ w, bf16* __restrict__ dz, int lddz,
                                                          float* __restrict__ dw, float* __restrict__ dbias,
                                                          long long NV, int vec, bool acc_dz) {
  constexpr int C = CG * 8, ROWS = 256 / CG, U = 4;
  __shared__ float sacc[K * C + K];
  for (int i = threadIdx.x; i < K * C + K; i += 256) sacc[i] = 0.f;
  __syncthreads();
  const int cg = threadIdx.x % CG, r = threadIdx.x / CG;
  float wr[K][8], acc[K][8], accb[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    accb[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { wr[k][j] = round_bf(w[k * C + cg * 8 + j]); acc[k][j] = 0.f; }
  }
  const long long stride = (long long)gridDim.x * ROWS;
  for (long long base = (long long)blockIdx.x * ROWS + r; base < NV; base += U * stride) {
    bf16x8 p[U];
    float g[U][K];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long v = base + u * stride;
      if (v < NV) {
        if (dw) p[u] = ldg16(z + v * ldz + cg * 8);
        load_row_k<K>(dl + v * ldl, vec != 0, g[u]);
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) g[u][k] = 0.f;
        p[u] = bf16x8{};
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long v = base + u * stride;
      if (v >= NV) continue;
      if (dz) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a = 0.f;
#pragma unroll
          for (int k = 0; k < K; ++k) a = fmaf(g[u][k], wr[k][j], a);
          o[j] = a;
        }
        store_dz(dz + v * lddz + cg * 8, o, acc_dz);
      }
      if (dw) {
        float f[8];
        unpack8(p[u], f);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          accb[k] += g[u][k];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[k][j] = fmaf(g[u][k], f[j], acc[k][j]);
        }
      }
    }
  }
  if (dw) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&sacc[k * C + cg * 8 + j], acc[k][j]);
      if (cg == 0) atomicAdd(&sacc[K * C + k], accb[k]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * C; i += 256) atomicAdd(&dw[i], sacc[i]);
    if (dbias && threadIdx.x < K) atomicAdd(&dbias[threadIdx.x], sacc[K * C + threadIdx.x]);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// cp.async-staged forms of the two kernels above for K = 4 classes with 8-byte logit rows (the production heads).
// The direct-load forms need 94 / 128 registers for 4 rows in flight (2 blocks/SM, ~49 KB in flight per SM: half the HBM
// copy rate in the step profile); here the rows wait in a private ring of kHeadStage slots per thread.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kHeadStage = 8;

__device__ __forceinline__ void cp_async8(uint32_t smem_addr, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(g) : "memory");
}

template <int CG>
__global__ void __launch_bounds__(256, 3) head_fwd_cg4_staged_kernel(const bf16* __restrict__ z, int ldz,
                                                                     const float* __restrict__ w,
                                                                     const float* __restrict__ bias,
                                                                     bf16* __restrict__ out, int ldl, long long NV) {
  constexpr int K = 4, C = CG * 8, ROWS = 256 / CG;
  extern __shared__ __align__(16) uint4 hring[];
  const int cg = threadIdx.x % CG, r = threadIdx.x / CG;
  float wr[K][8], bz[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    bz[k] = bias ? round_bf(bias[k]) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) wr[k][j] = round_bf(w[k * C + cg * 8 + j]);
  }
  const long long step = (long long)gridDim.x * ROWS;
  const long long v0 = (long long)blockIdx.x * ROWS + r;
  // all lanes of a warp must run the same number of steps (the class sums are reduced with warp shuffles)
  const long long vw = (long long)blockIdx.x * ROWS + (threadIdx.x & ~31) / CG;
  const long long iters = vw < NV ? (NV - vw + step - 1) / step : 0;
  const uint4* mine = hring + threadIdx.x;
  const uint32_t mine_u = (uint32_t)__cvta_generic_to_shared(mine);
  auto issue = [&](long long i) {
    const long long v = v0 + i * step;
    if (i < iters && v < NV) cp_async16(mine_u + (uint32_t)(i & (kHeadStage - 1)) * 4096, z + v * ldz + cg * 8);
    cp_async_commit();
  };
  for (int i = 0; i < kHeadStage - 1; ++i) issue(i);
  for (long long i = 0; i < iters; ++i) {
    issue(i + kHeadStage - 1);
    cp_async_wait<kHeadStage - 1>();
    const long long v = v0 + i * step;
    const bool ok = v < NV;
    float f[8];
    if (ok) {
      const uint4 u = mine[(int)(i & (kHeadStage - 1)) * 256];
      unpack8(*reinterpret_cast<const bf16x8*>(&u), f);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = 0.f;
    }
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) a = fmaf(f[j], wr[k][j], a);
#pragma unroll
      for (int off = CG / 2; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
      acc[k] = a + bz[k];
    }
    if (ok && cg == 0) {
      __nv_bfloat162 a = __floats2bfloat162_rn(acc[0], acc[1]), b = __floats2bfloat162_rn(acc[2], acc[3]);
      *reinterpret_cast<uint2*>(out + v * ldl) = make_uint2(*reinterpret_cast<unsigned*>(&a), *reinterpret_cast<unsigned*>(&b));
    }
  }
  cp_async_wait<0>();
}

template <int CG>
__global__ void __launch_bounds__(256, 2) head_bwd_cg4_staged_kernel(const bf16* __restrict__ dl, int ldl,
                                                                     const bf16* __restrict__ z, int ldz,
                                                                     const float* __restrict__ w, bf16* __restrict__ dz,
                                                                     int lddz, float* __restrict__ dw,
                                                                     float* __restrict__ dbias, long long NV,
                                                                     bool acc_dz) {
  constexpr int K = 4, C = CG * 8, ROWS = 256 / CG;
  extern __shared__ __align__(16) uint4 hring[];     // [stage][256] z vectors, then [stage][256] 8-byte logit-gradient rows
  __shared__ float sacc[K * C + K];
  for (int i = threadIdx.x; i < K * C + K; i += 256) sacc[i] = 0.f;
  __syncthreads();
  const int cg = threadIdx.x % CG, r = threadIdx.x / CG;
  float wr[K][8], acc[K][8], accb[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    accb[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { wr[k][j] = round_bf(w[k * C + cg * 8 + j]); acc[k][j] = 0.f; }
  }
  const long long step = (long long)gridDim.x * ROWS;
  const long long v0 = (long long)blockIdx.x * ROWS + r;
  const long long iters = v0 < NV ? (NV - v0 + step - 1) / step : 0;
  const uint4* mine = hring + threadIdx.x;
  const uint2* mine_g = reinterpret_cast<const uint2*>(hring + kHeadStage * 256) + threadIdx.x;
  const uint32_t mine_u = (uint32_t)__cvta_generic_to_shared(mine);
  const uint32_t mine_gu = (uint32_t)__cvta_generic_to_shared(mine_g);
  const bool need_z = dw != nullptr;
  auto issue = [&](long long i) {
    if (i < iters) {
      const long long v = v0 + i * step;
      const int st = (int)(i & (kHeadStage - 1));
      if (need_z) cp_async16(mine_u + (uint32_t)st * 4096, z + v * ldz + cg * 8);
      cp_async8(mine_gu + (uint32_t)st * 2048, dl + v * ldl);
    }
    cp_async_commit();
  };
  for (int i = 0; i < kHeadStage - 1; ++i) issue(i);
  for (long long i = 0; i < iters; ++i) {
    issue(i + kHeadStage - 1);
    cp_async_wait<kHeadStage - 1>();
    const int st = (int)(i & (kHeadStage - 1));
    const long long v = v0 + i * step;
    const uint2 gu = mine_g[st * 256];
    float g[K];
    g[0] = __uint_as_float(gu.x << 16); g[1] = __uint_as_float(gu.x & 0xffff0000u);
    g[2] = __uint_as_float(gu.y << 16); g[3] = __uint_as_float(gu.y & 0xffff0000u);
    if (dz) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) a = fmaf(g[k], wr[k][j], a);
        o[j] = a;
      }
      store_dz(dz + v * lddz + cg * 8, o, acc_dz);
    }
    if (need_z) {
      const uint4 u = mine[st * 256];
      float f[8];
      unpack8(*reinterpret_cast<const bf16x8*>(&u), f);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        accb[k] += g[k];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[k][j] = fmaf(g[k], f[j], acc[k][j]);
      }
    }
  }
  cp_async_wait<0>();
  if (dw) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&sacc[k * C + cg * 8 + j], acc[k][j]);
      if (cg == 0) atomicAdd(&sacc[K * C + k], accb[k]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * C; i += 256) atomicAdd(&dw[i], sacc[i]);
    if (dbias && threadIdx.x < K) atomicAdd(&dbias[threadIdx.x], sacc[K * C + threadIdx.x]);
  }
}

template <typename Kern>
static int head_staged_grid(Kern kern, size_t smem, long long NV, int rows) {
  int bps = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, 256, smem) != cudaSuccess || bps < 1) {
    (void)cudaGetLastError();
    bps = 2;
  }
  long long g = (long long)num_sms() * bps;
  const long long need = (NV + rows - 1) / rows;
  if (g > need) g = need;
  return (int)(g < 1 ? 1 : g);
}

template <typename Kern>
static int head_wave_grid(Kern kern, long long NV, int rows) {
  int bps = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, 256, 0) != cudaSuccess || bps < 1) {
    (void)cudaGetLastError();
    bps = 2;
  }
  long long g = (long long)num_sms() * bps;
  const long long need = (NV + rows - 1) / rows;
  if (g > need) g = need;
  return (int)(g < 1 ? 1 : g);
}

template <int K, int CG>
static int head_fwd_cg_launch(const bf16* z, int ldz, const float* w, const float* b, bf16* out, int ldl, long long NV,
                              cudaStream_t st) {
  const int vec = (K == 4 && ldl % 4 == 0 && ((uintptr_t)out & 7) == 0) ? 1 : 0;
  if constexpr (K == 4) {
    if (vec && ldz % 8 == 0 && ((uintptr_t)z & 15) == 0) {
      const size_t smem = (size_t)kHeadStage * 4096;
      const int grid = head_staged_grid(head_fwd_cg4_staged_kernel<CG>, smem, NV, 256 / CG);
      head_fwd_cg4_staged_kernel<CG><<<grid, 256, smem, st>>>(z, ldz, w, b, out, ldl, NV);
      MVD_LAUNCH_CHECK("head_fwd");
      return MVD_OK;
    }
  }
  const int grid = head_wave_grid(head_fwd_cg_kernel<K, CG>, NV, 256 / CG);
  head_fwd_cg_kernel<K, CG><<<grid, 256, 0, st>>>(z, ldz, w, b, out, ldl, NV, vec);
  MVD_LAUNCH_CHECK("head_fwd");
  return MVD_OK;
}

template <int K, int CG>
static int head_bwd_cg_launch(const bf16* dl, int ldl, const bf16* z, int ldz, const float* w, bf16* dz, int lddz,
                              float* dw, float* db, long long NV, bool acc_dz, cudaStream_t st) {
  const int vec = (K == 4 && ldl % 4 == 0 && ((uintptr_t)dl & 7) == 0) ? 1 : 0;
  if constexpr (K == 4) {
    if (vec && ldz % 8 == 0 && ((uintptr_t)z & 15) == 0 && (!dz || (lddz % 8 == 0 && ((uintptr_t)dz & 15) == 0))) {
      const size_t smem = (size_t)kHeadStage * (4096 + 2048);
      static bool attr_done = false;   // dynamic + static shared memory exceeds the 48 KB default
      if (!attr_done) {
        MVD_CUDA(cudaFuncSetAttribute(head_bwd_cg4_staged_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_done = true;
      }
      const int grid = head_staged_grid(head_bwd_cg4_staged_kernel<CG>, smem, NV, 256 / CG);
      head_bwd_cg4_staged_kernel<CG><<<grid, 256, smem, st>>>(dl, ldl, z, ldz, w, dz, lddz, dw, db, NV, acc_dz);
      MVD_LAUNCH_CHECK("head_bwd");
      return MVD_OK;
    }
  }
  const int grid = head_wave_grid(head_bwd_cg_kernel<K, CG>, NV, 256 / CG);
  head_bwd_cg_kernel<K, CG><<<grid, 256, 0, st>>>(dl, ldl, z, ldz, w, dz, lddz, dw, db, NV, vec, acc_dz);
  MVD_LAUNCH_CHECK("head_bwd");
  return MVD_OK;
}

template <int K>
static int head_fwd_launch(const bf16* z, int ldz, const float* w, const float* b, bf16* out, int ldl, long long NV,
                           int C, cudaStream_t st) {
  if (C == 32) return head_fwd_cg_launch<K, 4>(z, ldz, w, b, out, ldl, NV, st);
  if (C == 64) return head_fwd_cg_launch<K, 8>(z, ldz, w, b, out, ldl, NV, st);
  if (C == 128) return head_fwd_cg_launch<K, 16>(z, ldz, w, b, out, ldl, NV, st);
  if (C == 256) return head_fwd_cg_launch<K, 32>(z, ldz, w, b, out, ldl, NV, st);
  int grid = grid_for(NV, 256, num_sms() * 8);
  head_fwd_kernel<K><<<grid, 256, (K * C + K) * sizeof(float), st>>>(z, ldz, w, b, out, ldl, NV, C);
  MVD_LAUNCH_CHECK("head_fwd");
  return MVD_OK;
}

template <int K>
static int head_bwd_launch(const bf16* dl, int ldl, const bf16* z, int ldz, const float* w, bf16* dz, int lddz,
                           float* dw, float* db, long long NV, int C, bool acc_dz, cudaStream_t st) {
  if (C == 32) return head_bwd_cg_launch<K, 4>(dl, ldl, z, ldz, w, dz, lddz, dw, db, NV, acc_dz, st);
  if (C == 64) return head_bwd_cg_launch<K, 8>(dl, ldl, z, ldz, w, dz, lddz, dw, db, NV, acc_dz, st);
  if (C == 128) return head_bwd_cg_launch<K, 16>(dl, ldl, z, ldz, w, dz, lddz, dw, db, NV, acc_dz, st);
  if (C == 256) return head_bwd_cg_launch<K, 32>(dl, ldl, z, ldz, w, dz, lddz, dw, db, NV, acc_dz, st);
  if (dz) {
    int grid = grid_for(NV * (C / 8), 256 * 2, num_sms() * 8);
    head_bwd_data_kernel<K><<<grid, 256, K * C * sizeof(float), st>>>(dl, ldl, w, dz, lddz, NV, C, acc_dz);
    MVD_LAUNCH_CHECK("head_bwd_data");
  }
  if (dw) {
    long long rpb = 2048;
    long long nblk = (NV + rpb - 1) / rpb;
    while (nblk < (long long)num_sms() * 2 && rpb > 64) {
      rpb >>= 1;
      nblk = (NV + rpb - 1) / rpb;
    }
    head_bwd_weight_kernel<K><<<(unsigned)nblk, 256, (K * C + K) * sizeof(float), st>>>(dl, ldl, z, ldz, dw, db, NV, C, rpb);
    MVD_LAUNCH_CHECK("head_bwd_weight");
  }
  return MVD_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Sliding-window inference (inference/predict_from_raw_data.py:703-712): one tile's logits, weighted by the Gaussian
// importance map, are added into the full-volume accumulators
//     acc[k][z0+z][y0+y][x0+x] += pred[z][y][x][k] * g[z][y][x];   npred[z0+z][y0+y][x0+x] += g[z][y][x]
// (g = 1 when no map is given).  pred is the network's bf16 NDHWC output; acc / npred are fp32 (the reference keeps
// them in fp16: fp32 accumulation is the more accurate superset).  One thread per tile voxel, x fastest.
// ------------------------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256) sw_accumulate_kernel(const bf16* __restrict__ pred, int ldp,
                                                            const float* __restrict__ g, float scale,
                                                            float* __restrict__ acc, float* __restrict__ npred, int d,
                                                            int h, int w, int D, int H, int W, int z0, int y0, int x0,
                                                            int flip_mask) {
  const long long n = (long long)d * h * w;
  const long long vol = (long long)D * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const long long r = i / w;
    const int y = (int)(r % h), z = (int)(r / h);
    const float gw = g ? g[i] : 1.f;
    const long long o = ((long long)(z0 + z) * H + (y0 + y)) * W + (x0 + x);
    // a mirrored test-time-augmentation pass (predict_from_raw_data.py:562-589) is read back through the same flips
    const int sz = (flip_mask & 1) ? d - 1 - z : z, sy = (flip_mask & 2) ? h - 1 - y : y,
              sx = (flip_mask & 4) ? w - 1 - x : x;
    const long long src = ((long long)sz * h + sy) * w + sx;
    float v[K];
    load_row_k<K>(pred + src * ldp, K == 4 && (ldp % 4 == 0) && ((reinterpret_cast<uintptr_t>(pred) & 7) == 0), v);
#pragma unroll
    for (int k = 0; k < K; ++k) acc[(long long)k * vol + o] += v[k] * scale * gw;
    if (npred) npred[o] += gw;
  }
}

__global__ void __launch_bounds__(256) sw_finalize_kernel(float* __restrict__ acc, const float* __restrict__ npred,
                                                          int K, long long vol) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < vol; i += (long long)gridDim.x * blockDim.x) {
    const float inv = 1.f / npred[i];
    for (int k = 0; k < K; ++k) acc[(long long)k * vol + i] *= inv;
  }
}

}  // namespace mvd

using namespace mvd;

#define HEAD_DISPATCH(K, CALL)            \
  switch (K) {                            \
    case 1: return CALL(1);               \
    case 2: return CALL(2);               \
    case 3: return CALL(3);               \
    case 4: return CALL(4);               \
    case 5: return CALL(5);               \
    case 6: return CALL(6);               \
    case 7: return CALL(7);               \
    case 8: return CALL(8);               \
    default: break;                       \
  }

extern "C" {

int mvd_head_fwd(const void* z, int ldz, const float* w, const float* bias, void* logits, int ldl, long long NV,
                 int C, int K, mvd_stream_t stream) {
  MVD_REQUIRE(z && w && logits && NV > 0, "head_fwd: bad arguments");
  MVD_REQUIRE(K >= 1 && K <= kMaxHeadK, "head_fwd: num_classes must be in 1..8 (got %d)", K);
  MVD_REQUIRE(C % 8 == 0 && ldz % 8 == 0 && ((uintptr_t)z & 15) == 0 && C <= 1024, "head_fwd: C/alignment");
  MVD_REQUIRE(ldl >= K, "head_fwd: ldl < K");
#define CALL(KK) head_fwd_launch<KK>((const bf16*)z, ldz, w, bias, (bf16*)logits, ldl, NV, C, (cudaStream_t)stream)
  HEAD_DISPATCH(K, CALL)
#undef CALL
  return MVD_ERR_UNSUPPORTED;
}

int mvd_head_bwd(const void* dlogits, int ldl, const void* z, int ldz, const float* w, void* dz, int lddz,
                 float* dw, float* dbias, long long NV, int C, int K, int accumulate_dz, mvd_stream_t stream) {
  MVD_REQUIRE(dlogits && z && w && NV > 0, "head_bwd: bad arguments");
  MVD_REQUIRE(K >= 1 && K <= kMaxHeadK, "head_bwd: num_classes must be in 1..8 (got %d)", K);
  MVD_REQUIRE(C % 8 == 0 && ldz % 8 == 0 && ((uintptr_t)z & 15) == 0 && C <= 1024, "head_bwd: C/alignment");
  MVD_REQUIRE(!dz || (lddz % 8 == 0 && ((uintptr_t)dz & 15) == 0), "head_bwd: dz alignment");
#define CALL(KK)                                                                                               \
  head_bwd_launch<KK>((const bf16*)dlogits, ldl, (const bf16*)z, ldz, w, (bf16*)dz, lddz, dw, dbias, NV, C, \
                      accumulate_dz != 0, (cudaStream_t)stream)
  HEAD_DISPATCH(K, CALL)
#undef CALL
  return MVD_ERR_UNSUPPORTED;
}

int mvd_sw_accumulate(const void* pred, int ldp, const float* gaussian, float scale, float* acc, float* npred, int K,
                      int d, int h, int w, int D, int H, int W, int z0, int y0, int x0, int flip_mask,
                      mvd_stream_t stream) {
  MVD_REQUIRE(pred && acc && K >= 1 && K <= kMaxHeadK && ldp >= K && flip_mask >= 0 && flip_mask < 8,
              "sw_accumulate: bad arguments");
  MVD_REQUIRE(d > 0 && h > 0 && w > 0 && z0 >= 0 && y0 >= 0 && x0 >= 0 && z0 + d <= D && y0 + h <= H && x0 + w <= W,
              "sw_accumulate: tile outside the volume");
  const int grid = grid_for((long long)d * h * w, 256, num_sms() * 8);
#define SW_CASE(KK)                                                                                              \
  case KK:                                                                                                       \
    sw_accumulate_kernel<KK><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)pred, ldp, gaussian, scale, acc, \
                                                                     npred, d, h, w, D, H, W, z0, y0, x0,          \
                                                                     flip_mask);                                   \
    break;
  switch (K) {
    SW_CASE(1) SW_CASE(2) SW_CASE(3) SW_CASE(4) SW_CASE(5) SW_CASE(6) SW_CASE(7) SW_CASE(8)
    default: break;
  }
#undef SW_CASE
  MVD_LAUNCH_CHECK("sw_accumulate");
  return MVD_OK;
}

int mvd_sw_finalize(float* acc, const float* npred, int K, long long vol, mvd_stream_t stream) {
  MVD_REQUIRE(acc && npred && K >= 1 && vol > 0, "sw_finalize: bad arguments");
  sw_finalize_kernel<<<grid_for(vol, 256, num_sms() * 8), 256, 0, (cudaStream_t)stream>>>(acc, npred, K, vol);
  MVD_LAUNCH_CHECK("sw_finalize");
  return MVD_OK;
}

}  // extern "C"
