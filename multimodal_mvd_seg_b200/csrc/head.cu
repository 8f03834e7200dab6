// head.cu -- the decoder's 1x1x1 segmentation heads (nn.Conv3d(C, num_classes, 1), UNetDecoder.py:67-70).
// K = num_classes <= 8 outputs per voxel: HBM-bound, CUDA cores.  One thread per voxel, the [K][C] weights
// (rounded to bf16, as autocast feeds them to the reference conv) broadcast from shared memory.
#include "common.cuh"

namespace mvd {

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
      bf16x8 p = *reinterpret_cast<const bf16x8*>(zp + c0);
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
                                                            int lddz, long long NV, int C) {
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
    *reinterpret_cast<bf16x8*>(dz + v * lddz + cg * 8) = pack8(o);
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
      bf16x8 p = *reinterpret_cast<const bf16x8*>(z + v * ldz + cg * 8);
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

template <int K>
static int head_fwd_launch(const bf16* z, int ldz, const float* w, const float* b, bf16* out, int ldl, long long NV,
                           int C, cudaStream_t st) {
  int grid = grid_for(NV, 256, num_sms() * 8);
  head_fwd_kernel<K><<<grid, 256, (K * C + K) * sizeof(float), st>>>(z, ldz, w, b, out, ldl, NV, C);
  MVD_LAUNCH_CHECK("head_fwd");
  return MVD_OK;
}

template <int K>
static int head_bwd_launch(const bf16* dl, int ldl, const bf16* z, int ldz, const float* w, bf16* dz, int lddz,
                           float* dw, float* db, long long NV, int C, cudaStream_t st) {
  if (dz) {
    int grid = grid_for(NV * (C / 8), 256 * 2, num_sms() * 8);
    head_bwd_data_kernel<K><<<grid, 256, K * C * sizeof(float), st>>>(dl, ldl, w, dz, lddz, NV, C);
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
                 float* dw, float* dbias, long long NV, int C, int K, mvd_stream_t stream) {
  MVD_REQUIRE(dlogits && z && w && NV > 0, "head_bwd: bad arguments");
  MVD_REQUIRE(K >= 1 && K <= kMaxHeadK, "head_bwd: num_classes must be in 1..8 (got %d)", K);
  MVD_REQUIRE(C % 8 == 0 && ldz % 8 == 0 && ((uintptr_t)z & 15) == 0 && C <= 1024, "head_bwd: C/alignment");
  MVD_REQUIRE(!dz || (lddz % 8 == 0 && ((uintptr_t)dz & 15) == 0), "head_bwd: dz alignment");
#define CALL(KK)                                                                                               \
  head_bwd_launch<KK>((const bf16*)dlogits, ldl, (const bf16*)z, ldz, w, (bf16*)dz, lddz, dw, dbias, NV, C, \
                      (cudaStream_t)stream)
  HEAD_DISPATCH(K, CALL)
#undef CALL
  return MVD_ERR_UNSUPPORTED;
}

}  // extern "C"
