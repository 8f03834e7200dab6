// conv_generic.cu -- shape-generic CUDA-core implicit-GEMM convolution (any kernel / stride / padding / channel count,
// pitched NDHWC bf16 activations, fp32 accumulation).  It serves the layers the tcgen05 path does not take
// (Cin in {1,2} stem, tiny bottleneck volumes, odd strides) and is the in-library cross-check for the tensor-core
// kernels in tests.  Reference ops replaced: nn.Conv3d / nn.ConvTranspose3d fprop, dgrad, wgrad
// (get_network_from_plans.py:70-83, UNetDecoder.py:55-65).
//
//   fprop : y[m, n]  = sum_{tap,k} x[src(m,tap), k] * Wf[tap][n][k]      m = conv-output voxel, n = Cout, k = Cin
//   dgrad : x[m, n]  = sum_{tap,k} y[src'(m,tap), k] * Wd[tap][n][k]     m = conv-input voxel,  n = Cin,  k = Cout
//   wgrad : dw[co][ci][tap] = sum_v y[v, co] * x[src(v,tap), ci]
// Tiles: 64 voxels x BN channels x 16 k per step, 256 threads, 4 x BN/16 outputs per thread.
#include "common.cuh"
#include "conv_common.cuh"

namespace mvd {

constexpr int BM = 64, BK = 16;

// source coordinate along one axis. fprop: o*s - p + t.  dgrad: (i + p - t)/s when divisible.
template <bool DGRAD>
__device__ __forceinline__ bool src_axis(int m, int t, int s, int p, int len_src, int& out) {
  if constexpr (!DGRAD) {
    out = m * s - p + t;
    return out >= 0 && out < len_src;
  } else {
    int num = m + p - t;
    if (num < 0) return false;
    if (s == 1) { out = num; return out < len_src; }
    if (num % s) return false;
    out = num / s;
    return out < len_src;
  }
}

struct IgemmParams {
  const bf16* src; int lds; int Kc; int Ds, Hs, Ws;    // gathered operand (fprop: x, dgrad: y)
  bf16* dst; int ldd; int N; int Dm, Hm, Wm;           // produced operand
  const bf16* w;                                       // [tap][N][Kc]
  const float* bias;                                   // [N] or null
  int B, kd, kh, kw, sd, sh, sw, pd, ph, pw;
  int accumulate;
  int flatK;                                           // Kc % 4 != 0: treat (tap,k) as one flat K axis, scalar gathers
};

template <bool DGRAD, int BN>
__global__ void __launch_bounds__(256) igemm_kernel(IgemmParams P) {
  constexpr int TN = BN / 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const long long Mtot = (long long)P.B * P.Dm * P.Hm * P.Wm;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int taps = P.kd * P.kh * P.kw;

  // loader role: row = tid/4 (voxel for A, channel n for B), 4 consecutive k
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  long long mrow = m0 + lrow;
  const bool mvalid = mrow < Mtot;
  int mb = 0, md = 0, mh = 0, mw = 0;
  if (mvalid) {
    long long t = mrow;
    mw = (int)(t % P.Wm); t /= P.Wm;
    mh = (int)(t % P.Hm); t /= P.Hm;
    md = (int)(t % P.Dm); mb = (int)(t / P.Dm);
  }
  const int tx = tid & 15, ty = tid >> 4;  // compute role
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  if (!P.flatK) {
    for (int tap = 0; tap < taps; ++tap) {
      const int tw_ = tap % P.kw, th_ = (tap / P.kw) % P.kh, td_ = tap / (P.kw * P.kh);
      int sdz, shy, swx;
      bool ok = mvalid && src_axis<DGRAD>(md, td_, P.sd, P.pd, P.Ds, sdz) &&
                src_axis<DGRAD>(mh, th_, P.sh, P.ph, P.Hs, shy) && src_axis<DGRAD>(mw, tw_, P.sw, P.pw, P.Ws, swx);
      const bf16* arow = ok ? P.src + ((((long long)mb * P.Ds + sdz) * P.Hs + shy) * P.Ws + swx) * P.lds : nullptr;
      const bf16* wtap = P.w + (long long)tap * P.N * P.Kc;
      for (int k0 = 0; k0 < P.Kc; k0 += BK) {
        // A tile
        float a4[4] = {0.f, 0.f, 0.f, 0.f};
        if (arow && k0 + lk < P.Kc) {  // Kc % 4 == 0 here
          uint2 raw = *reinterpret_cast<const uint2*>(arow + k0 + lk);
          float2 f0 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&raw.x));
          float2 f1 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&raw.y));
          a4[0] = f0.x; a4[1] = f0.y; a4[2] = f1.x; a4[3] = f1.y;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) As[lk + i][lrow] = a4[i];
        // B tile
        if (lrow < BN) {
          float b4[4] = {0.f, 0.f, 0.f, 0.f};
          const int n = n0 + lrow;
          if (n < P.N && k0 + lk < P.Kc) {
            uint2 raw = *reinterpret_cast<const uint2*>(wtap + (long long)n * P.Kc + k0 + lk);
            float2 f0 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&raw.x));
            float2 f1 = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&raw.y));
            b4[0] = f0.x; b4[1] = f0.y; b4[2] = f1.x; b4[3] = f1.y;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) Bs[lk + i][lrow] = b4[i];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
          float a[4], b[TN];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
          for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
      }
    }
  } else {
    const int Ktot = taps * P.Kc;
    for (int k0 = 0; k0 < Ktot; k0 += BK) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + lk + i;
        float av = 0.f, bv = 0.f;
        if (k < Ktot) {
          const int tap = k / P.Kc, kc = k - tap * P.Kc;
          const int tw_ = tap % P.kw, th_ = (tap / P.kw) % P.kh, td_ = tap / (P.kw * P.kh);
          int sdz, shy, swx;
          if (mvalid && src_axis<DGRAD>(md, td_, P.sd, P.pd, P.Ds, sdz) &&
              src_axis<DGRAD>(mh, th_, P.sh, P.ph, P.Hs, shy) && src_axis<DGRAD>(mw, tw_, P.sw, P.pw, P.Ws, swx))
            av = bf2f(P.src[((((long long)mb * P.Ds + sdz) * P.Hs + shy) * P.Ws + swx) * P.lds + kc]);
          const int n = n0 + lrow;
          if (lrow < BN && n < P.N) bv = bf2f(P.w[((long long)tap * P.N + n) * P.Kc + kc]);
        }
        As[lk + i][lrow] = av;
        if (lrow < BN) Bs[lk + i][lrow] = bv;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[4], b[TN];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  // epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= Mtot) continue;
    bf16* orow = P.dst + m * P.ldd;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= P.N) continue;
      float v = acc[i][j] + (P.bias ? round_bf(P.bias[n]) : 0.f);
      if (P.accumulate) v += bf2f(orow[n]);
      orow[n] = f2bf(v);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// wgrad: per tap a [Cout x Cin] GEMM reduced over a voxel run; fp32 atomics into dw (torch layout).
// flat mode (Cin % 4 != 0): the N axis is the flattened (tap, ci) pair.
// ---------------------------------------------------------------------------------------------------------------
struct WgradParams {
  const bf16* x; int ldx; int Cin; int Di, Hi, Wi;
  const bf16* y; int ldy; int Cout; int Do, Ho, Wo;
  float* dw; float* dbias;
  int B, kd, kh, kw, sd, sh, sw, pd, ph, pw;
  long long vox_per_block;
  int flat;
};

template <int BNW>
__global__ void __launch_bounds__(256) wgrad_kernel(WgradParams P) {
  constexpr int TN = BNW / 16;
  __shared__ float Ys[BK][BM + 4];   // [k = voxel][m = co]
  __shared__ float Xs[BK][BNW + 4];  // [k = voxel][n = ci or (tap,ci)]
  const int tid = threadIdx.x;
  const int taps = P.kd * P.kh * P.kw;
  const int ntile_n = P.flat ? (taps * P.Cin + BNW - 1) / BNW : (P.Cin + BNW - 1) / BNW;
  const int co0 = (blockIdx.x / ntile_n) * BM;
  const int n0 = (blockIdx.x % ntile_n) * BNW;
  const int tap_fixed = blockIdx.y;  // unused in flat mode
  const long long Vtot = (long long)P.B * P.Do * P.Ho * P.Wo;
  const long long v0 = (long long)blockIdx.z * P.vox_per_block;
  long long v1 = v0 + P.vox_per_block;
  if (v1 > Vtot) v1 = Vtot;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;  // dbias partial for co = co0 + tid (tid < 64), only in the n-tile 0 / tap 0 blocks
  const bool do_bias = P.dbias && (blockIdx.x % ntile_n) == 0 && (P.flat || tap_fixed == 0);

  // loader roles
  const int yk = tid >> 4, ym = (tid & 15) * 4;          // Ys: 16 voxels x 64 co, 4 co per thread
  for (long long vb = v0; vb < v1; vb += BK) {
    // ---- Ys
    {
      const long long v = vb + yk;
      float f[4] = {0.f, 0.f, 0.f, 0.f};
      if (v < v1) {
        const bf16* yr = P.y + v * P.ldy;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int co = co0 + ym + i;
          if (co < P.Cout) f[i] = bf2f(yr[co]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) Ys[yk][ym + i] = f[i];
    }
    // ---- Xs: BK voxels x BNW columns = 16*BNW elements, 256 threads
    for (int e = tid; e < BK * BNW; e += 256) {
      const int k = e / BNW, nn = e % BNW;
      const long long v = vb + k;
      float val = 0.f;
      const int n = n0 + nn;
      int tap, ci;
      bool nvalid;
      if (P.flat) { tap = n / P.Cin; ci = n - tap * P.Cin; nvalid = tap < taps; }
      else { tap = tap_fixed; ci = n; nvalid = ci < P.Cin; }
      if (v < v1 && nvalid) {
        long long t = v;
        const int ow = (int)(t % P.Wo); t /= P.Wo;
        const int oh = (int)(t % P.Ho); t /= P.Ho;
        const int od = (int)(t % P.Do); const int b = (int)(t / P.Do);
        const int tw_ = tap % P.kw, th_ = (tap / P.kw) % P.kh, td_ = tap / (P.kw * P.kh);
        const int iz = od * P.sd - P.pd + td_, iy = oh * P.sh - P.ph + th_, ix = ow * P.sw - P.pw + tw_;
        if (iz >= 0 && iz < P.Di && iy >= 0 && iy < P.Hi && ix >= 0 && ix < P.Wi)
          val = bf2f(P.x[((((long long)b * P.Di + iz) * P.Hi + iy) * P.Wi + ix) * P.ldx + ci]);
      }
      Xs[k][nn] = val;
    }
    __syncthreads();
    if (do_bias && tid < BM) {
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) bsum += Ys[kk][tid];
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[TN];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Ys[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Xs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= P.Cout) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      int tap, ci;
      if (P.flat) { tap = n / P.Cin; ci = n - tap * P.Cin; if (tap >= taps) continue; }
      else { tap = tap_fixed; ci = n; if (ci >= P.Cin) continue; }
      atomicAdd(&P.dw[((long long)co * P.Cin + ci) * taps + tap], acc[i][j]);
    }
  }
  if (do_bias && tid < BM && co0 + tid < P.Cout) atomicAdd(&P.dbias[co0 + tid], bsum);
}

// per-channel sum over voxels of a pitched bf16 tensor (bias gradient of a transposed conv)
__global__ void __launch_bounds__(256) channel_sum_kernel(const bf16* __restrict__ g, int ld, long long NV, int C,
                                                          float* __restrict__ out, long long rows_per_block) {
  extern __shared__ float sacc[];
  for (int i = threadIdx.x; i < C; i += 256) sacc[i] = 0.f;
  __syncthreads();
  long long v0 = (long long)blockIdx.x * rows_per_block, v1 = v0 + rows_per_block;
  if (v1 > NV) v1 = NV;
  const int cpt = (C + 255) / 256;  // channels per thread when C > 256
  if (C <= 256) {
    const int rows = 256 / C;
    const int c = threadIdx.x % C, r = threadIdx.x / C;
    float a = 0.f;
    if (r < rows)
      for (long long v = v0 + r; v < v1; v += rows) a += bf2f(g[v * ld + c]);
    if (r < rows) atomicAdd(&sacc[c], a);
  } else {
    for (int q = 0; q < cpt; ++q) {
      const int c = threadIdx.x + q * 256;
      if (c >= C) break;
      float a = 0.f;
      for (long long v = v0; v < v1; ++v) a += bf2f(g[v * ld + c]);
      atomicAdd(&sacc[c], a);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) atomicAdd(&out[i], sacc[i]);
}

// vector form (C % 8 == 0, 16-byte aligned pitched rows): thread = (voxel row r, 8-channel group cg), one resident wave,
// band sweep staged through a private cp.async ring of eight 16-byte slots per thread (a read-only stream needs ~100 KB
// in flight per SM to reach the HBM copy rate, scripts/probes/stream_probe.cu; plain load batches get serialised by ptxas); per-thread fp32 partials, one shared-memory reduction
// and one atomicAdd per (channel, block).  The scalar kernel above moved 2 bytes per load (0.2 ms per step at cfg-2).
__global__ void __launch_bounds__(256) channel_sum_vec_kernel(const bf16* __restrict__ g, int ld, long long NV, int C,
                                                              float* __restrict__ out) {
  extern __shared__ __align__(16) float sacc[];   // [C] sums, then the cp.async staging ring (8 x 256 x 16 B)
  for (int i = threadIdx.x; i < C; i += 256) sacc[i] = 0.f;
  __syncthreads();
  const int CG = C >> 3, rows = 256 / CG;
  const int cg = threadIdx.x % CG, r = threadIdx.x / CG;
  if (r < rows) {
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 0.f;
    const long long step = (long long)gridDim.x * rows;
    long long v = (long long)blockIdx.x * rows + r;
    const bf16* p = g + v * ld + cg * 8;
    const long long ps = step * ld;
    auto add = [&](const bf16x8& q) {
      float f[8];
      unpack8(q, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] += f[i];
    };
    const long long iters = v < NV ? (NV - v + step - 1) / step : 0;
    const bf16x8* ring = reinterpret_cast<const bf16x8*>(sacc + C) + threadIdx.x;   // private slots: [stage][256]
    const uint32_t ring_u = (uint32_t)__cvta_generic_to_shared(ring);
    constexpr int D = 8;
    auto issue = [&](long long i) {
      if (i < iters) cp_async16(ring_u + (uint32_t)(i & (D - 1)) * 4096, p + i * ps);
      cp_async_commit();
    };
    for (int i = 0; i < D - 1; ++i) issue(i);
    for (long long i = 0; i < iters; ++i) {
      issue(i + D - 1);
      cp_async_wait<D - 1>();
      add(ring[(int)(i & (D - 1)) * 256]);
    }
    cp_async_wait<0>();
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&sacc[cg * 8 + i], s[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) atomicAdd(&out[i], sacc[i]);
}

int generic_fprop(const mvd_conv3d_args* a, cudaStream_t st) {
  IgemmParams P;
  P.src = (const bf16*)a->x; P.lds = a->ldx; P.Kc = a->Cin; P.Ds = a->Di; P.Hs = a->Hi; P.Ws = a->Wi;
  P.dst = (bf16*)a->y; P.ldd = a->ldy; P.N = a->Cout; P.Dm = a->Do; P.Hm = a->Ho; P.Wm = a->Wo;
  P.w = (const bf16*)a->w; P.bias = a->bias; P.B = a->B;
  P.kd = a->kd; P.kh = a->kh; P.kw = a->kw; P.sd = a->sd; P.sh = a->sh; P.sw = a->sw;
  P.pd = a->pd; P.ph = a->ph; P.pw = a->pw; P.accumulate = a->accumulate;
  P.flatK = (a->Cin % 4 != 0) || (a->ldx % 4 != 0) || (((uintptr_t)a->x) & 7) || (((uintptr_t)a->w) & 7);
  const long long M = (long long)a->B * a->Do * a->Ho * a->Wo;
  if (a->Cout <= 32) {
    dim3 grid((unsigned)((M + BM - 1) / BM), (a->Cout + 31) / 32);
    igemm_kernel<false, 32><<<grid, 256, 0, st>>>(P);
  } else {
    dim3 grid((unsigned)((M + BM - 1) / BM), (a->Cout + 63) / 64);
    igemm_kernel<false, 64><<<grid, 256, 0, st>>>(P);
  }
  MVD_LAUNCH_CHECK("conv3d_fprop(generic)");
  return MVD_OK;
}

int generic_dgrad(const mvd_conv3d_args* a, cudaStream_t st) {
  IgemmParams P;
  P.src = (const bf16*)a->y; P.lds = a->ldy; P.Kc = a->Cout; P.Ds = a->Do; P.Hs = a->Ho; P.Ws = a->Wo;
  P.dst = (bf16*)a->x; P.ldd = a->ldx; P.N = a->Cin; P.Dm = a->Di; P.Hm = a->Hi; P.Wm = a->Wi;
  P.w = (const bf16*)a->w; P.bias = a->bias; P.B = a->B;
  P.kd = a->kd; P.kh = a->kh; P.kw = a->kw; P.sd = a->sd; P.sh = a->sh; P.sw = a->sw;
  P.pd = a->pd; P.ph = a->ph; P.pw = a->pw; P.accumulate = a->accumulate;
  P.flatK = (a->Cout % 4 != 0) || (a->ldy % 4 != 0) || (((uintptr_t)a->y) & 7) || (((uintptr_t)a->w) & 7);
  const long long M = (long long)a->B * a->Di * a->Hi * a->Wi;
  if (a->Cin <= 32) {
    dim3 grid((unsigned)((M + BM - 1) / BM), (a->Cin + 31) / 32);
    igemm_kernel<true, 32><<<grid, 256, 0, st>>>(P);
  } else {
    dim3 grid((unsigned)((M + BM - 1) / BM), (a->Cin + 63) / 64);
    igemm_kernel<true, 64><<<grid, 256, 0, st>>>(P);
  }
  MVD_LAUNCH_CHECK("conv3d_dgrad(generic)");
  return MVD_OK;
}

int generic_wgrad(const mvd_conv3d_args* a, cudaStream_t st) {
  const int taps = a->kd * a->kh * a->kw;
  MVD_CUDA(cudaMemsetAsync(a->dw, 0, sizeof(float) * (size_t)a->Cout * a->Cin * taps, st));
  if (a->dbias) MVD_CUDA(cudaMemsetAsync(a->dbias, 0, sizeof(float) * (size_t)a->Cout, st));
  WgradParams P;
  P.x = (const bf16*)a->x; P.ldx = a->ldx; P.Cin = a->Cin; P.Di = a->Di; P.Hi = a->Hi; P.Wi = a->Wi;
  P.y = (const bf16*)a->y; P.ldy = a->ldy; P.Cout = a->Cout; P.Do = a->Do; P.Ho = a->Ho; P.Wo = a->Wo;
  P.dw = a->dw; P.dbias = a->dbias; P.B = a->B;
  P.kd = a->kd; P.kh = a->kh; P.kw = a->kw; P.sd = a->sd; P.sh = a->sh; P.sw = a->sw;
  P.pd = a->pd; P.ph = a->ph; P.pw = a->pw;
  P.flat = (a->Cin < 16);
  const long long V = (long long)a->B * a->Do * a->Ho * a->Wo;
  const int bnw = (P.flat || a->Cin > 32) ? 64 : 32;
  const int ntile_n = P.flat ? (taps * a->Cin + bnw - 1) / bnw : (a->Cin + bnw - 1) / bnw;
  const int ntile_m = (a->Cout + BM - 1) / BM;
  const int gy = P.flat ? 1 : taps;
  const long long base_blocks = (long long)ntile_n * ntile_m * gy;
  long long want = (long long)num_sms() * 8;
  long long splits = (want + base_blocks - 1) / base_blocks;
  long long max_splits = (V + 255) / 256;  // at least 256 voxels per block
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  long long vpb = (V + splits - 1) / splits;
  vpb = ((vpb + BK - 1) / BK) * BK;
  splits = (V + vpb - 1) / vpb;
  P.vox_per_block = vpb;
  dim3 grid((unsigned)(ntile_n * ntile_m), (unsigned)gy, (unsigned)splits);
  if (bnw == 64) wgrad_kernel<64><<<grid, 256, 0, st>>>(P);
  else wgrad_kernel<32><<<grid, 256, 0, st>>>(P);
  MVD_LAUNCH_CHECK("conv3d_wgrad(generic)");
  return MVD_OK;
}

}  // namespace mvd

using namespace mvd;

extern "C" int mvd_channel_sum(const void* g, int ld, long long NV, int C, float* out, mvd_stream_t stream) {
  MVD_REQUIRE(g && out && NV > 0 && C > 0 && ld >= C && C <= 4096, "channel_sum: bad arguments");
  long long rpb = 4096;
  long long nblk = (NV + rpb - 1) / rpb;
  while (nblk < (long long)num_sms() * 2 && rpb > 64) {
    rpb >>= 1;
    nblk = (NV + rpb - 1) / rpb;
  }
  MVD_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)C, (cudaStream_t)stream));
  if (C % 8 == 0 && C <= 2048 && ld % 8 == 0 && ((uintptr_t)g & 15) == 0) {
    const int rows = 256 / (C / 8);
    long long nb = (long long)num_sms() * 4;
    const long long need = (NV + rows - 1) / rows;
    if (nb > need) nb = need;
    channel_sum_vec_kernel<<<(unsigned)nb, 256, C * sizeof(float) + 8 * 4096, (cudaStream_t)stream>>>((const bf16*)g, ld, NV, C, out);
    MVD_LAUNCH_CHECK("channel_sum");
    return MVD_OK;
  }
  channel_sum_kernel<<<(unsigned)nblk, 256, C * sizeof(float), (cudaStream_t)stream>>>((const bf16*)g, ld, NV, C, out, rpb);
  MVD_LAUNCH_CHECK("channel_sum");
  return MVD_OK;
}
