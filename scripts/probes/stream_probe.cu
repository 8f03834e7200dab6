// stream_probe.cu -- which launch shape lets an InstanceNorm-apply-like streaming kernel (read bf16, FMA + LeakyReLU,
// write bf16; 2 + 2 bytes per element) reach the HBM copy rate on B200?  Standalone (no library), rotating operand
// sets larger than L2.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_probe stream_probe.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include <algorithm>

struct __align__(16) V8 { __nv_bfloat162 v[4]; };
__device__ __forceinline__ void body8(V8& p, const float* sc, const float* sh, float slope) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p.v[i]);
    float a = fmaf(t.x, sc[2 * i], sh[2 * i]), b = fmaf(t.y, sc[2 * i + 1], sh[2 * i + 1]);
    __nv_bfloat162 r = __floats2bfloat162_rn(a, b);
    float2 u = __bfloat1622float2(r);
    u.x = u.x > 0.f ? u.x : slope * u.x;
    u.y = u.y > 0.f ? u.y : slope * u.y;
    p.v[i] = __floats2bfloat162_rn(u.x, u.y);
  }
}
template <int HINT> __device__ __forceinline__ V8 ld(const V8* p) {
  if (HINT & 1) { uint4 u = __ldcs(reinterpret_cast<const uint4*>(p)); return *reinterpret_cast<V8*>(&u); }
  return *p;
}
template <int HINT> __device__ __forceinline__ void st(V8* p, const V8& v) {
  if (HINT & 2) __stcs(reinterpret_cast<uint4*>(p), *reinterpret_cast<const uint4*>(&v));
  else *p = v;
}

// A: persistent band sweep (the library's shape): vector i of iteration k = k*4*G*T + u*G*T + blockIdx*T + tid
template <int HINT, int U, bool MATH>
__global__ void __launch_bounds__(256) k_band(const V8* __restrict__ x, V8* __restrict__ y, long long n, float slope) {
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc[i] = 1.f + 0.01f * ((threadIdx.x & 3) * 8 + i); sh[i] = 0.001f * i; }
  const long long step = (long long)gridDim.x * 256;
  long long v = (long long)blockIdx.x * 256 + threadIdx.x;
  for (; v + (U - 1) * step < n; v += U * step) {
    V8 p[U];
#pragma unroll
    for (int u = 0; u < U; ++u) p[u] = ld<HINT>(x + v + u * step);
#pragma unroll
    for (int u = 0; u < U; ++u) { if (MATH) body8(p[u], sc, sh, slope); st<HINT>(y + v + u * step, p[u]); }
  }
  for (; v < n; v += step) { V8 p = ld<HINT>(x + v); if (MATH) body8(p, sc, sh, slope); st<HINT>(y + v, p); }
}
// C: persistent, block-contiguous chunks of U*256 vectors
template <int HINT, int U>
__global__ void __launch_bounds__(256) k_chunk(const V8* __restrict__ x, V8* __restrict__ y, long long n, float slope) {
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc[i] = 1.f + 0.01f * ((threadIdx.x & 3) * 8 + i); sh[i] = 0.001f * i; }
  const long long chunk = (long long)U * 256, nchunks = (n + chunk - 1) / chunk;
  for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const long long v0 = c * chunk + threadIdx.x;
    V8 p[U];
#pragma unroll
    for (int u = 0; u < U; ++u) if (v0 + u * 256 < n) p[u] = ld<HINT>(x + v0 + u * 256);
#pragma unroll
    for (int u = 0; u < U; ++u) if (v0 + u * 256 < n) { body8(p[u], sc, sh, slope); st<HINT>(y + v0 + u * 256, p[u]); }
  }
}
// E: per-thread cp.async pipeline (depth D), band sweep
template <int D>
__global__ void __launch_bounds__(256) k_cpasync(const V8* __restrict__ x, V8* __restrict__ y, long long n, float slope) {
  extern __shared__ __align__(16) uint8_t smem[];
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc[i] = 1.f + 0.01f * ((threadIdx.x & 3) * 8 + i); sh[i] = 0.001f * i; }
  V8* mine = reinterpret_cast<V8*>(smem) + threadIdx.x;   // stage s at mine[s * 256]
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(mine);
  const long long step = (long long)gridDim.x * 256;
  const long long v0 = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long iters = v0 < n ? (n - v0 + step - 1) / step : 0;
  auto issue = [&](long long i) {
    if (i < iters) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + (uint32_t)((i % D) * 256 * 16)), "l"(x + v0 + i * step));
    asm volatile("cp.async.commit_group;");
  };
  for (int i = 0; i < D - 1; ++i) issue(i);
  for (long long i = 0; i < iters; ++i) {
    issue(i + D - 1);
    asm volatile("cp.async.wait_group %0;" ::"n"(D - 1));
    V8 p = mine[(i % D) * 256];
    body8(p, sc, sh, slope);
    y[v0 + i * step] = p;
  }
}

// two inputs, one output (bwd_apply-like), band sweep
template <int HINT, int U>
__global__ void __launch_bounds__(256) k_band2(const V8* __restrict__ x, const V8* __restrict__ g, V8* __restrict__ y, long long n, float slope) {
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc[i] = 1.f + 0.01f * ((threadIdx.x & 3) * 8 + i); sh[i] = 0.001f * i; }
  const long long step = (long long)gridDim.x * 256;
  long long v = (long long)blockIdx.x * 256 + threadIdx.x;
  for (; v + (U - 1) * step < n; v += U * step) {
    V8 p[U], q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { p[u] = ld<HINT>(x + v + u * step); q[u] = ld<HINT>(g + v + u * step); }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      body8(p[u], sc, sh, slope);
#pragma unroll
      for (int i = 0; i < 4; ++i) p[u].v[i] = __hadd2(p[u].v[i], q[u].v[i]);
      st<HINT>(y + v + u * step, p[u]);
    }
  }
}
// inputs only (stats-like: NIN = 1, bwd_stats-like: NIN = 2)
template <int HINT, int U, int NIN>
__global__ void __launch_bounds__(256) k_reduce(const V8* __restrict__ x, const V8* __restrict__ g, float* __restrict__ out, long long n) {
  const long long step = (long long)gridDim.x * 256;
  long long v = (long long)blockIdx.x * 256 + threadIdx.x;
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  for (; v + (U - 1) * step < n; v += U * step) {
    V8 p[U], q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { p[u] = ld<HINT>(x + v + u * step); if (NIN == 2) q[u] = ld<HINT>(g + v + u * step); }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 t = __bfloat1622float2(p[u].v[i]);
        if (NIN == 2) { float2 w = __bfloat1622float2(q[u].v[i]); s[2 * i] = fmaf(t.x, w.x, s[2 * i]); s[2 * i + 1] = fmaf(t.y, w.y, s[2 * i + 1]); }
        else { s[2 * i] = fmaf(t.x, t.x, s[2 * i]); s[2 * i + 1] = fmaf(t.y, t.y, s[2 * i + 1]); }
      }
  }
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += s[i];
  if (tot == 123.456f) out[0] = tot;
}

// cp.async-staged inputs (per-thread private ring of D stages per input: no register cost for the bytes in flight)
template <int D, int NIN, bool WRITE>
__global__ void __launch_bounds__(256) k_cpa2(const V8* __restrict__ x, const V8* __restrict__ g, V8* __restrict__ y, float* __restrict__ out, long long n, float slope) {
  extern __shared__ __align__(16) uint8_t smem[];
  float sc[8], sh[8], s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc[i] = 1.f + 0.01f * ((threadIdx.x & 3) * 8 + i); sh[i] = 0.001f * i; s[i] = 0.f; }
  V8* mine = reinterpret_cast<V8*>(smem) + threadIdx.x;   // stage s of input j at mine[(j * D + s) * 256]
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(mine);
  const long long step = (long long)gridDim.x * 256;
  const long long v0 = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long iters = v0 < n ? (n - v0 + step - 1) / step : 0;
  auto issue = [&](long long i) {
    if (i < iters) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + (uint32_t)((i % D) * 4096)), "l"(x + v0 + i * step));
      if (NIN == 2) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + (uint32_t)((D + i % D) * 4096)), "l"(g + v0 + i * step));
    }
    asm volatile("cp.async.commit_group;");
  };
  for (int i = 0; i < D - 1; ++i) issue(i);
  for (long long i = 0; i < iters; ++i) {
    issue(i + D - 1);
    asm volatile("cp.async.wait_group %0;" ::"n"(D - 1));
    V8 p = mine[(i % D) * 256], q;
    if (NIN == 2) q = mine[(D + i % D) * 256];
    if (WRITE) {
      body8(p, sc, sh, slope);
      if (NIN == 2) {
#pragma unroll
        for (int k = 0; k < 4; ++k) p.v[k] = __hadd2(p.v[k], q.v[k]);
      }
      st<2>(y + v0 + i * step, p);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 t = __bfloat1622float2(p.v[k]);
        float2 w = NIN == 2 ? __bfloat1622float2(q.v[k]) : t;
        s[2 * k] = fmaf(t.x, w.x, s[2 * k]); s[2 * k + 1] = fmaf(t.y, w.y, s[2 * k + 1]);
      }
    }
  }
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += s[i];
  if (tot == 123.456f) out[0] = tot;
}

int main() {
  const long long n = 2LL * 128 * 128 * 128 * 32 / 8;   // vectors of 8 bf16 (268 MB per tensor)
  const int SETS = 4;
  V8 *x[SETS], *y[SETS];
  for (int s = 0; s < SETS; ++s) { cudaMalloc(&x[s], n * 16); cudaMalloc(&y[s], n * 16); cudaMemset(x[s], 0x3c, n * 16); cudaMemset(y[s], 0x3c, n * 16); }
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* name, auto launch, double passes = 2.0) {
    for (int i = 0; i < 3; ++i) launch(x[i % SETS], y[i % SETS]);
    cudaDeviceSynchronize();
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      for (int i = 0; i < 20; ++i) launch(x[i % SETS], y[i % SETS]);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms / 20);
    }
    cudaError_t e = cudaGetLastError();
    printf("%-46s %7.1f us  %7.1f GB/s %s\n", name, best * 1e3, passes * n * 16 / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
  };
  timeit("cudaMemcpyAsync D2D", [&](V8* a, V8* b) { cudaMemcpyAsync(b, a, n * 16, cudaMemcpyDeviceToDevice); });
#define CPA2(D, NIN, W, BPS, PASSES) { cudaFuncSetAttribute(k_cpa2<D, NIN, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, D * NIN * 4096); \
  timeit("cp.async depth=" #D " inputs=" #NIN " write=" #W " blocks/SM=" #BPS, [&](V8* a, V8* b) { k_cpa2<D, NIN, W><<<sms * BPS, 256, D * NIN * 4096>>>(a, W ? x[3] : b, b, (float*)y[0], n, 0.01f); }, PASSES); }
  CPA2(4, 2, false, 3, 2.0) CPA2(4, 2, false, 4, 2.0) CPA2(6, 2, false, 3, 2.0) CPA2(6, 2, false, 4, 2.0) CPA2(8, 2, false, 3, 2.0) CPA2(8, 2, false, 2, 2.0)
  CPA2(4, 2, true, 3, 3.0) CPA2(4, 2, true, 4, 3.0) CPA2(6, 2, true, 3, 3.0) CPA2(6, 2, true, 4, 3.0) CPA2(8, 2, true, 3, 3.0)
  CPA2(4, 1, true, 4, 2.0) CPA2(6, 1, true, 4, 2.0) CPA2(8, 1, true, 4, 2.0) CPA2(8, 1, true, 3, 2.0)
#define BAND2(H, U, BPS) timeit("2in-1out hint=" #H " U=" #U " blocks/SM=" #BPS, [&](V8* a, V8* b) { k_band2<H, U><<<sms * BPS, 256>>>(a, x[3], b, n, 0.01f); }, 3.0);
  BAND2(2, 2, 3) BAND2(2, 2, 4)
#define RED(H, U, NIN, BPS) timeit("reduce hint=" #H " U=" #U " inputs=" #NIN " blocks/SM=" #BPS, [&](V8* a, V8* b) { k_reduce<H, U, NIN><<<sms * BPS, 256>>>(a, b, (float*)y[0], n); }, (double)NIN);
  RED(0, 2, 2, 3) RED(0, 2, 2, 4)
  return 0;
}
