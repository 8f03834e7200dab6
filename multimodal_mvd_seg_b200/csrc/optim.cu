// optim.cu -- the optimiser tail of train_step (MVDTrainer.py:975-985, 482-486):
//   torch.nn.utils.clip_grad_norm_(params, 12)  +  torch.optim.SGD(lr, weight_decay=3e-5, momentum=0.99, nesterov=True)
// as two multi-tensor kernels over a device-resident pointer table (one launch each for all ~110 tensors of a net):
//   grad_sqnorm        : sum of squares of every gradient (fp32 in, double out), 4 B/param of HBM traffic
//   sgd_nesterov_clip  : p, g, buf read + p, buf written = 20 B/param
// HBM-bound; a chunk = 4096 contiguous elements of one tensor handled by one 256-thread block with float4 accesses.
#include "common.cuh"

namespace mvd {

constexpr int kChunk = 4096;

__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const uint64_t* __restrict__ ptrs,
                                                          const long long* __restrict__ numel,
                                                          const int* __restrict__ chunk_tensor,
                                                          const long long* __restrict__ chunk_offset,
                                                          double* __restrict__ sqnorm) {
  const int t = chunk_tensor[blockIdx.x];
  const long long off = chunk_offset[blockIdx.x];
  const float* g = reinterpret_cast<const float*>(ptrs[3 * t + 1]) + off;
  long long n = numel[t] - off;
  if (n > kChunk) n = kChunk;
  float acc = 0.f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const int n4 = (int)(n >> 2);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (int i = threadIdx.x; i < n4; i += 256) {
      float4 v = g4[i];
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += 256) acc += g[i] * g[i];
  } else {
    for (int i = threadIdx.x; i < n; i += 256) acc += g[i] * g[i];
  }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < 8; ++w) a += (double)red[w];
    atomicAdd(sqnorm, a);
  }
}

__global__ void __launch_bounds__(256) sgd_nesterov_clip_kernel(const uint64_t* __restrict__ ptrs,
                                                                const long long* __restrict__ numel,
                                                                const int* __restrict__ chunk_tensor,
                                                                const long long* __restrict__ chunk_offset,
                                                                const double* __restrict__ sqnorm, float gscale,
                                                                float max_norm, float lr, float wd, float mom) {
  const int t = chunk_tensor[blockIdx.x];
  const long long off = chunk_offset[blockIdx.x];
  float* p = reinterpret_cast<float*>(ptrs[3 * t + 0]) + off;
  const float* g = reinterpret_cast<const float*>(ptrs[3 * t + 1]) + off;
  float* b = reinterpret_cast<float*>(ptrs[3 * t + 2]) + off;
  long long n = numel[t] - off;
  if (n > kChunk) n = kChunk;
  float coef = gscale;
  if (max_norm > 0.f) {
    const float total = (float)sqrt(sqnorm[0]) * gscale;
    float c = max_norm / (total + 1e-6f);
    coef *= (c < 1.f ? c : 1.f);
  }
  for (int i = threadIdx.x; i < n; i += 256) {
    float pv = p[i];
    float gv = fmaf(wd, pv, g[i] * coef);
    float bv = fmaf(mom, b[i], gv);
    b[i] = bv;
    p[i] = pv - lr * fmaf(mom, bv, gv);
  }
}

// fp32 torch layout [Cout][Cin][taps] -> bf16 [tap][Cout][Cin] and/or [tap][Cin][Cout], for a whole TABLE of layers in
// one launch (mvd_pack_desc, device memory).  Block = a 16 (co) x 16 (ci) tile of one layer: every co row of the tile is
// 16*taps contiguous floats (coalesced, all loads of a thread independent); both packed layouts are then written in
// 32-byte runs (16 consecutive ci, resp. 16 consecutive co).
constexpr int kPackTile = 16, kPackMaxTaps = 27;
__global__ void __launch_bounds__(256) pack_conv_weights_kernel(const mvd_pack_desc* __restrict__ descs, int n) {
  __shared__ float tile[kPackTile][kPackTile * kPackMaxTaps + 1];
  // which layer does this block belong to (n <= a few dozen: linear scan)
  int li = 0;
  while (li + 1 < n && (int)blockIdx.x >= descs[li + 1].block_begin) ++li;
  const mvd_pack_desc d = descs[li];
  const int Cout = d.Cout, Cin = d.Cin, taps = d.taps;
  const int tiles_ci = (Cin + kPackTile - 1) / kPackTile;
  const int tb = (int)blockIdx.x - d.block_begin;
  const int ci0 = (tb % tiles_ci) * kPackTile, co0 = (tb / tiles_ci) * kPackTile;
  const int nci = min(kPackTile, Cin - ci0), nco = min(kPackTile, Cout - co0);
  const int rowlen = nci * taps;
  const float* w = d.w;
  for (int i = threadIdx.x; i < nco * rowlen; i += 256) {
    const int r = i / rowlen, c = i - r * rowlen;
    tile[r][c] = w[((long long)(co0 + r) * Cin + ci0) * taps + c];
  }
  __syncthreads();
  const int a = threadIdx.x >> 4, b = threadIdx.x & 15;
  bf16* wf = (bf16*)d.w_fprop;
  bf16* wd = (bf16*)d.w_dgrad;
  if (d.stem_kpad > 0) {                   // stem layout [Cout][kpad], k = tap*Cin + ci (Cin <= 16: one ci tile)
    const int kpad = d.stem_kpad;
    if (wf)
      for (int i = threadIdx.x; i < nco * kpad; i += 256) {
        const int r = i / kpad, k = i - r * kpad;
        const int t = k / Cin, ci = k - t * Cin;
        wf[(long long)(co0 + r) * kpad + k] = f2bf(k < taps * Cin ? tile[r][ci * taps + t] : 0.f);
      }
    return;
  }
  if (wf && a < nco && b < nci) {          // a = co, b = ci (fastest)
#pragma unroll 9
    for (int t = 0; t < taps; ++t)
      wf[((long long)t * Cout + co0 + a) * Cin + ci0 + b] = f2bf(tile[a][b * taps + t]);
  }
  if (wd && a < nci && b < nco) {          // a = ci, b = co (fastest)
#pragma unroll 9
    for (int t = 0; t < taps; ++t)
      wd[((long long)t * Cin + ci0 + a) * Cout + co0 + b] = f2bf(tile[b][a * taps + t]);
  }
}

__global__ void __launch_bounds__(128) zero_regions_kernel(float* __restrict__ base, const long long* __restrict__ table) {
  const long long off = table[2 * blockIdx.x], n = table[2 * blockIdx.x + 1];
  for (long long i = threadIdx.x; i < n; i += blockDim.x) base[off + i] = 0.f;
}

// out[c] = sum_b stats[b][c0 + c][0]   (the channel sums a conv epilogue left in its statistics buffer)
__global__ void stats_channel_sum_kernel(const double* __restrict__ stats, int B, int C, int c0, int n, float* __restrict__ out) {
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    double a = 0.0;
    for (int b = 0; b < B; ++b) a += stats[((long long)b * C + c0 + c) * 2];
    out[c] = (float)a;
  }
}

// SGD update of the conv weights FUSED with the refresh of their bf16 GEMM layouts (mvd_sgd_pack_desc): a block owns a
// 16 (co) x 16 (ci) x taps tile of one layer, applies  g = grad*coef + wd*p;  buf = mom*buf + g;  p -= lr*(g + mom*buf)
// on it (rows of 16*taps contiguous floats: coalesced reads / writes of p, grad, buf) and writes the two packed layouts
// of the UPDATED weights from the same tile -- the next forward pass finds them ready, the separate pack pass (a read of
// all fp32 weights per step) disappears.
__global__ void __launch_bounds__(256) sgd_pack_kernel(const mvd_sgd_pack_desc* __restrict__ descs, int n,
                                                       const double* __restrict__ sqnorm, float gscale, float max_norm,
                                                       float lr, float wd_, float mom) {
  __shared__ float tile[kPackTile][kPackTile * kPackMaxTaps + 1];
  int li = 0;
  while (li + 1 < n && (int)blockIdx.x >= descs[li + 1].pack.block_begin) ++li;
  const mvd_sgd_pack_desc d = descs[li];
  const int Cout = d.pack.Cout, Cin = d.pack.Cin, taps = d.pack.taps;
  const int tiles_ci = (Cin + kPackTile - 1) / kPackTile;
  const int tb = (int)blockIdx.x - d.pack.block_begin;
  const int ci0 = (tb % tiles_ci) * kPackTile, co0 = (tb / tiles_ci) * kPackTile;
  const int nci = min(kPackTile, Cin - ci0), nco = min(kPackTile, Cout - co0);
  const int rowlen = nci * taps;
  float coef = gscale;
  if (max_norm > 0.f) {
    const float total = (float)sqrt(sqnorm[0]) * gscale;
    const float c = max_norm / (total + 1e-6f);
    coef *= (c < 1.f ? c : 1.f);
  }
  float* w = const_cast<float*>(d.pack.w);
  // four independent element triples (p, grad, momentum) per thread in flight: the loop body is three dependent-free loads
  // and two stores, and a single iteration per trip leaves the memory pipe mostly idle (56 % of the copy rate)
  const int total = nco * rowlen;
  for (int i0 = threadIdx.x; i0 < total; i0 += 4 * 256) {
    float pv[4], gr[4], mv[4];
    long long o[4];
    int rr[4], cc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * 256;
      rr[u] = i / rowlen; cc[u] = i - rr[u] * rowlen;
      o[u] = ((long long)(co0 + rr[u]) * Cin + ci0) * taps + cc[u];
      if (i < total) { pv[u] = w[o[u]]; gr[u] = d.grad[o[u]]; mv[u] = d.momentum[o[u]]; }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (i0 + u * 256 < total) {
        const float gv = fmaf(wd_, pv[u], gr[u] * coef);
        const float bv = fmaf(mom, mv[u], gv);
        d.momentum[o[u]] = bv;
        const float pn = pv[u] - lr * fmaf(mom, bv, gv);
        w[o[u]] = pn;
        tile[rr[u]][cc[u]] = pn;
      }
    }
  }
  __syncthreads();
  const int a = threadIdx.x >> 4, b = threadIdx.x & 15;
  bf16* wf = (bf16*)d.pack.w_fprop;
  bf16* wdg = (bf16*)d.pack.w_dgrad;
  if (d.pack.stem_kpad > 0) {
    const int kpad = d.pack.stem_kpad;
    if (wf)
      for (int i = threadIdx.x; i < nco * kpad; i += 256) {
        const int r = i / kpad, k = i - r * kpad;
        const int t = k / Cin, ci = k - t * Cin;
        wf[(long long)(co0 + r) * kpad + k] = f2bf(k < taps * Cin ? tile[r][ci * taps + t] : 0.f);
      }
    return;
  }
  if (wf && a < nco && b < nci) {
#pragma unroll 9
    for (int t = 0; t < taps; ++t) wf[((long long)t * Cout + co0 + a) * Cin + ci0 + b] = f2bf(tile[a][b * taps + t]);
  }
  if (wdg && a < nci && b < nco) {
#pragma unroll 9
    for (int t = 0; t < taps; ++t) wdg[((long long)t * Cin + ci0 + a) * Cout + co0 + b] = f2bf(tile[b][a * taps + t]);
  }
}

__global__ void scalar_axpy_kernel(const double* __restrict__ in, float scale, float* __restrict__ out, int accumulate) {
  if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + (float)((double)scale * in[0]);
}

}  // namespace mvd

using namespace mvd;

extern "C" {

int mvd_grad_sqnorm(const uint64_t* ptrs, const long long* numel, const int* chunk_tensor,
                    const long long* chunk_offset, int n_chunks, double* sqnorm, mvd_stream_t stream) {
  MVD_REQUIRE(ptrs && numel && chunk_tensor && chunk_offset && sqnorm && n_chunks > 0, "grad_sqnorm: bad arguments");
  grad_sqnorm_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(ptrs, numel, chunk_tensor, chunk_offset, sqnorm);
  MVD_LAUNCH_CHECK("grad_sqnorm");
  return MVD_OK;
}

int mvd_sgd_nesterov_clip(const uint64_t* ptrs, const long long* numel, const int* chunk_tensor,
                          const long long* chunk_offset, int n_chunks, const double* sqnorm, float gscale,
                          float max_norm, float lr, float weight_decay, float momentum, mvd_stream_t stream) {
  MVD_REQUIRE(ptrs && numel && chunk_tensor && chunk_offset && n_chunks > 0, "sgd_nesterov_clip: bad arguments");
  MVD_REQUIRE(max_norm <= 0.f || sqnorm, "sgd_nesterov_clip: clipping needs the squared norm");
  sgd_nesterov_clip_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(ptrs, numel, chunk_tensor, chunk_offset, sqnorm,
                                                                       gscale, max_norm, lr, weight_decay, momentum);
  MVD_LAUNCH_CHECK("sgd_nesterov_clip");
  return MVD_OK;
}

int mvd_pack_conv_weights_multi(const mvd_pack_desc* descs_device, int n, int total_blocks, mvd_stream_t stream) {
  MVD_REQUIRE(descs_device && n > 0 && total_blocks > 0, "pack_conv_weights_multi: bad arguments");
  pack_conv_weights_kernel<<<total_blocks, 256, 0, (cudaStream_t)stream>>>(descs_device, n);
  MVD_LAUNCH_CHECK("pack_conv_weights");
  return MVD_OK;
}

int mvd_sgd_pack_conv_weights(const mvd_sgd_pack_desc* descs_device, int n, int total_blocks, const double* sqnorm,
                               float gscale, float max_norm, float lr, float weight_decay, float momentum,
                               mvd_stream_t stream) {
  MVD_REQUIRE(descs_device && n > 0 && total_blocks > 0, "sgd_pack_conv_weights: bad arguments");
  MVD_REQUIRE(max_norm <= 0.f || sqnorm, "sgd_pack_conv_weights: clipping needs the squared norm");
  sgd_pack_kernel<<<total_blocks, 256, 0, (cudaStream_t)stream>>>(descs_device, n, sqnorm, gscale, max_norm, lr,
                                                                  weight_decay, momentum);
  MVD_LAUNCH_CHECK("sgd_pack_conv_weights");
  return MVD_OK;
}

int mvd_pack_blocks(int Cout, int Cin) {
  return ((Cout + kPackTile - 1) / kPackTile) * ((Cin + kPackTile - 1) / kPackTile);
}

int mvd_stats_channel_sum(const double* stats, int B, int C, int c0, int n, float* out, mvd_stream_t stream) {
  MVD_REQUIRE(stats && out && B > 0 && c0 >= 0 && n > 0 && c0 + n <= C, "stats_channel_sum: bad arguments");
  stats_channel_sum_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(stats, B, C, c0, n, out);
  MVD_LAUNCH_CHECK("stats_channel_sum");
  return MVD_OK;
}

int mvd_zero_regions(float* base, const long long* table_device, int n, mvd_stream_t stream) {
  MVD_REQUIRE(base && table_device && n > 0, "zero_regions: bad arguments");
  zero_regions_kernel<<<n, 128, 0, (cudaStream_t)stream>>>(base, table_device);
  MVD_LAUNCH_CHECK("zero_regions");
  return MVD_OK;
}

int mvd_zero_bytes(void* ptr, size_t bytes, mvd_stream_t stream) {
  MVD_REQUIRE(ptr && bytes > 0, "zero_bytes: bad arguments");
  MVD_CUDA(cudaMemsetAsync(ptr, 0, bytes, (cudaStream_t)stream));
  return MVD_OK;
}

int mvd_scalar_axpy(const double* in, float scale, float* out, int accumulate, mvd_stream_t stream) {
  MVD_REQUIRE(in && out, "scalar_axpy: bad arguments");
  scalar_axpy_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(in, scale, out, accumulate);
  MVD_LAUNCH_CHECK("scalar_axpy");
  return MVD_OK;
}

}  // extern "C"
