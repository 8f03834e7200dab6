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

// fp32 torch layout [Cout][Cin][taps] -> bf16 [tap][Cout][Cin] and/or [tap][Cin][Cout]
__global__ void __launch_bounds__(256) pack_conv_weights_kernel(const float* __restrict__ w, int Cout, int Cin,
                                                                int taps, bf16* __restrict__ wf,
                                                                bf16* __restrict__ wd) {
  const long long total = (long long)Cout * Cin * taps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // i indexes the fprop layout [tap][co][ci] so that its writes are coalesced
    int ci = (int)(i % Cin);
    long long r = i / Cin;
    int co = (int)(r % Cout);
    int tap = (int)(r / Cout);
    bf16 v = f2bf(w[((long long)co * Cin + ci) * taps + tap]);
    if (wf) wf[i] = v;
    if (wd) wd[((long long)tap * Cin + ci) * Cout + co] = v;
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

int mvd_pack_conv_weights(const float* w, int Cout, int Cin, int taps, void* w_fprop, void* w_dgrad,
                          mvd_stream_t stream) {
  MVD_REQUIRE(w && Cout > 0 && Cin > 0 && taps > 0 && (w_fprop || w_dgrad), "pack_conv_weights: bad arguments");
  int grid = grid_for((long long)Cout * Cin * taps, 256, num_sms() * 8);
  pack_conv_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, taps, (bf16*)w_fprop, (bf16*)w_dgrad);
  MVD_LAUNCH_CHECK("pack_conv_weights");
  return MVD_OK;
}

int mvd_scalar_axpy(const double* in, float scale, float* out, int accumulate, mvd_stream_t stream) {
  MVD_REQUIRE(in && out, "scalar_axpy: bad arguments");
  scalar_axpy_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(in, scale, out, accumulate);
  MVD_LAUNCH_CHECK("scalar_axpy");
  return MVD_OK;
}

}  // extern "C"
