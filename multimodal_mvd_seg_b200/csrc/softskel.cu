// softskel.cu -- soft erosion / dilation / skeleton levels of training/loss/soft_skeleton.py:6-37 on fp32 volumes
// [B][D][H][W], forward and backward with PyTorch's tie rules:
//   * -max_pool3d(-x) along one axis -> gradient to the FIRST minimum in scan order (out-of-volume taps ignored)
//   * max_pool3d 3x3x3               -> gradient to the FIRST maximum in (d,h,w) scan order
//   * torch.min(a,b)                 -> 0.5 / 0.5 on ties
//   * relu'(0) = 0
// Round-1 form: one streaming stencil kernel per level (neighbour taps served by L1/L2; 4 B/voxel of HBM per read
// or written volume).  The backward scatters with fp32 atomics.
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
