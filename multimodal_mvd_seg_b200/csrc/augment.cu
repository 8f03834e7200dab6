// augment.cu -- training-batch augmentation on the GPU.
//
// Replaces the CPU transform chain of nnUNetTrainer.get_training_transforms (MVDTrainer.py:700-765), which the reference
// runs in 12+ batchgenerators worker processes per GPU: SpatialTransform (rotation + isotropic scaling through
// scipy.ndimage.map_coordinates, order 3 for the images / order 1 per label for the segmentation, constant border),
// GaussianNoise, GaussianBlur, BrightnessMultiplicative, ContrastAugmentation, Gamma (plain and inverted, retain_stats)
// and Mirror.  Every kernel takes the RANDOM DRAWS as explicit per-plane parameter arrays (the host samples them with the
// reference's distributions, multimodal_mvd_seg_b200/augment.py), so each transform is a deterministic function that the
// oracle (oracle/augment.py, numpy / scipy) restates and the tests compare.
// Volumes are fp32 [N][D][H][W] planes (N = B * C: NCDHW batches as the loader delivers them), a few tens of MB per batch:
// every kernel is a plain HBM-bound sweep (one thread per voxel, or per line for the recursive spline prefilter).
// batchgenerators itself is NOT part of the reference tree (third-party dependency): semantics restated from its
// published sources, see oracle/augment.py ("parity unpinned").
#include "common.cuh"

namespace mvd {
namespace {

constexpr float kPole = -0.26794919243112270647f;   // sqrt(3) - 2: the pole of the cubic B-spline prefilter

// ---- cubic B-spline prefilter along one axis, in place (scipy.ndimage.spline_filter1d, order 3, mode 'mirror' -- what
// ---- map_coordinates(mode='constant') applies before interpolating).  One thread per line.
__global__ void __launch_bounds__(128) spline_prefilter_kernel(float* __restrict__ vol, long long n_lines, int len,
                                                               long long inner, long long stride,
                                                               const unsigned char* __restrict__ apply,
                                                               long long lines_per_plane) {
  const long long line = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= n_lines) return;
  if (apply && !apply[line / lines_per_plane]) return;
  // line -> base offset: lines are enumerated as (outer, inner) with the filtered axis removed
  const long long outer = line / inner, in = line - outer * inner;
  float* c = vol + outer * (long long)len * stride + in;
  if (len < 2) return;
  const double z = (double)kPole;
  const double lambda = (1.0 - z) * (1.0 - 1.0 / z);   // 6
  // causal initialisation, mirror boundary (whole-sample symmetric), exact sum over one period
  double zi = z, sum = (double)c[0] * lambda;
  for (int i = 1; i < len; ++i) { sum += zi * (double)c[(long long)i * stride] * lambda; zi *= z; }
  for (int i = len - 2; i > 0; --i) { sum += zi * (double)c[(long long)i * stride] * lambda; zi *= z; }
  double prev = sum / (1.0 - zi);
  c[0] = (float)prev;
  // NOTE: running values are kept in double, stored as float (scipy filters in double and rounds once at the end; the
  // difference is below 1e-6 relative for image-range data)
  for (int i = 1; i < len; ++i) {
    prev = (double)c[(long long)i * stride] * lambda + z * prev;
    c[(long long)i * stride] = (float)prev;
  }
  // anticausal initialisation + recursion
  double last = (z / (z * z - 1.0)) * (z * (double)c[(long long)(len - 2) * stride] + (double)c[(long long)(len - 1) * stride]);
  c[(long long)(len - 1) * stride] = (float)last;
  for (int i = len - 2; i >= 0; --i) {
    last = z * (last - (double)c[(long long)i * stride]);
    c[(long long)i * stride] = (float)last;
  }
}

__device__ __forceinline__ int mirror_index(int i, int n) {   // whole-sample symmetric: -1 -> 1, n -> n - 2
  if (n == 1) return 0;
  const int period = 2 * n - 2;
  i = i % period;
  if (i < 0) i += period;
  return i < n ? i : period - i;
}

struct SpatialParams {
  const float* src; float* dst;
  int B, C, Di, Hi, Wi, D, H, W;
  const float* mat;    // [B][9] row-major: source offset (d,h,w) = M * centred output index (d,h,w)
  const int* mode;     // [B] 0 = centre crop (no interpolation), 1 = interpolate
  int order;           // 1 or 3 (src holds spline coefficients for 3)
  float cval;
};

__device__ __forceinline__ void cubic_weights(float t, float* w) {
  const float t2 = t * t, t3 = t2 * t, u = 1.f - t;
  w[0] = u * u * u * (1.f / 6.f);
  w[1] = (4.f - 6.f * t2 + 3.f * t3) * (1.f / 6.f);
  w[2] = (1.f + 3.f * t + 3.f * t2 - 3.f * t3) * (1.f / 6.f);
  w[3] = t3 * (1.f / 6.f);
}

// source coordinates of output voxel (d, h, w) of sample b (augment_spatial: zero-centred mesh -> rotate / scale ->
// + centre of the source volume)
__device__ __forceinline__ void source_coords(const float* M, int d, int h, int w, int D, int H, int W, int Di, int Hi,
                                              int Wi, float& cd, float& ch, float& cw) {
  const float zd = (float)d - 0.5f * (float)(D - 1), zh = (float)h - 0.5f * (float)(H - 1), zw = (float)w - 0.5f * (float)(W - 1);
  cd = M[0] * zd + M[1] * zh + M[2] * zw + (0.5f * (float)Di - 0.5f);
  ch = M[3] * zd + M[4] * zh + M[5] * zw + (0.5f * (float)Hi - 0.5f);
  cw = M[6] * zd + M[7] * zh + M[8] * zw + (0.5f * (float)Wi - 0.5f);
}

__global__ void __launch_bounds__(256) spatial_kernel(const SpatialParams P) {
  const long long V = (long long)P.D * P.H * P.W, total = (long long)P.B * P.C * V;
  const long long Vi = (long long)P.Di * P.Hi * P.Wi;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long j = i;
    const int w = (int)(j % P.W); j /= P.W;
    const int h = (int)(j % P.H); j /= P.H;
    const int d = (int)(j % P.D); j /= P.D;
    const int c = (int)(j % P.C);
    const int b = (int)(j / P.C);
    const float* src = P.src + ((long long)b * P.C + c) * Vi;
    float out;
    if (P.mode[b] == 0) {   // centre crop (batchgenerators center_crop: lower bound (in - out) // 2)
      const int sd = d + (P.Di - P.D) / 2, sh = h + (P.Hi - P.H) / 2, sw = w + (P.Wi - P.W) / 2;
      out = __ldg(src + ((long long)sd * P.Hi + sh) * P.Wi + sw);
    } else {
      float cd, ch, cw;
      source_coords(P.mat + b * 9, d, h, w, P.D, P.H, P.W, P.Di, P.Hi, P.Wi, cd, ch, cw);
      if (cd < 0.f || cd > (float)(P.Di - 1) || ch < 0.f || ch > (float)(P.Hi - 1) || cw < 0.f || cw > (float)(P.Wi - 1)) {
        out = P.cval;       // scipy NI_EXTEND_CONSTANT: coordinates outside [0, n - 1] take cval
      } else if (P.order == 3) {
        const int fd = (int)floorf(cd), fh = (int)floorf(ch), fw = (int)floorf(cw);
        float wd[4], wh[4], ww[4];
        cubic_weights(cd - (float)fd, wd);
        cubic_weights(ch - (float)fh, wh);
        cubic_weights(cw - (float)fw, ww);
        int iw[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) iw[k] = mirror_index(fw - 1 + k, P.Wi);
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const long long od = (long long)mirror_index(fd - 1 + a, P.Di) * P.Hi;
          float accd = 0.f;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float* row = src + (od + mirror_index(fh - 1 + e, P.Hi)) * P.Wi;
            accd += wh[e] * (ww[0] * __ldg(row + iw[0]) + ww[1] * __ldg(row + iw[1]) + ww[2] * __ldg(row + iw[2]) +
                             ww[3] * __ldg(row + iw[3]));
          }
          acc += wd[a] * accd;
        }
        out = acc;
      } else {
        const int fd = min((int)floorf(cd), P.Di - 2 < 0 ? 0 : P.Di - 2), fh = min((int)floorf(ch), P.Hi - 2 < 0 ? 0 : P.Hi - 2),
                  fw = min((int)floorf(cw), P.Wi - 2 < 0 ? 0 : P.Wi - 2);
        const float td = cd - (float)fd, th = ch - (float)fh, tw = cw - (float)fw;
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const int id = min(fd + a, P.Di - 1), ih = min(fh + e, P.Hi - 1), iw = min(fw + g, P.Wi - 1);
              acc += (a ? td : 1.f - td) * (e ? th : 1.f - th) * (g ? tw : 1.f - tw) *
                     __ldg(src + ((long long)id * P.Hi + ih) * P.Wi + iw);
            }
        out = acc;
      }
    }
    P.dst[i] = out;
  }
}

// segmentation: interpolate_img(..., is_seg=True, order=1): for every label c (ascending) the mask (seg == c) is
// interpolated linearly and voxels with a value >= 0.5 take c; outside the volume nothing is set (stays 0).
constexpr int kMaxLabels = 16;
__global__ void __launch_bounds__(256) spatial_seg_kernel(const SpatialParams P, int n_labels) {
  const long long V = (long long)P.D * P.H * P.W, total = (long long)P.B * P.C * V;
  const long long Vi = (long long)P.Di * P.Hi * P.Wi;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long j = i;
    const int w = (int)(j % P.W); j /= P.W;
    const int h = (int)(j % P.H); j /= P.H;
    const int d = (int)(j % P.D); j /= P.D;
    const int c = (int)(j % P.C);
    const int b = (int)(j / P.C);
    const float* src = P.src + ((long long)b * P.C + c) * Vi;
    float out = 0.f;
    if (P.mode[b] == 0) {
      const int sd = d + (P.Di - P.D) / 2, sh = h + (P.Hi - P.H) / 2, sw = w + (P.Wi - P.W) / 2;
      out = __ldg(src + ((long long)sd * P.Hi + sh) * P.Wi + sw);
    } else {
      float cd, ch, cw;
      source_coords(P.mat + b * 9, d, h, w, P.D, P.H, P.W, P.Di, P.Hi, P.Wi, cd, ch, cw);
      if (!(cd < 0.f || cd > (float)(P.Di - 1) || ch < 0.f || ch > (float)(P.Hi - 1) || cw < 0.f || cw > (float)(P.Wi - 1))) {
        const int fd = min((int)floorf(cd), P.Di - 2 < 0 ? 0 : P.Di - 2), fh = min((int)floorf(ch), P.Hi - 2 < 0 ? 0 : P.Hi - 2),
                  fw = min((int)floorf(cw), P.Wi - 2 < 0 ? 0 : P.Wi - 2);
        const float td = cd - (float)fd, th = ch - (float)fh, tw = cw - (float)fw;
        float wsum[kMaxLabels];
#pragma unroll
        for (int l = 0; l < kMaxLabels; ++l) wsum[l] = 0.f;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const int id = min(fd + a, P.Di - 1), ih = min(fh + e, P.Hi - 1), iw = min(fw + g, P.Wi - 1);
              const float wt = (a ? td : 1.f - td) * (e ? th : 1.f - th) * (g ? tw : 1.f - tw);
              const int lab = (int)__ldg(src + ((long long)id * P.Hi + ih) * P.Wi + iw);
#pragma unroll
              for (int l = 0; l < kMaxLabels; ++l)
                if (l == lab) wsum[l] += wt;
            }
#pragma unroll
        for (int l = 1; l < kMaxLabels; ++l)
          if (l < n_labels && wsum[l] >= 0.5f) out = (float)l;   // later (larger) labels overwrite, as the loop over labels does
      }
    }
    P.dst[i] = out;
  }
}

// ---- intensity transforms (per-plane parameters) ---------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {   // splitmix64 finaliser
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

// x += N(0, sigma[plane]) -- GaussianNoiseTransform (sigma = the drawn "variance" value, used as the scale of
// np.random.normal exactly as batchgenerators does); counter-based generator: (seed, plane, voxel) -> Box-Muller
__global__ void __launch_bounds__(256) gaussian_noise_kernel(float* __restrict__ x, long long V, int N,
                                                             const float* __restrict__ sigma, unsigned long long seed) {
  const long long total = (long long)N * V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i / V);
    const float s = sigma[p];
    if (s == 0.f) continue;
    const unsigned long long r = mix64(seed ^ mix64((unsigned long long)i));
    const float u1 = ((float)(unsigned)(r >> 40) + 1.f) * (1.f / 16777217.f);        // (0, 1)
    const float u2 = (float)(unsigned)((r >> 8) & 0xffffffu) * (1.f / 16777216.f);   // [0, 1)
    x[i] += s * sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
  }
}

// one axis of scipy.ndimage.gaussian_filter (truncate 4, mode 'reflect' = half-sample symmetric), sigma per plane
__global__ void __launch_bounds__(256) blur_axis_kernel(const float* __restrict__ src, float* __restrict__ dst, int N,
                                                        int D, int H, int W, int axis, const float* __restrict__ sigma) {
  const long long V = (long long)D * H * W, total = (long long)N * V;
  const int len = axis == 0 ? D : (axis == 1 ? H : W);
  const long long stride = axis == 0 ? (long long)H * W : (axis == 1 ? W : 1);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i / V);
    const float s = sigma[p];
    if (s <= 0.f) { dst[i] = src[i]; continue; }
    const long long v = i - (long long)p * V;
    const int pos = axis == 0 ? (int)(v / ((long long)H * W)) : (axis == 1 ? (int)((v / W) % H) : (int)(v % W));
    const int radius = (int)(4.f * s + 0.5f);
    const float inv2 = -0.5f / (s * s);
    float acc = 0.f, wsum = 0.f;
    for (int k = -radius; k <= radius; ++k) {
      int q = pos + k;
      // reflect: (d c b a | a b c d | d c b a)
      while (q < 0 || q >= len) q = q < 0 ? -q - 1 : 2 * len - 1 - q;
      const float wt = __expf(inv2 * (float)(k * k));
      acc += wt * __ldg(src + i + (long long)(q - pos) * stride);
      wsum += wt;
    }
    dst[i] = acc / wsum;
  }
}

// per-plane sum, sum of squares, minimum, maximum -> out[N][4] doubles (caller initialises: 0, 0, +inf, -inf)
__device__ __forceinline__ void atomic_min_double(double* a, double v) {
  unsigned long long* p = reinterpret_cast<unsigned long long*>(a);
  unsigned long long old = *p;
  while (__longlong_as_double((long long)old) > v) {
    const unsigned long long prev = atomicCAS(p, old, (unsigned long long)__double_as_longlong(v));
    if (prev == old) break;
    old = prev;
  }
}
__device__ __forceinline__ void atomic_max_double(double* a, double v) {
  unsigned long long* p = reinterpret_cast<unsigned long long*>(a);
  unsigned long long old = *p;
  while (__longlong_as_double((long long)old) < v) {
    const unsigned long long prev = atomicCAS(p, old, (unsigned long long)__double_as_longlong(v));
    if (prev == old) break;
    old = prev;
  }
}

__global__ void __launch_bounds__(256) plane_stats_kernel(const float* __restrict__ x, long long V, double* __restrict__ out) {
  const int p = blockIdx.y;
  const float* xp = x + (long long)p * V;
  double s = 0.0, q = 0.0;
  float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (long long)gridDim.x * blockDim.x) {
    const float v = __ldg(xp + i);
    s += (double)v;
    q += (double)v * (double)v;
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  __shared__ double ss[256], sq[256];
  __shared__ float smn[256], smx[256];
  ss[threadIdx.x] = s; sq[threadIdx.x] = q; smn[threadIdx.x] = mn; smx[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      ss[threadIdx.x] += ss[threadIdx.x + o];
      sq[threadIdx.x] += sq[threadIdx.x + o];
      smn[threadIdx.x] = fminf(smn[threadIdx.x], smn[threadIdx.x + o]);
      smx[threadIdx.x] = fmaxf(smx[threadIdx.x], smx[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    atomicAdd(out + 4 * p, ss[0]);
    atomicAdd(out + 4 * p + 1, sq[0]);
    atomic_min_double(out + 4 * p + 2, (double)smn[0]);
    atomic_max_double(out + 4 * p + 3, (double)smx[0]);
  }
}

// op 0: x *= a[p]                                                    (BrightnessMultiplicativeTransform; a = 1 skips)
// op 1: x = clip((x - mean) * a[p] + mean, min, max), stats = st0     (ContrastAugmentationTransform, preserve_range; a = 0 skips)
// op 2: x = ((s x - mn) / (rng + 1e-7)) ^ a[p] * rng + mn, sign s = invert ? -1 : 1, (mn, rng) of s x from st0; result
//       is left in the s x domain                                      (augment_gamma before retain_stats; a = 0 skips)
// op 3: x = s * ((x - mean1) / (std1 + 1e-8) * std0 + mean0): st0 = statistics of s x before the gamma, st1 = after
//       (retain_stats, then the inversion is undone; a = 0 skips)
__global__ void __launch_bounds__(256) intensity_kernel(float* __restrict__ x, long long V, int N, int op,
                                                        const float* __restrict__ a, const double* __restrict__ st0,
                                                        const double* __restrict__ st1, int invert) {
  const long long total = (long long)N * V;
  const double dV = (double)V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i / V);
    const float ap = a[p];
    float v = x[i];
    if (op == 0) {
      v *= ap;
    } else if (op == 1) {
      if (ap == 0.f) continue;
      const float mean = (float)(st0[4 * p] / dV), mn = (float)st0[4 * p + 2], mx = (float)st0[4 * p + 3];
      v = fminf(fmaxf((v - mean) * ap + mean, mn), mx);
    } else if (op == 2) {
      if (ap == 0.f) continue;
      // statistics were taken of the stored (un-inverted) data: min(-x) = -max(x)
      const float mn = invert ? -(float)st0[4 * p + 3] : (float)st0[4 * p + 2];
      const float mx = invert ? -(float)st0[4 * p + 2] : (float)st0[4 * p + 3];
      const float rng = mx - mn;
      const float sx = invert ? -v : v;
      v = powf((sx - mn) / (rng + 1e-7f), ap) * rng + mn;
    } else {
      if (ap == 0.f) continue;
      const double m0 = (invert ? -1.0 : 1.0) * st0[4 * p] / dV;
      const double var0 = fmax(st0[4 * p + 1] / dV - (st0[4 * p] / dV) * (st0[4 * p] / dV), 0.0);
      const double m1 = st1[4 * p] / dV;
      const double var1 = fmax(st1[4 * p + 1] / dV - m1 * m1, 0.0);
      const float r = (float)((double)(v - (float)m1) / (sqrt(var1) + 1e-8) * sqrt(var0) + m0);
      v = invert ? -r : r;
    }
    x[i] = v;
  }
}

// ---- SimulateLowResolutionTransform (MVDTrainer.py:721-725): per selected plane, nearest-neighbour down-sampling to
// ---- round(shape * zoom) followed by cubic up-sampling back, both skimage.transform.resize(mode='edge',
// ---- anti_aliasing=False) = scipy.ndimage.zoom(grid_mode=True, mode='nearest') + clipping to the input range.
// The low-resolution volume is written edge-padded by 12 voxels per side (what scipy pre-pads before its spline filter for
// mode 'nearest'; the prefilter's own boundary rule then only matters to z^12 = 1.4e-7), prefiltered in place, and sampled
// at (o + 0.5) * t / n - 0.5 + 12.
constexpr int kLowresPad = 12;
struct LowresParams {
  float* x;            // [N][D][H][W]
  float* buf;          // [N][buf_stride]: the padded low-resolution volumes, compact strides (t + 24 per axis)
  long long buf_stride;
  int N, D, H, W;
  const int* tshape;   // [N][3] target (low-resolution) shape, 0 = plane not selected
  double* minmax;      // [N][2] minimum / maximum of the low-resolution volume (caller initialises +inf / -inf)
};

__global__ void __launch_bounds__(256) lowres_down_kernel(const LowresParams P) {
  const int p = blockIdx.y;
  const int td = P.tshape[3 * p], th = P.tshape[3 * p + 1], tw = P.tshape[3 * p + 2];
  if (td <= 0) return;
  const int pd = td + 2 * kLowresPad, ph = th + 2 * kLowresPad, pw = tw + 2 * kLowresPad;
  const long long total = (long long)pd * ph * pw;
  const float* src = P.x + (long long)p * P.D * P.H * P.W;
  float* dst = P.buf + (long long)p * P.buf_stride;
  const double rd = (double)P.D / td, rh = (double)P.H / th, rw = (double)P.W / tw;
  float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long j = i;
    const int k = (int)(j % pw); j /= pw;
    const int e = (int)(j % ph);
    const int a = (int)(j / ph);
    const int ld = min(max(a - kLowresPad, 0), td - 1), lh = min(max(e - kLowresPad, 0), th - 1),
              lw = min(max(k - kLowresPad, 0), tw - 1);
    const int sd = min((int)floor((ld + 0.5) * rd), P.D - 1), sh = min((int)floor((lh + 0.5) * rh), P.H - 1),
              sw = min((int)floor((lw + 0.5) * rw), P.W - 1);
    const float v = __ldg(src + ((long long)sd * P.H + sh) * P.W + sw);
    dst[i] = v;
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  __shared__ float smn[256], smx[256];
  smn[threadIdx.x] = mn; smx[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      smn[threadIdx.x] = fminf(smn[threadIdx.x], smn[threadIdx.x + o]);
      smx[threadIdx.x] = fmaxf(smx[threadIdx.x], smx[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    atomic_min_double(P.minmax + 2 * p, (double)smn[0]);
    atomic_max_double(P.minmax + 2 * p + 1, (double)smx[0]);
  }
}

// cubic prefilter of the padded low-resolution volumes, one axis per launch, one thread per line (per-plane extents)
__global__ void __launch_bounds__(128) lowres_prefilter_kernel(const LowresParams P, int axis) {
  const int p = blockIdx.y;
  const int td = P.tshape[3 * p];
  if (td <= 0) return;
  const int dims[3] = {td + 2 * kLowresPad, P.tshape[3 * p + 1] + 2 * kLowresPad, P.tshape[3 * p + 2] + 2 * kLowresPad};
  const int len = dims[axis];
  const long long inner = axis == 0 ? (long long)dims[1] * dims[2] : (axis == 1 ? dims[2] : 1);
  const long long n_lines = (long long)dims[0] * dims[1] * dims[2] / len;
  const long long line = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (line >= n_lines) return;
  const long long outer = line / inner, in = line - outer * inner;
  float* c = P.buf + (long long)p * P.buf_stride + outer * (long long)len * inner + in;
  const long long stride = inner;
  const double z = (double)kPole;
  const double lambda = (1.0 - z) * (1.0 - 1.0 / z);
  double zi = z, sum = (double)c[0] * lambda;
  for (int i = 1; i < len; ++i) { sum += zi * (double)c[(long long)i * stride] * lambda; zi *= z; }
  for (int i = len - 2; i > 0; --i) { sum += zi * (double)c[(long long)i * stride] * lambda; zi *= z; }
  double prev = sum / (1.0 - zi);
  c[0] = (float)prev;
  for (int i = 1; i < len; ++i) {
    prev = (double)c[(long long)i * stride] * lambda + z * prev;
    c[(long long)i * stride] = (float)prev;
  }
  double last = (z / (z * z - 1.0)) * (z * (double)c[(long long)(len - 2) * stride] + (double)c[(long long)(len - 1) * stride]);
  c[(long long)(len - 1) * stride] = (float)last;
  for (int i = len - 2; i >= 0; --i) {
    last = z * (last - (double)c[(long long)i * stride]);
    c[(long long)i * stride] = (float)last;
  }
}

__global__ void __launch_bounds__(256) lowres_up_kernel(const LowresParams P) {
  const int p = blockIdx.y;
  const int td = P.tshape[3 * p], th = P.tshape[3 * p + 1], tw = P.tshape[3 * p + 2];
  if (td <= 0) return;
  const int ph = th + 2 * kLowresPad, pw = tw + 2 * kLowresPad;
  const long long V = (long long)P.D * P.H * P.W;
  const float* co = P.buf + (long long)p * P.buf_stride;
  float* dst = P.x + (long long)p * V;
  const float lo = (float)P.minmax[2 * p], hi = (float)P.minmax[2 * p + 1];
  const float rd = (float)td / (float)P.D, rh = (float)th / (float)P.H, rw = (float)tw / (float)P.W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += (long long)gridDim.x * blockDim.x) {
    long long j = i;
    const int w = (int)(j % P.W); j /= P.W;
    const int h = (int)(j % P.H);
    const int d = (int)(j / P.H);
    const float cd = ((float)d + 0.5f) * rd - 0.5f + (float)kLowresPad, ch = ((float)h + 0.5f) * rh - 0.5f + (float)kLowresPad,
                cw = ((float)w + 0.5f) * rw - 0.5f + (float)kLowresPad;
    const int fd = (int)floorf(cd), fh = (int)floorf(ch), fw = (int)floorf(cw);
    float wd[4], wh[4], ww[4];
    cubic_weights(cd - (float)fd, wd);
    cubic_weights(ch - (float)fh, wh);
    cubic_weights(cw - (float)fw, ww);
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float accd = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float* row = co + ((long long)(fd - 1 + a) * ph + (fh - 1 + e)) * pw + (fw - 1);
        accd += wh[e] * (ww[0] * __ldg(row) + ww[1] * __ldg(row + 1) + ww[2] * __ldg(row + 2) + ww[3] * __ldg(row + 3));
      }
      acc += wd[a] * accd;
    }
    dst[i] = fminf(fmaxf(acc, lo), hi);      // skimage resize(clip=True)
  }
}

// MirrorTransform: per sample, flip along d / h / w when flips[b][axis] != 0 (images and segmentation alike)
__global__ void __launch_bounds__(256) mirror_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C,
                                                     int D, int H, int W, const unsigned char* __restrict__ flips) {
  const long long V = (long long)D * H * W, total = (long long)B * C * V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long j = i;
    int w = (int)(j % W); j /= W;
    int h = (int)(j % H); j /= H;
    int d = (int)(j % D); j /= D;
    const long long plane = j;
    const int b = (int)(plane / C);
    if (flips[3 * b]) d = D - 1 - d;
    if (flips[3 * b + 1]) h = H - 1 - h;
    if (flips[3 * b + 2]) w = W - 1 - w;
    dst[i] = __ldg(src + plane * V + ((long long)d * H + h) * W + w);
  }
}

}  // namespace
}  // namespace mvd

using namespace mvd;

extern "C" {

int mvd_aug_spline_prefilter(float* vol, int N, int D, int H, int W, const unsigned char* apply, mvd_stream_t stream) {
  MVD_REQUIRE(vol && N > 0 && D > 0 && H > 0 && W > 0, "aug_spline_prefilter: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const long long V = (long long)D * H * W;
  // axis 0 (d): lines enumerated by (plane, h*W + w); axis 1 (h): (plane*D + d, w); axis 2 (w): (plane*D*H + d*H + h)
  const long long lines0 = (long long)N * H * W, lines1 = (long long)N * D * W, lines2 = (long long)N * D * H;
  spline_prefilter_kernel<<<(unsigned)((lines0 + 127) / 128), 128, 0, st>>>(vol, lines0, D, (long long)H * W, (long long)H * W,
                                                                              apply, (long long)H * W);
  MVD_LAUNCH_CHECK("aug_spline_prefilter(d)");
  spline_prefilter_kernel<<<(unsigned)((lines1 + 127) / 128), 128, 0, st>>>(vol, lines1, H, W, W, apply, (long long)D * W);
  MVD_LAUNCH_CHECK("aug_spline_prefilter(h)");
  spline_prefilter_kernel<<<(unsigned)((lines2 + 127) / 128), 128, 0, st>>>(vol, lines2, W, 1, 1, apply, (long long)D * H);
  MVD_LAUNCH_CHECK("aug_spline_prefilter(w)");
  (void)V;
  return MVD_OK;
}

int mvd_aug_spatial(const float* src, int B, int C, int Di, int Hi, int Wi, float* dst, int D, int H, int W,
                    const float* mat, const int* mode, int order, float cval, int seg_labels, mvd_stream_t stream) {
  MVD_REQUIRE(src && dst && mat && mode && B > 0 && C > 0, "aug_spatial: bad arguments");
  MVD_REQUIRE(Di >= D && Hi >= H && Wi >= W && D > 0 && H > 0 && W > 0, "aug_spatial: the source must be at least the output size");
  MVD_REQUIRE(seg_labels > 0 || order == 1 || order == 3, "aug_spatial: order must be 1 or 3");
  MVD_REQUIRE(seg_labels <= kMaxLabels, "aug_spatial: at most %d labels", kMaxLabels);
  SpatialParams P;
  P.src = src; P.dst = dst; P.B = B; P.C = C; P.Di = Di; P.Hi = Hi; P.Wi = Wi; P.D = D; P.H = H; P.W = W;
  P.mat = mat; P.mode = mode; P.order = order; P.cval = cval;
  const long long total = (long long)B * C * D * H * W;
  const int grid = grid_for(total, 256, num_sms() * 16);
  if (seg_labels > 0) spatial_seg_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P, seg_labels);
  else spatial_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
  MVD_LAUNCH_CHECK("aug_spatial");
  return MVD_OK;
}

int mvd_aug_gaussian_noise(float* x, long long V, int N, const float* sigma, unsigned long long seed, mvd_stream_t stream) {
  MVD_REQUIRE(x && sigma && V > 0 && N > 0, "aug_gaussian_noise: bad arguments");
  gaussian_noise_kernel<<<grid_for((long long)N * V, 256 * 4, num_sms() * 16), 256, 0, (cudaStream_t)stream>>>(x, V, N, sigma, seed);
  MVD_LAUNCH_CHECK("aug_gaussian_noise");
  return MVD_OK;
}

int mvd_aug_gaussian_blur(float* x, float* tmp, int N, int D, int H, int W, const float* sigma, mvd_stream_t stream) {
  MVD_REQUIRE(x && tmp && sigma && N > 0 && D > 0 && H > 0 && W > 0, "aug_gaussian_blur: bad arguments");
  const long long total = (long long)N * D * H * W;
  const int grid = grid_for(total, 256, num_sms() * 16);
  cudaStream_t st = (cudaStream_t)stream;
  blur_axis_kernel<<<grid, 256, 0, st>>>(x, tmp, N, D, H, W, 0, sigma);
  MVD_LAUNCH_CHECK("aug_gaussian_blur(d)");
  blur_axis_kernel<<<grid, 256, 0, st>>>(tmp, x, N, D, H, W, 1, sigma);
  MVD_LAUNCH_CHECK("aug_gaussian_blur(h)");
  blur_axis_kernel<<<grid, 256, 0, st>>>(x, tmp, N, D, H, W, 2, sigma);
  MVD_LAUNCH_CHECK("aug_gaussian_blur(w)");
  cudaError_t e = cudaMemcpyAsync(x, tmp, (size_t)total * sizeof(float), cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) { set_error("aug_gaussian_blur: %s", cudaGetErrorString(e)); return MVD_ERR_CUDA; }
  return MVD_OK;
}

int mvd_aug_plane_stats(const float* x, long long V, int N, double* out, mvd_stream_t stream) {
  MVD_REQUIRE(x && out && V > 0 && N > 0, "aug_plane_stats: bad arguments");
  int bx = (int)((V + 256 * 8 - 1) / (256 * 8));
  const int cap = (num_sms() * 8 + N - 1) / N;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  plane_stats_kernel<<<dim3(bx, N), 256, 0, (cudaStream_t)stream>>>(x, V, out);
  MVD_LAUNCH_CHECK("aug_plane_stats");
  return MVD_OK;
}

int mvd_aug_intensity(float* x, long long V, int N, int op, const float* a, const double* stats0, const double* stats1,
                      int invert, mvd_stream_t stream) {
  MVD_REQUIRE(x && a && V > 0 && N > 0 && op >= 0 && op <= 3, "aug_intensity: bad arguments");
  MVD_REQUIRE(op == 0 || stats0, "aug_intensity: op %d needs the plane statistics", op);
  MVD_REQUIRE(op != 3 || stats1, "aug_intensity: op 3 needs the statistics after the gamma");
  intensity_kernel<<<grid_for((long long)N * V, 256 * 4, num_sms() * 16), 256, 0, (cudaStream_t)stream>>>(x, V, N, op, a, stats0,
                                                                                                          stats1, invert);
  MVD_LAUNCH_CHECK("aug_intensity");
  return MVD_OK;
}

int mvd_aug_simulate_lowres(float* x, int N, int D, int H, int W, const int* tshape, float* buf, long long buf_stride,
                            double* minmax, mvd_stream_t stream) {
  MVD_REQUIRE(x && tshape && buf && minmax && N > 0 && D > 0 && H > 0 && W > 0, "aug_simulate_lowres: bad arguments");
  MVD_REQUIRE(buf_stride >= (long long)(D + 2 * kLowresPad) * (H + 2 * kLowresPad) * (W + 2 * kLowresPad),
              "aug_simulate_lowres: the scratch needs (D + 24)(H + 24)(W + 24) floats per plane");
  LowresParams P;
  P.x = x; P.buf = buf; P.buf_stride = buf_stride; P.N = N; P.D = D; P.H = H; P.W = W; P.tshape = tshape; P.minmax = minmax;
  cudaStream_t st = (cudaStream_t)stream;
  const int per_plane = (num_sms() * 8 + N - 1) / N;
  lowres_down_kernel<<<dim3(per_plane, N), 256, 0, st>>>(P);
  MVD_LAUNCH_CHECK("aug_simulate_lowres(down)");
  const long long max_lines = (long long)(D + 2 * kLowresPad) * (H + 2 * kLowresPad) * (W + 2 * kLowresPad) /
                              (D < H ? (D < W ? D : W) : (H < W ? H : W));
  for (int axis = 0; axis < 3; ++axis) {
    lowres_prefilter_kernel<<<dim3((unsigned)((max_lines + 127) / 128), N), 128, 0, st>>>(P, axis);
    MVD_LAUNCH_CHECK("aug_simulate_lowres(prefilter)");
  }
  lowres_up_kernel<<<dim3(per_plane, N), 256, 0, st>>>(P);
  MVD_LAUNCH_CHECK("aug_simulate_lowres(up)");
  return MVD_OK;
}

int mvd_aug_mirror(const float* src, float* dst, int B, int C, int D, int H, int W, const unsigned char* flips,
                   mvd_stream_t stream) {
  MVD_REQUIRE(src && dst && flips && src != dst && B > 0 && C > 0 && D > 0 && H > 0 && W > 0, "aug_mirror: bad arguments");
  mirror_kernel<<<grid_for((long long)B * C * D * H * W, 256 * 4, num_sms() * 16), 256, 0, (cudaStream_t)stream>>>(
      src, dst, B, C, D, H, W, flips);
  MVD_LAUNCH_CHECK("aug_mirror");
  return MVD_OK;
}

}  // extern "C"
