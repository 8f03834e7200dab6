// losses_multi.cu -- the deep-supervision loss of a whole step in THREE launches, and the distillation KL in one.
//
//   DeepSupervisionWrapper(DC_and_CE_loss)(outputs, targets) for every active scale of every network
//   (nnUNetTrainer.py:359-374; robust_ce_loss.py:12-16) used to be fwd + finalize + bwd PER SCALE (12 launches per
//   network at cfg-2).  Here a table of "segments" (one per (network, scale): logits, targets, voxel count, weight) is
//   passed BY VALUE as a kernel parameter and
//     dice_ce_multi_fwd : one grid sweeps all segments (blocks are dealt out proportionally to the voxel counts), block
//                         sums go to fp64 accumulators, and the LAST block to finish (ticket counter) runs the scalar
//                         algebra for all segments: Dice coefficients for the backward + the weighted loss;
//     dice_ce_multi_bwd : one grid writes dlogits of all segments.
//   With batch_dice under data parallelism the per-segment sums are all-reduced between the two phases, so the
//   finalisation is available as its own launch as well (counter == NULL in the forward call).
//   distill_kl (other_loss.py:51-64): its gradients are point-wise, so ONE pass reads both logit tensors once and writes
//   the loss sum and both gradients, pre-multiplied by a caller-supplied upstream factor (lambda1 * T^2 / numel); a
//   rescale launch that exits at once when the real upstream gradient equals the assumed one keeps autograd exact.
//   Production shape only: C = 4 classes, dense 8-byte logit rows, V % 4 == 0 (other shapes: losses.cu, per scale).
#include <string.h>
#include "common.cuh"

namespace mvd {
namespace {

constexpr int kMaxSeg = MVD_DICE_CE_MAX_SEGMENTS;
constexpr int kStage = 4, kThreads = 256, C4 = 4;

struct Seg {
  const bf16* logits;
  const float* target;
  bf16* dlogits;
  long long V;
  float weight;
  int block_begin, bps;      // first block of the segment in the fused grid, blocks per sample
};
struct SegTable {
  Seg s[kMaxSeg];
  int n, B;
};
struct FinalizeArgs {
  float smooth, w_ce, w_dice;
  int do_bg, batch_dice;
};

__device__ __forceinline__ void unpack2(uint32_t lo, uint32_t hi, float* z) {
  z[0] = __uint_as_float(lo << 16); z[1] = __uint_as_float(lo & 0xffff0000u);
  z[2] = __uint_as_float(hi << 16); z[3] = __uint_as_float(hi & 0xffff0000u);
}
__device__ __forceinline__ float softmax4(float* z) {   // in place; returns the log-sum-exp
  const float m = fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3]));
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c) { z[c] = __expf(z[c] - m); s += z[c]; }
  const float inv = __fdividef(1.f, s);
#pragma unroll
  for (int c = 0; c < 4; ++c) z[c] *= inv;
  return m + logf(s);
}

// software pipeline over `iters` steps through a private per-thread cp.async ring (see losses.cu: staged_sweep)
template <typename Issue, typename Body>
__device__ __forceinline__ void sweep(long long iters, Issue issue_step, Body body) {
  auto issue = [&](long long i) {
    if (i < iters) issue_step(i, (int)(i & (kStage - 1)));
    cp_async_commit();
  };
  for (int i = 0; i < kStage - 1; ++i) issue(i);
  for (long long i = 0; i < iters; ++i) {
    issue(i + kStage - 1);
    cp_async_wait<kStage - 1>();
    body(i, (int)(i & (kStage - 1)));
  }
  cp_async_wait<0>();
}

__device__ __forceinline__ int find_segment(const SegTable& T, int block) {
  int si = 0;
#pragma unroll 1
  for (int i = 1; i < T.n; ++i)
    if (block >= T.s[i].block_begin) si = i;
  return si;
}

// acc layout per segment: [B][4][3] (intersect, sum_pred, sum_gt) + 2 (sum of -log p[target], number of valid voxels);
//   stride acc_stride doubles.  A voxel whose target is outside [0, 4) is IGNORED (the reference's ignore_label, which
//   nnU-Net places behind the last class): it contributes to none of the sums (loss_mask of MemoryEfficientSoftDiceLoss,
//   ignore_index of the cross entropy) and gets a zero gradient.
// coef layout per segment: [B][4][2] (A, E): d(w_dice * Dice)/dp_vc = A * y_vc - E, + 1 (1 / number of valid voxels)
__device__ void finalize_segment(const double* __restrict__ acc, int B, long long V, const FinalizeArgs& F,
                                 float* __restrict__ coef, double& loss) {
  const int C = C4, c0 = F.do_bg ? 0 : 1, nC = C - c0;
  double dc_sum = 0.0;
  if (F.batch_dice) {
    const double nterms = (double)nC;
    for (int c = 0; c < C; ++c) {
      double I = 0, P = 0, G = 0;
      for (int b = 0; b < B; ++b) {
        I += __ldcg(&acc[(b * C + c) * 3 + 0]);
        P += __ldcg(&acc[(b * C + c) * 3 + 1]);
        G += __ldcg(&acc[(b * C + c) * 3 + 2]);
      }
      const double num = 2.0 * I + F.smooth;
      double den = G + P + F.smooth;
      if (den < 1e-8) den = 1e-8;
      float A = 0.f, E = 0.f;
      if (c >= c0) {
        dc_sum += num / den;
        A = (float)(-(double)F.w_dice * (2.0 / den) / nterms);
        E = (float)(-(double)F.w_dice * (num / (den * den)) / nterms);
      }
      for (int b = 0; b < B; ++b) { coef[(b * C + c) * 2] = A; coef[(b * C + c) * 2 + 1] = E; }
    }
    dc_sum /= nterms;
  } else {
    const double nterms = (double)B * nC;
    for (int b = 0; b < B; ++b)
      for (int c = 0; c < C; ++c) {
        const double I = __ldcg(&acc[(b * C + c) * 3 + 0]), P = __ldcg(&acc[(b * C + c) * 3 + 1]),
                     G = __ldcg(&acc[(b * C + c) * 3 + 2]);
        const double num = 2.0 * I + F.smooth;
        double den = G + P + F.smooth;
        if (den < 1e-8) den = 1e-8;
        float A = 0.f, E = 0.f;
        if (c >= c0) {
          dc_sum += num / den;
          A = (float)(-(double)F.w_dice * (2.0 / den) / nterms);
          E = (float)(-(double)F.w_dice * (num / (den * den)) / nterms);
        }
        coef[(b * C + c) * 2] = A;
        coef[(b * C + c) * 2 + 1] = E;
      }
    dc_sum /= nterms;
  }
  // cross entropy: mean over the VALID voxels (targets outside [0, C) are the reference's ignore label: F.cross_entropy(
  // ignore_index) divides by the number of non-ignored targets; DC_and_CE_loss returns 0 for the term when there are none)
  const double nvalid = __ldcg(&acc[B * C * 3 + 1]);
  const double ce = nvalid > 0.0 ? __ldcg(&acc[B * C * 3]) / nvalid : 0.0;
  coef[B * C * 2] = nvalid > 0.0 ? (float)(1.0 / nvalid) : 0.f;
  (void)V;
  loss = (double)F.w_ce * ce + (double)F.w_dice * (-dc_sum);
}

// all segments: thread s finalises segment s, thread 0 adds the weighted terms up (fixed order: deterministic)
__device__ void finalize_all(const SegTable& T, const FinalizeArgs& F, const double* __restrict__ acc, int acc_stride,
                             float* __restrict__ coef, float* __restrict__ loss_out, double* sh_loss) {
  if ((int)threadIdx.x < T.n) {
    double l = 0.0;
    finalize_segment(acc + (long long)threadIdx.x * acc_stride, T.B, T.s[threadIdx.x].V, F,
                     coef + (long long)threadIdx.x * (T.B * C4 * 2 + 1), l);
    sh_loss[threadIdx.x] = (double)T.s[threadIdx.x].weight * l;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < T.n; ++i) tot += sh_loss[i];
    loss_out[0] = (float)tot;
  }
}

__global__ void __launch_bounds__(kThreads) dice_ce_multi_fwd_kernel(const __grid_constant__ SegTable T,
                                                                     const __grid_constant__ FinalizeArgs F,
                                                                     double* __restrict__ acc, int acc_stride,
                                                                     float* __restrict__ coef,
                                                                     float* __restrict__ loss_out,
                                                                     unsigned* __restrict__ counter) {
  constexpr int NS = 3;
  extern __shared__ __align__(16) uint4 ring4[];
  __shared__ float red[kThreads / 32][3 * C4 + 2];
  __shared__ double sh_loss[kMaxSeg];
  __shared__ bool is_last;
  const int si = find_segment(T, blockIdx.x);
  const Seg& S = T.s[si];
  const int local = blockIdx.x - S.block_begin;
  const int b = local / S.bps, blk = local - b * S.bps;
  const uint4* lq = reinterpret_cast<const uint4*>(S.logits + (long long)b * S.V * C4);
  const uint4* tq = reinterpret_cast<const uint4*>(S.target + (long long)b * S.V);
  const long long Q = S.V >> 2, step = (long long)S.bps * kThreads;
  const long long q0 = (long long)blk * kThreads + threadIdx.x;
  const long long iters = q0 < Q ? (Q - q0 + step - 1) / step : 0;
  const uint4* mine = ring4 + threadIdx.x;
  const uint32_t mine_u = (uint32_t)__cvta_generic_to_shared(mine);
  float vals[3 * C4 + 2];
#pragma unroll
  for (int i = 0; i < 3 * C4 + 2; ++i) vals[i] = 0.f;
  sweep(
      iters,
      [&](long long i, int st) {
        const long long q = q0 + i * step;
        const uint32_t d = mine_u + (uint32_t)(st * NS) * (kThreads * 16);
        cp_async16(d, lq + 2 * q);
        cp_async16(d + kThreads * 16, lq + 2 * q + 1);
        cp_async16(d + 2 * kThreads * 16, tq + q);
      },
      [&](long long, int st) {
        const uint4 l0 = mine[(st * NS) * kThreads], l1 = mine[(st * NS + 1) * kThreads];
        const uint4 tv = mine[(st * NS + 2) * kThreads];
        const uint32_t lw[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
        const float tf[4] = {__uint_as_float(tv.x), __uint_as_float(tv.y), __uint_as_float(tv.z), __uint_as_float(tv.w)};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float z[C4];
          unpack2(lw[2 * j], lw[2 * j + 1], z);
          const int t = (int)tf[j];
          float zt = 0.f;
#pragma unroll
          for (int c = 0; c < C4; ++c) zt = (c == t) ? z[c] : zt;
          const float lse = softmax4(z);
          const bool valid = t >= 0 && t < C4;
#pragma unroll
          for (int c = 0; c < C4; ++c) {
            const float y = (c == t) ? 1.f : 0.f;
            vals[3 * c + 0] += z[c] * y;
            vals[3 * c + 1] += valid ? z[c] : 0.f;
            vals[3 * c + 2] += y;
          }
          vals[3 * C4] += valid ? (lse - zt) : 0.f;
          vals[3 * C4 + 1] += valid ? 1.f : 0.f;
        }
      });
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 3 * C4 + 2; ++i) {
    const float sum = warp_sum(vals[i]);
    if (lane == 0) red[warp][i] = sum;
  }
  __syncthreads();
  double* a = acc + (long long)si * acc_stride;
  if (threadIdx.x < 3 * C4 + 2) {
    double v = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) v += (double)red[w][threadIdx.x];
    if (threadIdx.x < 3 * C4) atomicAdd(&a[(long long)b * C4 * 3 + threadIdx.x], v);
    else atomicAdd(&a[(long long)T.B * C4 * 3 + (threadIdx.x - 3 * C4)], v);
  }
  if (counter == nullptr) return;
  // ---- last block done: scalar algebra of all segments (threadfence reduction pattern)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicAdd(counter, 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  finalize_all(T, F, acc, acc_stride, coef, loss_out, sh_loss);
  if (threadIdx.x == 0) *counter = 0u;     // ready for the next launch / graph replay
}

__global__ void dice_ce_multi_finalize_kernel(const __grid_constant__ SegTable T, const __grid_constant__ FinalizeArgs F,
                                              const double* __restrict__ acc, int acc_stride,
                                              float* __restrict__ coef, float* __restrict__ loss_out) {
  __shared__ double sh_loss[kMaxSeg];
  finalize_all(T, F, acc, acc_stride, coef, loss_out, sh_loss);
}

__global__ void __launch_bounds__(kThreads) dice_ce_multi_bwd_kernel(const __grid_constant__ SegTable T,
                                                                     const float* __restrict__ coef, float w_ce,
                                                                     float coef_scale,
                                                                     const float* __restrict__ gout) {
  constexpr int NS = 3;
  extern __shared__ __align__(16) uint4 ring4[];
  const int si = find_segment(T, blockIdx.x);
  const Seg& S = T.s[si];
  const int local = blockIdx.x - S.block_begin;
  const int b = local / S.bps, blk = local - b * S.bps;
  const uint4* lq = reinterpret_cast<const uint4*>(S.logits + (long long)b * S.V * C4);
  const uint4* tq = reinterpret_cast<const uint4*>(S.target + (long long)b * S.V);
  uint4* dq = reinterpret_cast<uint4*>(S.dlogits + (long long)b * S.V * C4);
  float A[C4], E[C4];
  const float* cseg = coef + (long long)si * (T.B * C4 * 2 + 1);
  const float* cf = cseg + (long long)b * C4 * 2;
#pragma unroll
  for (int c = 0; c < C4; ++c) { A[c] = cf[2 * c] * coef_scale; E[c] = cf[2 * c + 1] * coef_scale; }
  const float g = (gout ? gout[0] : 1.f) * S.weight;
  const float ce_scale = w_ce * cseg[T.B * C4 * 2];        // 1 / number of valid voxels of the segment
  const long long Q = S.V >> 2, step = (long long)S.bps * kThreads;
  const long long q0 = (long long)blk * kThreads + threadIdx.x;
  const long long iters = q0 < Q ? (Q - q0 + step - 1) / step : 0;
  const uint4* mine = ring4 + threadIdx.x;
  const uint32_t mine_u = (uint32_t)__cvta_generic_to_shared(mine);
  sweep(
      iters,
      [&](long long i, int st) {
        const long long q = q0 + i * step;
        const uint32_t d = mine_u + (uint32_t)(st * NS) * (kThreads * 16);
        cp_async16(d, lq + 2 * q);
        cp_async16(d + kThreads * 16, lq + 2 * q + 1);
        cp_async16(d + 2 * kThreads * 16, tq + q);
      },
      [&](long long i, int st) {
        const uint4 l0 = mine[(st * NS) * kThreads], l1 = mine[(st * NS + 1) * kThreads];
        const uint4 tv = mine[(st * NS + 2) * kThreads];
        const uint32_t lw[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
        const float tf[4] = {__uint_as_float(tv.x), __uint_as_float(tv.y), __uint_as_float(tv.z), __uint_as_float(tv.w)};
        uint32_t ow[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float p[C4];
          unpack2(lw[2 * j], lw[2 * j + 1], p);
          const int t = (int)tf[j];
          softmax4(p);
          float qv[C4], dot = 0.f;
#pragma unroll
          for (int c = 0; c < C4; ++c) {
            qv[c] = ((c == t) ? A[c] : 0.f) - E[c];
            dot = fmaf(p[c], qv[c], dot);
          }
          float o[C4];
          const bool valid = t >= 0 && t < C4;      // ignored voxels: no gradient
#pragma unroll
          for (int c = 0; c < C4; ++c) {
            const float dce = (p[c] - ((c == t) ? 1.f : 0.f)) * ce_scale;
            o[c] = valid ? g * (dce + p[c] * (qv[c] - dot)) : 0.f;
          }
          __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]), bb = __floats2bfloat162_rn(o[2], o[3]);
          ow[2 * j] = *reinterpret_cast<uint32_t*>(&a);
          ow[2 * j + 1] = *reinterpret_cast<uint32_t*>(&bb);
        }
        const long long q = q0 + i * step;
        dq[2 * q] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        dq[2 * q + 1] = make_uint4(ow[4], ow[5], ow[6], ow[7]);
      });
}

// ---- distillation KL, one pass: loss sum + both gradients ------------------------------------------------------
__global__ void __launch_bounds__(kThreads) kl_fused_kernel(const bf16* __restrict__ ys, const bf16* __restrict__ yt,
                                                            long long NV, float invT, float gscale,
                                                            double* __restrict__ loss_sum, bf16* __restrict__ dys,
                                                            bf16* __restrict__ dyt) {
  constexpr int NS = 4;
  extern __shared__ __align__(16) uint4 ring4[];
  __shared__ float red[kThreads / 32];
  const uint4* sq = reinterpret_cast<const uint4*>(ys);
  const uint4* tq = reinterpret_cast<const uint4*>(yt);
  const long long Q = NV >> 2, step = (long long)gridDim.x * kThreads;
  const long long q0 = (long long)blockIdx.x * kThreads + threadIdx.x;
  const long long iters = q0 < Q ? (Q - q0 + step - 1) / step : 0;
  const uint4* mine = ring4 + threadIdx.x;
  const uint32_t mine_u = (uint32_t)__cvta_generic_to_shared(mine);
  const float g = gscale * invT;
  float acc = 0.f;
  sweep(
      iters,
      [&](long long i, int st) {
        const long long q = q0 + i * step;
        const uint32_t d = mine_u + (uint32_t)(st * NS) * (kThreads * 16);
        cp_async16(d, sq + 2 * q);
        cp_async16(d + kThreads * 16, sq + 2 * q + 1);
        cp_async16(d + 2 * kThreads * 16, tq + 2 * q);
        cp_async16(d + 3 * kThreads * 16, tq + 2 * q + 1);
      },
      [&](long long i, int st) {
        const uint4 s0 = mine[(st * NS) * kThreads], s1 = mine[(st * NS + 1) * kThreads];
        const uint4 t0 = mine[(st * NS + 2) * kThreads], t1 = mine[(st * NS + 3) * kThreads];
        const uint32_t sw[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const uint32_t tw[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
        uint32_t os[8], ot[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float us[C4], ut[C4], ps[C4], pt[C4];
          unpack2(sw[2 * j], sw[2 * j + 1], us);
          unpack2(tw[2 * j], tw[2 * j + 1], ut);
#pragma unroll
          for (int c = 0; c < C4; ++c) { us[c] *= invT; ut[c] *= invT; ps[c] = us[c]; pt[c] = ut[c]; }
          const float lse_s = softmax4(ps);
          const float lse_t = softmax4(pt);
          float d[C4], dot = 0.f;
#pragma unroll
          for (int c = 0; c < C4; ++c) {
            d[c] = (ut[c] - lse_t) - (us[c] - lse_s);
            const float term = pt[c] * d[c];
            dot += term;
            acc += pt[c] > 0.f ? term : 0.f;
          }
          float o[C4];
#pragma unroll
          for (int c = 0; c < C4; ++c) o[c] = g * (ps[c] - pt[c]);
          __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]), b = __floats2bfloat162_rn(o[2], o[3]);
          os[2 * j] = *reinterpret_cast<uint32_t*>(&a); os[2 * j + 1] = *reinterpret_cast<uint32_t*>(&b);
#pragma unroll
          for (int c = 0; c < C4; ++c) o[c] = g * pt[c] * (d[c] - dot);
          a = __floats2bfloat162_rn(o[0], o[1]); b = __floats2bfloat162_rn(o[2], o[3]);
          ot[2 * j] = *reinterpret_cast<uint32_t*>(&a); ot[2 * j + 1] = *reinterpret_cast<uint32_t*>(&b);
        }
        const long long q = q0 + i * step;
        if (dys) {
          uint4* o4 = reinterpret_cast<uint4*>(dys) + 2 * q;
          o4[0] = make_uint4(os[0], os[1], os[2], os[3]);
          o4[1] = make_uint4(os[4], os[5], os[6], os[7]);
        }
        if (dyt) {
          uint4* o4 = reinterpret_cast<uint4*>(dyt) + 2 * q;
          o4[0] = make_uint4(ot[0], ot[1], ot[2], ot[3]);
          o4[1] = make_uint4(ot[4], ot[5], ot[6], ot[7]);
        }
      });
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float s = warp_sum(acc);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) v += (double)red[w];
    atomicAdd(loss_sum, v);
  }
}

// g <- g * (gout / assumed) for up to two bf16 tensors of n8 16-byte vectors each; every block leaves at once when the
// upstream gradient is the assumed one (the case the trainer arranges), so the launch costs only its latency
__global__ void __launch_bounds__(256) rescale_bf16_kernel(uint4* __restrict__ a, uint4* __restrict__ b, long long n8,
                                                           const float* __restrict__ gout, float assumed) {
  const float r = gout[0] / assumed;
  if (r == 1.f) return;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      uint4* p = k ? b : a;
      if (!p) continue;
      bf16x8 v = *reinterpret_cast<const bf16x8*>(&p[i]);
      float f[8];
      unpack8(v, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= r;
      stg16(&p[i], pack8(f));
    }
  }
}

template <typename K>
int resident_blocks(K kernel, size_t smem, bool& attr_done) {
  if (!attr_done) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess) {
      (void)cudaGetLastError();
      return 0;
    }
    attr_done = true;
  }
  int bps = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kernel, kThreads, smem) != cudaSuccess || bps < 1) {
    (void)cudaGetLastError();
    bps = 2;
  }
  return num_sms() * bps;
}

// host: validate the table, deal the resident blocks out to the segments in proportion to their voxel counts
int build_table(const mvd_dice_ce_segment* segs, int n_seg, int B, bool need_dlogits, int budget, SegTable& T,
                const char* who) {
  MVD_REQUIRE(segs && n_seg >= 1 && n_seg <= kMaxSeg && B >= 1, "%s: 1..%d segments", who, kMaxSeg);
  memset(&T, 0, sizeof(T));
  T.n = n_seg;
  T.B = B;
  long long quads = 0;
  for (int i = 0; i < n_seg; ++i) {
    const mvd_dice_ce_segment& s = segs[i];
    MVD_REQUIRE(s.logits && s.target && s.V > 0 && (s.V & 3) == 0 && ((uintptr_t)s.logits & 15) == 0 &&
                    ((uintptr_t)s.target & 15) == 0 && ((B == 1) || ((s.V * 8) % 16 == 0)),
                "%s: segment %d is not the dense C = 4 shape (V %% 4 == 0, 16-byte aligned)", who, i);
    MVD_REQUIRE(!need_dlogits || (s.dlogits && ((uintptr_t)s.dlogits & 15) == 0), "%s: segment %d: bad dlogits", who, i);
    quads += (long long)B * (s.V >> 2);
  }
  int blocks = 0;
  for (int i = 0; i < n_seg; ++i) {
    const long long q = segs[i].V >> 2;                          // quads per sample
    long long bps = (long long)((double)budget * (double)q / (double)quads + 0.5);
    const long long need = (q + kThreads - 1) / kThreads;
    if (bps > need) bps = need;
    if (bps < 1) bps = 1;
    T.s[i].logits = (const bf16*)segs[i].logits;
    T.s[i].target = segs[i].target;
    T.s[i].dlogits = (bf16*)segs[i].dlogits;
    T.s[i].V = segs[i].V;
    T.s[i].weight = segs[i].weight;
    T.s[i].block_begin = blocks;
    T.s[i].bps = (int)bps;
    blocks += (int)bps * B;
  }
  return blocks;
}

}  // namespace
}  // namespace mvd

using namespace mvd;

extern "C" {

int mvd_dice_ce_multi_fwd(const mvd_dice_ce_segment* segs, int n_seg, int B, int C, float smooth, int do_bg,
                          int batch_dice, float w_ce, float w_dice, double* acc, float* coef, float* loss_out,
                          unsigned* counter, mvd_stream_t stream) {
  MVD_REQUIRE(C == C4, "dice_ce_multi_fwd: C must be 4 (got %d); other shapes go through mvd_dice_ce_fwd per scale", C);
  MVD_REQUIRE(acc && (counter == nullptr || (coef && loss_out)), "dice_ce_multi_fwd: null buffers");
  constexpr size_t smem = (size_t)kStage * 3 * kThreads * 16;
  static bool attr = false;
  const int budget = resident_blocks(dice_ce_multi_fwd_kernel, smem, attr);
  MVD_REQUIRE(budget > 0, "dice_ce_multi_fwd: cannot configure the kernel");
  SegTable T;
  const int blocks = build_table(segs, n_seg, B, false, budget, T, "dice_ce_multi_fwd");
  if (blocks < 0) return blocks;
  FinalizeArgs F{smooth, w_ce, w_dice, do_bg, batch_dice};
  dice_ce_multi_fwd_kernel<<<blocks, kThreads, smem, (cudaStream_t)stream>>>(T, F, acc, B * C4 * 3 + 2, coef, loss_out,
                                                                             counter);
  MVD_LAUNCH_CHECK("dice_ce_multi_fwd");
  return MVD_OK;
}

int mvd_dice_ce_multi_finalize(const mvd_dice_ce_segment* segs, int n_seg, int B, int C, float smooth, int do_bg,
                               int batch_dice, float w_ce, float w_dice, const double* acc, float* coef,
                               float* loss_out, mvd_stream_t stream) {
  MVD_REQUIRE(C == C4 && acc && coef && loss_out, "dice_ce_multi_finalize: bad arguments");
  SegTable T;
  const int blocks = build_table(segs, n_seg, B, false, 1, T, "dice_ce_multi_finalize");
  if (blocks < 0) return blocks;
  FinalizeArgs F{smooth, w_ce, w_dice, do_bg, batch_dice};
  dice_ce_multi_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(T, F, acc, B * C4 * 3 + 2, coef, loss_out);
  MVD_LAUNCH_CHECK("dice_ce_multi_finalize");
  return MVD_OK;
}

int mvd_dice_ce_multi_bwd(const mvd_dice_ce_segment* segs, int n_seg, int B, int C, const float* coef, float w_ce,
                          float coef_scale, const float* gout, mvd_stream_t stream) {
  MVD_REQUIRE(C == C4 && coef, "dice_ce_multi_bwd: bad arguments");
  constexpr size_t smem = (size_t)kStage * 3 * kThreads * 16;
  static bool attr = false;
  const int budget = resident_blocks(dice_ce_multi_bwd_kernel, smem, attr);
  MVD_REQUIRE(budget > 0, "dice_ce_multi_bwd: cannot configure the kernel");
  SegTable T;
  const int blocks = build_table(segs, n_seg, B, true, budget, T, "dice_ce_multi_bwd");
  if (blocks < 0) return blocks;
  dice_ce_multi_bwd_kernel<<<blocks, kThreads, smem, (cudaStream_t)stream>>>(T, coef, w_ce, coef_scale, gout);
  MVD_LAUNCH_CHECK("dice_ce_multi_bwd");
  return MVD_OK;
}

int mvd_kl_fused(const void* ys, const void* yt, long long NV, int C, float T, float gscale, double* loss_sum,
                 void* dys, void* dyt, mvd_stream_t stream) {
  MVD_REQUIRE(ys && yt && loss_sum && NV > 0 && T > 0.f, "kl_fused: bad arguments");
  MVD_REQUIRE(C == C4 && (NV & 3) == 0 && (((uintptr_t)ys | (uintptr_t)yt | (uintptr_t)dys | (uintptr_t)dyt) & 15) == 0,
              "kl_fused: dense C = 4 logits with NV %% 4 == 0 only (other shapes: mvd_kl_fwd / mvd_kl_bwd)");
  constexpr size_t smem = (size_t)kStage * 4 * kThreads * 16;
  static bool attr = false;
  int nb = resident_blocks(kl_fused_kernel, smem, attr);
  MVD_REQUIRE(nb > 0, "kl_fused: cannot configure the kernel");
  const long long need = ((NV >> 2) + kThreads - 1) / kThreads;
  if (nb > need) nb = (int)need;
  kl_fused_kernel<<<nb, kThreads, smem, (cudaStream_t)stream>>>((const bf16*)ys, (const bf16*)yt, NV, 1.f / T, gscale,
                                                                loss_sum, (bf16*)dys, (bf16*)dyt);
  MVD_LAUNCH_CHECK("kl_fused");
  return MVD_OK;
}

int mvd_rescale_bf16_pair(void* a, void* b, long long n_elems, const float* gout, float assumed, mvd_stream_t stream) {
  MVD_REQUIRE((a || b) && gout && n_elems > 0 && (n_elems & 7) == 0 && assumed != 0.f &&
                  (((uintptr_t)a | (uintptr_t)b) & 15) == 0, "rescale_bf16_pair: bad arguments");
  const long long n8 = n_elems >> 3;
  rescale_bf16_kernel<<<grid_for(n8, 256, num_sms() * 8), 256, 0, (cudaStream_t)stream>>>((uint4*)a, (uint4*)b, n8, gout,
                                                                                         assumed);
  MVD_LAUNCH_CHECK("rescale_bf16_pair");
  return MVD_OK;
}

}  // extern "C"
