// conv_tc_halo.cu -- tcgen05 implicit GEMM for 3x3x3 / stride 1 / pad 1 convolutions (fprop, and dgrad via flipped
// taps) with SHARED-MEMORY HALO REUSE.  These layers carry ~90 % of the network's FLOPs.
//
// The tap-by-tap kernel (conv_tc.cu) re-fetches every input voxel 27 times through TMA; ncu shows it bound by the
// TMA request rate (~0.27 rows of <=128 B per clock per SM), not by the tensor pipe.  Here an input plane
// [18 h][10 w][KC] (the 16 x 8 output brick plus a one-voxel halo) is loaded ONCE per channel chunk and feeds all nine
// in-plane taps: the UMMA shared-memory descriptor is simply started (oy*10 + ox) rows further, with the stride
// between 8-row groups (SBO) set to the plane's row pitch of 10 voxels.  This relies on the tensor core applying the
// 128B/64B swizzle as a pure function of the shared-memory address (verified by scripts/umma_probe.py,
// profiles/r1_umma_descriptor_probe.txt).  A CTA additionally produces MT output planes (consecutive d) per work
// item from MT+2 input planes, so each plane and each weight tile serves up to MT x 9 (x3) MMAs:
//   TMA rows per 128-voxel tile: 27*(128+N)  ->  ((MT+2)*180 + 27*N) / MT.
// Roles (224 threads): warp 0 plane producer, warp 6 weight producer, warp 1 MMA issuer, warps 2..5 epilogue.
// Accumulators: 2 (double buffer) x MT x N fp32 columns of TMEM.
#include "conv_common.cuh"
#include "tc_common.cuh"
#include "tc_epilogue.cuh"

namespace mvd {
namespace {

using namespace tc;

constexpr int kThreads = 224;
constexpr size_t kFoldBudget = 215 * 1024;   // dynamic shared memory of the depth-folded kernel (weights + plane ring)
constexpr int TILE_W = 8, TILE_H = 16, HALO_W = TILE_W + 2, HALO_H = TILE_H + 2, PLANE_ROWS = HALO_W * HALO_H;
constexpr int kMaxRing = 8, kMaxWStages = 16;
constexpr int kMaxN = 1024;   // largest produced-channel count the tap-major kernel stages a bias vector for

struct alignas(64) HaloMaps {
  CUtensorMap a;   // (C, W, H, D, B) box (KC, 10, 18, 1, 1)
  CUtensorMap b;   // weights [27*N][K] box (KC, n_tile)
};

struct HaloParams {
  int B, D, H, W;
  int tiles_w, tiles_h, dgroups, num_n_tiles, total_items;
  int MT, n_tile, kchunks, ring, wstages;
  uint32_t idesc, tmem_cols;
  bf16* out;
  long long sb, sd, sh, sw;
  const float* bias;
  int accumulate;
  double* stats;     // optional [B][N][2] running (sum, sumsq) of the bf16-rounded output (InstanceNorm statistics)
  int Ntot;
  long long* prof;   // optional [gridDim.x][8] cycle counters of the MMA issuer (debug / DESIGN.md evidence)
  int wrow[27];
  int lane_own;      // epilogue stores each lane's own accumulator row (no shared-memory transpose), see tc_epilogue.cuh
  // data gradient of the depth-folded kernel only: backward statistics of the InstanceNorm + LeakyReLU in front of the
  // produced tensor (mvd_norm_bwd_stats_args); nb_out == nullptr switches it off
  const bf16* nb_y;
  int nb_ld;
  const double* nb_stats;
  const float *nb_gamma, *nb_beta;
  float nb_eps, nb_slope;
  double* nb_out;
};

// work item -> (N tile, sample, first output plane, brick origin)
__device__ __forceinline__ void halo_decode(const HaloParams& P, int MT, int item, int& n0, int& b, int& d0, int& h0,
                                            int& w0) {
  const int nt = item % P.num_n_tiles;
  int m = item / P.num_n_tiles;
  n0 = nt * P.n_tile;
  w0 = (m % P.tiles_w) * TILE_W; m /= P.tiles_w;
  h0 = (m % P.tiles_h) * TILE_H; m /= P.tiles_h;
  d0 = (m % P.dgroups) * MT;
  b = m / P.dgroups;
}

// epilogue warps (4 warps, q = TMEM lane quarter): accumulators [acc][t][n_tile] -> bias, InstanceNorm sums, bf16,
// coalesced stores.  Shared by the tap-major and the depth-folded kernels.
//  * bias comes pre-rounded from shared memory (sbias[0..Ntot));
//  * one packed conversion (F2FP, ALU pipe) per column pair serves both the store and the statistics;
//  * InstanceNorm sums (fused for n_tile <= 64): every lane keeps fp32 partial sums of ITS accumulator row for all
//    columns across the CTA's work items; the cross-lane reduction + fp64 atomics run once per (sample, CTA).
template <int MT, int SC>
__device__ __forceinline__ void halo_epilogue_sc(const HaloParams& P, uint32_t tmem_base, uint64_t* bar_tfull,
                                              uint64_t* bar_tempty, uint8_t* stage, const float* sbias, int q,
                                              int lane) {
  int acc = 0;
  uint32_t accphase = 0;
  constexpr bool do_stats = SC > 0;           // host guarantees n_tile == Ntot == 32 * SC when statistics are fused
  LaneStats<SC> hs;
  hs.reset(-1);
  // 32-channel layers: the bias lives in registers (the shared-memory pipe belongs to the tensor core's operands)
  const bool reg_bias = P.bias != nullptr && P.n_tile == 32 && P.Ntot == 32;
  float2 breg[16];
  if (reg_bias) {
#pragma unroll
    for (int j = 0; j < 16; ++j) breg[j] = *reinterpret_cast<const float2*>(sbias + 2 * j);
  }
  for (int item = blockIdx.x; item < P.total_items; item += gridDim.x) {
    int n0, b, d0, h0, w0;
    halo_decode(P, MT, item, n0, b, d0, h0, w0);
    if (do_stats && b != hs.b) {
      if (hs.b >= 0) hs.flush(P.stats, P.Ntot, lane);
      hs.reset(b);
    }
    mbar_wait(&bar_tfull[acc], accphase, 36);
    tcgen05_fence_after();
    const int H = P.H, W = P.W;
    const long long sh = P.sh, sw = P.sw;
    const int rr = q * 32 + lane;
    const bool ok = (h0 + (rr >> 3) < H) && (w0 + (rr & 7) < W);
    for (int t = 0; t < MT; ++t) {
      const int d = d0 + t;
      if (d >= P.D) break;   // uniform across the CTA
      bf16* tile_base = P.out + (long long)b * P.sb + (long long)d * P.sd + n0;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * MT + t) * P.n_tile);
      for (int c = 0; c < P.n_tile; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + (uint32_t)c, v);
        tmem_ld_wait();
        uint32_t w2[16];
        epilogue_chunk<SC>(v, P.bias ? sbias + n0 + c : nullptr, hs, c, ok, w2, reg_bias ? breg : nullptr);
        if (P.lane_own) {
          if (ok)
            store_row_lane_own(tile_base + (long long)(h0 + (rr >> 3)) * sh + (long long)(w0 + (rr & 7)) * sw + c, w2,
                               P.accumulate != 0);
        } else {
          store_rows_coalesced_packed(stage, lane, w2, [&](int R) -> bf16* {
            const int r2 = q * 32 + R;
            const int h = h0 + (r2 >> 3), w = w0 + (r2 & 7);
            return (h < H && w < W) ? tile_base + (long long)h * sh + (long long)w * sw + c : nullptr;
          }, P.accumulate != 0);
        }
      }
    }
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar_tempty[acc]);
    acc ^= 1;
    if (acc == 0) accphase ^= 1;
  }
  if (do_stats && hs.b >= 0) hs.flush(P.stats, P.Ntot, lane);
}

// ---- data-gradient epilogue (N = 32) that also leaves the backward statistics of the InstanceNorm + LeakyReLU in front
// of the produced tensor (csrc/norm_act.cu: inorm_lrelu_bwd_stats_kernel), from the values it stores:
//   g' = g * (gamma*xhat + beta > 0 ? 1 : slope),   S1 = sum g',   S2 = sum g' xhat,   xhat = (y - mean) * rstd.
// Everything stays in the lane-own layout (lane = voxel, 32 channels in registers): the raw conv outputs y of the lane's
// voxel are 64 contiguous bytes, the LeakyReLU mask is a comparison of y with a per-channel threshold (the sign of
// gamma*rstd folded in by flipping the signs of y and of the threshold), and every lane keeps fp32 partial sums of g' and
// g'*(+-y) for all 32 channels across the CTA's work items; cross-lane reduction, the affine map sum g'y -> sum g'xhat
// and the fp64 atomics run once per (sample, CTA).  No shared memory in the item loop (see store_row_lane_own).
struct NbTables {
  float thr[32];       // threshold of (+-y) above which the pre-activation is positive (+-inf for a constant mask)
  uint32_t neg;        // bit c: gamma*rstd < 0, i.e. the comparison runs on -y
};

template <int MT>
__device__ __forceinline__ void halo_epilogue_nb(const HaloParams& P, uint32_t tmem_base, uint64_t* bar_tfull,
                                                 uint64_t* bar_tempty, const float* sbias, NbTables* tab, int q,
                                                 int lane) {
  int acc = 0;
  uint32_t accphase = 0;
  int cur_b = -1;
  const int H = P.H, W = P.W;
  const long long ysw = P.nb_ld, ysh = ysw * W, ysd = ysh * H;
  const double Vd = (double)P.D * (double)H * (double)W;
  LaneStats<0> none;
  float thr[32], s1[32], s2[32];
  uint32_t neg = 0;
  auto mean_rstd = [&](int b, int c, float& mean, float& rstd) {     // exactly as load_scale_shift (norm_act.cu)
    const double a1 = P.nb_stats[((long long)b * 32 + c) * 2 + 0];
    const double a2 = P.nb_stats[((long long)b * 32 + c) * 2 + 1];
    const double m = a1 / Vd;
    double var = a2 / Vd - m * m;
    if (var < 0.0) var = 0.0;
    rstd = (float)(1.0 / sqrt(var + (double)P.nb_eps));
    mean = (float)m;
  };
  auto flush = [&](int b) {
    // column sums over the warp's 32 lanes: afterwards lane l owns channel l
    const float S1 = warp_column_sum32(s1, lane), A = warp_column_sum32(s2, lane);
    float mean, rstd;
    mean_rstd(b, lane, mean, rstd);
    const double Ay = ((tab->neg >> lane) & 1u) ? -(double)A : (double)A;      // sum g' y
    double* d = P.nb_out + ((long long)b * 32 + lane) * 2;
    atomicAdd(d, (double)S1);
    atomicAdd(d + 1, (double)rstd * (Ay - (double)mean * (double)S1));          // sum g' xhat
  };
  for (int item = blockIdx.x; item < P.total_items; item += gridDim.x) {
    int n0, b, d0, h0, w0;
    halo_decode(P, MT, item, n0, b, d0, h0, w0);
    if (b != cur_b) {       // all four epilogue warps arrive here at the same items
      if (cur_b >= 0) flush(cur_b);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (q == 0) {
        float mean, rstd;
        mean_rstd(b, lane, mean, rstd);
        const float g = P.nb_gamma ? P.nb_gamma[lane] : 1.f;
        const float be = P.nb_beta ? P.nb_beta[lane] : 0.f;
        const float sc = g * rstd, shf = be - mean * g * rstd;      // pre-activation t = y * sc + shf;  mask = t > 0
        float th;
        if (sc > 0.f) th = -shf / sc;                 // y > th
        else if (sc < 0.f) th = shf / sc;             // y < -shf/sc  <=>  -y > shf/sc
        else th = __int_as_float(shf > 0.f ? 0xff800000 : 0x7f800000);   // constant mask: -inf (always) / +inf (never)
        tab->thr[lane] = th;
        const uint32_t negb = __ballot_sync(0xffffffffu, sc < 0.f);
        if (lane == 0) tab->neg = negb;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      neg = tab->neg;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        thr[j] = tab->thr[j];
        s1[j] = s2[j] = 0.f;
      }
      cur_b = b;
    }
    const float slope = P.nb_slope;
    const int rr = q * 32 + lane;
    const int h = h0 + (rr >> 3), w = w0 + (rr & 7);
    const bool ok = h < H && w < W;
    const int mt = min(MT, P.D - d0);
    bf16* out_row = P.out + (long long)b * P.sb + (long long)d0 * P.sd + (long long)h * P.sh + (long long)w * P.sw;
    const bf16* y_row = P.nb_y + ((long long)b * P.D + d0) * ysd + (long long)h * ysh + (long long)w * ysw;
    // the lane's 64 bytes of y for plane t, two planes ahead of the one being reduced (MT is a compile-time constant:
    // the loop below is fully unrolled and the three buffers keep static names)
    uint4 yb[3][4];
    auto load_y = [&](int t, uint4* dst) {
      if (ok && t < mt) {
        const bf16* src = y_row + (long long)t * ysd;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(dst[g].x), "=r"(dst[g].y), "=r"(dst[g].z), "=r"(dst[g].w) : "l"(src + 8 * g) : "memory");
      }
    };
    load_y(0, yb[0]);
    load_y(1, yb[1]);
    mbar_wait(&bar_tfull[acc], accphase, 36);
    tcgen05_fence_after();
#pragma unroll
    for (int t = 0; t < MT; ++t) {
      if (t < mt) {      // uniform across the CTA
        load_y(t + 2, yb[(t + 2) % 3]);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * MT + t) * 32);
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr, v);
        tmem_ld_wait();
        uint32_t w2[16];
        epilogue_chunk<0>(v, P.bias ? sbias : nullptr, none, 0, false, w2);
        if (ok) {
          store_row_lane_own(out_row + (long long)t * P.sd, w2, P.accumulate != 0);
          const uint4* yv = yb[t % 3];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t yw[4] = {yv[g].x, yv[g].y, yv[g].z, yv[g].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int j = 8 * g + 2 * k;          // channels j, j + 1
              const float g0 = __uint_as_float(w2[4 * g + k] << 16), g1 = __uint_as_float(w2[4 * g + k] & 0xffff0000u);
              const float y0 = __uint_as_float((yw[k] << 16) ^ ((neg << (31 - j)) & 0x80000000u));
              const float y1 = __uint_as_float((yw[k] & 0xffff0000u) ^ ((neg << (30 - j)) & 0x80000000u));
              const float p0 = g0 * (y0 > thr[j] ? 1.f : slope);
              const float p1 = g1 * (y1 > thr[j + 1] ? 1.f : slope);
              s1[j] += p0;
              s2[j] = fmaf(p0, y0, s2[j]);
              s1[j + 1] += p1;
              s2[j + 1] = fmaf(p1, y1, s2[j + 1]);
            }
          }
        }
      }
    }
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar_tempty[acc]);
    acc ^= 1;
    if (acc == 0) accphase ^= 1;
  }
  if (cur_b >= 0) flush(cur_b);
}

template <int MT>
__device__ __forceinline__ void halo_epilogue(const HaloParams& P, uint32_t tmem_base, uint64_t* bar_tfull,
                                              uint64_t* bar_tempty, uint8_t* stage, const float* sbias, int q,
                                              int lane) {
  if (!P.stats) halo_epilogue_sc<MT, 0>(P, tmem_base, bar_tfull, bar_tempty, stage, sbias, q, lane);
  else if (P.n_tile <= 32) halo_epilogue_sc<MT, 1>(P, tmem_base, bar_tfull, bar_tempty, stage, sbias, q, lane);
  else halo_epilogue_sc<MT, 2>(P, tmem_base, bar_tfull, bar_tempty, stage, sbias, q, lane);
}

// bias (rounded to bf16 as the reference's autocast conv does) -> shared memory, zero when absent; all threads, once
__device__ __forceinline__ void halo_stage_bias(const HaloParams& P, float* sbias) {
  for (int i = threadIdx.x; i < P.Ntot; i += blockDim.x) sbias[i] = P.bias ? round_bf(__ldg(P.bias + i)) : 0.f;
}

template <int KC, int MT>
__global__ void __launch_bounds__(kThreads, 1) conv_halo_kernel(const __grid_constant__ HaloMaps maps,
                                                                const __grid_constant__ HaloParams P) {
  constexpr int ROWB = KC * 2;
  constexpr int PLANE_TX = PLANE_ROWS * ROWB;                       // bytes TMA delivers per plane
  constexpr int PLANE_BYTES = (PLANE_TX + 1023) & ~1023;            // slot pitch (1 KB aligned)
  constexpr uint64_t LAYOUT = (KC == 64) ? kLayoutSw128 : kLayoutSw64;
  constexpr uint32_t A_SBO = HALO_W * ROWB;                         // next 8-voxel row group = next h line of the plane
  constexpr uint32_t B_SBO = 8 * ROWB;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar_pfull[kMaxRing], bar_pempty[kMaxRing], bar_wfull[kMaxWStages], bar_wempty[kMaxWStages],
      bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(16) uint8_t s_stage[4][2048];   // per epilogue warp: 32 rows x 64 B transpose buffer
  __shared__ __align__(16) float s_bias[kMaxN];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int w_bytes = P.n_tile * ROWB;
  uint8_t* smem_p = smem;
  uint8_t* smem_w = smem + (size_t)P.ring * PLANE_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NP = MT + 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.ring; ++s) { mbar_init(&bar_pfull[s], 1); mbar_init(&bar_pempty[s], 1); }
    for (int s = 0; s < P.wstages; ++s) { mbar_init(&bar_wfull[s], 1); mbar_init(&bar_wempty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&bar_tfull[a], 1); mbar_init(&bar_tempty[a], 4); }
    fence_barrier_init();
  }
  halo_stage_bias(P, s_bias);
  if (warp == 1) tmem_alloc(&s_tmem_base, P.tmem_cols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  auto decode = [&](int item, int& n0, int& b, int& d0, int& h0, int& w0) {
    halo_decode(P, MT, item, n0, b, d0, h0, w0);
  };

  if (warp == 0) {
    if (elect_one_sync()) {
      // ================= plane producer =================
      int slot = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < P.total_items; item += gridDim.x) {
        int n0, b, d0, h0, w0;
        decode(item, n0, b, d0, h0, w0);
        for (int kc = 0; kc < P.kchunks; ++kc)
          for (int p = 0; p < NP; ++p) {
            mbar_wait(&bar_pempty[slot], phase ^ 1, 31);
            mbar_arrive_expect_tx(&bar_pfull[slot], (uint32_t)PLANE_TX);
            tma_load_5d(&maps.a, smem_p + (size_t)slot * PLANE_BYTES, &bar_pfull[slot], kc * KC, w0 - 1, h0 - 1,
                        d0 + p - 1, b);
            if (++slot == P.ring) { slot = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 6) {
    if (elect_one_sync()) {
      // ================= weight producer =================
      int ws = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < P.total_items; item += gridDim.x) {
        const int n0 = (item % P.num_n_tiles) * P.n_tile;
        for (int kc = 0; kc < P.kchunks; ++kc)
          for (int o = 0; o < 27; ++o) {
            mbar_wait(&bar_wempty[ws], phase ^ 1, 32);
            mbar_arrive_expect_tx(&bar_wfull[ws], (uint32_t)w_bytes);
            tma_load_2d(&maps.b, smem_w + (size_t)ws * w_bytes, &bar_wfull[ws], kc * KC, P.wrow[o] + n0);
            if (++ws == P.wstages) { ws = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      // ================= MMA issuer =================
      // everything loop-invariant lives in registers: the single issuing thread has no ILP to hide constant-bank
      // reloads or 64-bit address arithmetic between two tcgen05.mma
      const int ring = P.ring, wstages = P.wstages, kchunks = P.kchunks, n_tile = P.n_tile;
      const int total_items = P.total_items, gstride = gridDim.x;
      const uint32_t idesc = P.idesc;
      const uint32_t a_hi = (uint32_t)(make_smem_desc(0, 16, A_SBO, LAYOUT) >> 32);
      const uint32_t b_hi = (uint32_t)(make_smem_desc(0, 16, B_SBO, LAYOUT) >> 32);
      const uint32_t p_base = (smem_u32(smem_p) >> 4) | (1u << 16);     // descriptor low word of ring slot 0
      const uint32_t w_base = (smem_u32(smem_w) >> 4) | (1u << 16);
      const uint32_t w_step = (uint32_t)w_bytes >> 4;
      int acc = 0, ws = 0;
      uint32_t accphase = 0, wphase = 0;
      int base_slot = 0;          // ring slot of plane 0 of the current (item, chunk)
      uint32_t base_phase = 0;    // its full-barrier parity; slots that wrap past the ring end use the flipped parity
      long long c_tempty = 0, c_wfull = 0, c_pfull = 0;
      const long long c_start = clock64();
      for (int item = blockIdx.x; item < total_items; item += gstride) {
        long long c0 = clock64();
        mbar_wait(&bar_tempty[acc], accphase ^ 1, 33);
        c_tempty += clock64() - c0;
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * MT * n_tile);
        for (int kc = 0; kc < kchunks; ++kc) {
          uint32_t plo[NP];       // descriptor low words of this chunk's planes
          uint32_t ppar = 0;      // bit p: parity of plane p's full barrier
          int pslot[NP];
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            int slot = base_slot + p;
            uint32_t par = base_phase;
            if (slot >= ring) { slot -= ring; par ^= 1; }
            pslot[p] = slot;
            ppar |= par << p;
            plo[p] = p_base + (uint32_t)slot * (uint32_t)(PLANE_BYTES >> 4);
          }
#pragma unroll
          for (int oz = 0; oz < 3; ++oz) {
            for (int oyx = 0; oyx < 9; ++oyx) {
              const int oy = oyx / 3, ox = oyx - oy * 3;
              c0 = clock64();
              mbar_wait(&bar_wfull[ws], wphase, 34);
              c_wfull += clock64() - c0;
              tcgen05_fence_after();
              const uint32_t wlo = w_base + (uint32_t)ws * w_step;
              const uint32_t tap_off = (uint32_t)(((oy * HALO_W + ox) * ROWB) >> 4);
#pragma unroll
              for (int t = 0; t < MT; ++t) {
                const int p = t + oz;
                // a plane is first touched at the first tap of the first oz phase that uses it
                if (oyx == 0 && (oz == 0 || t == MT - 1)) {
                  c0 = clock64();
                  mbar_wait(&bar_pfull[pslot[p]], (ppar >> p) & 1u, 35);
                  c_pfull += clock64() - c0;
                  tcgen05_fence_after();
                }
                const uint64_t adesc = ((uint64_t)a_hi << 32) | (uint64_t)(plo[p] + tap_off);
                const uint64_t bdesc = ((uint64_t)b_hi << 32) | (uint64_t)wlo;
#pragma unroll
                for (int k = 0; k < KC / 16; ++k)
                  umma_bf16(d_tmem + (uint32_t)(t * n_tile), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                            (kc | oz | oyx | k) ? 1u : 0u);
              }
              umma_commit(&bar_wempty[ws]);
              if (++ws == wstages) { ws = 0; wphase ^= 1; }
            }
            // planes whose last use was this oz phase: p with min(p, 2) == oz
            if (oz < 2) {
              umma_commit(&bar_pempty[pslot[oz]]);
            } else {
#pragma unroll
              for (int p = 2; p < NP; ++p) umma_commit(&bar_pempty[pslot[p]]);
            }
          }
          base_slot += NP;
          if (base_slot >= ring) { base_slot -= ring; base_phase ^= 1; }
        }
        umma_commit(&bar_tfull[acc]);
        acc ^= 1;
        if (acc == 0) accphase ^= 1;
      }
      if (P.prof) {
        long long* o = P.prof + (long long)blockIdx.x * 8;
        o[0] = clock64() - c_start; o[1] = c_tempty; o[2] = c_wfull; o[3] = c_pfull;
      }
    }
  } else if (warp >= 2 && warp <= 5) {
    // ================= epilogue =================
    halo_epilogue<MT>(P, tmem_base, bar_tfull, bar_tempty, s_stage[warp & 3], s_bias, warp & 3, lane);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, P.tmem_cols);
}

// ------------------------------------------------------------------------------------------------------------------
// Depth-folded variant for narrow layers (N = 32 / 64) whose 27 weight tiles fit in shared memory.
//
// One tcgen05.mma M128 x N x K16 costs max(N/2, 32 + N/4) cycles (profiles/r1_mma_issue_rate.txt): at N = 32 the 4 KB
// A-operand read, not the tensor pipe, is the limit (40 cycles for 16 cycles of math).  Here the three depth taps are
// folded into the N axis: input plane a of a work item contributes to the output planes t = a - oz (oz = 0..2), whose
// accumulators sit side by side in TMEM ([t][N] columns), so ONE MMA with B = the weight rows [oz = 2 | 1 | 0] x N of
// an in-plane tap (oy, ox) and N' = 3N columns does the work of three (56 instead of 120 cycles at N = 32).  The shift
// along depth is carried by the TMEM column address -- no shift-add epilogue.  At the ends of the MT-plane run the MMA
// narrows to 2N / N columns.
// Loop order is plane-major (for plane a: all 9 in-plane taps x K steps), which needs all weight tiles resident (loaded
// once per CTA) and lets the planes stream through a short ring across work items; MT is limited by TMEM only
// (2 x MT x N <= 512 columns).  The first MMA that touches a new output plane is split off (accumulate = 0).
// ------------------------------------------------------------------------------------------------------------------
template <int KC, int MT, bool NB = false>   // NB: data gradient + norm-backward statistics (halo_epilogue_nb)
__global__ void __launch_bounds__(kThreads, 1) conv_halo_fold_kernel(const __grid_constant__ HaloMaps maps,
                                                                     const __grid_constant__ HaloParams P) {
  constexpr int ROWB = KC * 2;
  constexpr int PLANE_TX = PLANE_ROWS * ROWB;
  constexpr int PLANE_BYTES = (PLANE_TX + 1023) & ~1023;
  constexpr uint64_t LAYOUT = (KC == 64) ? kLayoutSw128 : kLayoutSw64;
  constexpr uint32_t A_SBO = HALO_W * ROWB;
  constexpr uint32_t B_SBO = 8 * ROWB;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar_pfull[kMaxRing], bar_pempty[kMaxRing], bar_wres, bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(16) uint8_t s_stage[4][2048];
  __shared__ __align__(16) float s_bias[64];
  __shared__ NbTables s_nbtab;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int N = P.n_tile;
  const int wtile_bytes = 3 * N * ROWB;                   // one (chunk, in-plane tap): rows [oz=2 | oz=1 | oz=0] x N
  uint8_t* smem_w = smem;
  uint8_t* smem_p = smem + (size_t)P.kchunks * 9 * wtile_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.ring; ++s) { mbar_init(&bar_pfull[s], 1); mbar_init(&bar_pempty[s], 1); }
    mbar_init(&bar_wres, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&bar_tfull[a], 1); mbar_init(&bar_tempty[a], 4); }
    fence_barrier_init();
  }
  halo_stage_bias(P, s_bias);
  if (warp == 1) tmem_alloc(&s_tmem_base, P.tmem_cols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp == 0) {
    if (elect_one_sync()) {
      // ================= plane producer =================
      int slot = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < P.total_items; item += gridDim.x) {
        int n0, b, d0, h0, w0;
        halo_decode(P, MT, item, n0, b, d0, h0, w0);
        const int mt = min(MT, P.D - d0);
        for (int kc = 0; kc < P.kchunks; ++kc)
          for (int a = 0; a < mt + 2; ++a) {
            mbar_wait(&bar_pempty[slot], phase ^ 1, 41);
            mbar_arrive_expect_tx(&bar_pfull[slot], (uint32_t)PLANE_TX);
            tma_load_5d(&maps.a, smem_p + (size_t)slot * PLANE_BYTES, &bar_pfull[slot], kc * KC, w0 - 1, h0 - 1,
                        d0 + a - 1, b);
            if (++slot == P.ring) { slot = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 6) {
    if (elect_one_sync()) {
      // ================= weights: all kchunks x 27 tiles, once =================
      mbar_arrive_expect_tx(&bar_wres, (uint32_t)(P.kchunks * 9 * wtile_bytes));
      for (int kc = 0; kc < P.kchunks; ++kc)
        for (int oyx = 0; oyx < 9; ++oyx)
          for (int j = 0; j < 3; ++j)   // row block j holds depth tap oz = 2 - j
            tma_load_2d(&maps.b, smem_w + (size_t)(kc * 9 + oyx) * wtile_bytes + (size_t)j * N * ROWB, &bar_wres, kc * KC,
                        P.wrow[(2 - j) * 9 + oyx]);
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      // ================= MMA issuer =================
      const int ring = P.ring, kchunks = P.kchunks;
      const int total_items = P.total_items, gstride = gridDim.x, D = P.D;
      const uint32_t a_hi = (uint32_t)(make_smem_desc(0, 16, A_SBO, LAYOUT) >> 32);
      const uint32_t b_hi = (uint32_t)(make_smem_desc(0, 16, B_SBO, LAYOUT) >> 32);
      const uint32_t p_base = (smem_u32(smem_p) >> 4) | (1u << 16);
      const uint32_t w_base = (smem_u32(smem_w) >> 4) | (1u << 16);
      const uint32_t wtile16 = (uint32_t)wtile_bytes >> 4;
      const uint32_t nrow16 = (uint32_t)(N * ROWB) >> 4;
      const uint32_t idesc1 = make_idesc_bf16(128, N, 0, 0), idesc2 = make_idesc_bf16(128, 2 * N, 0, 0),
                     idesc3 = make_idesc_bf16(128, 3 * N, 0, 0);
      int acc = 0, slot = 0;
      uint32_t accphase = 0, pphase = 0;
      long long c_tempty = 0, c_pfull = 0;
      const long long c_start = clock64();
      mbar_wait(&bar_wres, 0, 42);
      tcgen05_fence_after();
      for (int item = blockIdx.x; item < total_items; item += gstride) {
        const int d0 = ((item / (P.tiles_w * P.tiles_h)) % P.dgroups) * MT;   // num_n_tiles == 1 in this kernel
        const int mt = min(MT, D - d0);
        long long c0 = clock64();
        mbar_wait(&bar_tempty[acc], accphase ^ 1, 43);
        c_tempty += clock64() - c0;
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * MT * N);
        for (int kc = 0; kc < kchunks; ++kc) {
          const uint32_t wk = w_base + (uint32_t)(kc * 9) * wtile16;
          for (int a = 0; a < mt + 2; ++a) {
            const int oz_hi = min(2, a), oz_lo = max(0, a - (mt - 1));
            const int cnt = oz_hi - oz_lo + 1;                    // output planes t = a - oz_hi .. a - oz_lo
            const uint32_t d_col = d_tmem + (uint32_t)((a - oz_hi) * N);
            const uint32_t wrow0 = wk + (uint32_t)(2 - oz_hi) * nrow16;   // first weight row block used
            const uint32_t idesc = (cnt == 3) ? idesc3 : ((cnt == 2) ? idesc2 : idesc1);
            const bool fresh = (kc == 0) && (oz_lo == 0);         // plane t = a gets its first contribution here
            c0 = clock64();
            mbar_wait(&bar_pfull[slot], pphase, 44);
            c_pfull += clock64() - c0;
            tcgen05_fence_after();
            const uint32_t plo = p_base + (uint32_t)slot * (uint32_t)(PLANE_BYTES >> 4);
#pragma unroll
            for (int oyx = 0; oyx < 9; ++oyx) {
              const int oy = oyx / 3, ox = oyx - oy * 3;
              const uint32_t tap_off = (uint32_t)(((oy * HALO_W + ox) * ROWB) >> 4);
              const uint64_t adesc = ((uint64_t)a_hi << 32) | (uint64_t)(plo + tap_off);
              const uint64_t bdesc = ((uint64_t)b_hi << 32) | (uint64_t)(wrow0 + (uint32_t)oyx * wtile16);
#pragma unroll
              for (int k = 0; k < KC / 16; ++k) {
                if (oyx == 0 && k == 0 && fresh) {
                  // the newest plane (last N columns / last row block) starts from zero; the older ones accumulate
                  umma_bf16(d_col + (uint32_t)((cnt - 1) * N), adesc, bdesc + (uint64_t)((uint32_t)(cnt - 1) * nrow16),
                            idesc1, 0u);
                  if (cnt > 1) umma_bf16(d_col, adesc, bdesc, (cnt == 3) ? idesc2 : idesc1, 1u);
                } else {
                  umma_bf16(d_col, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
                }
              }
            }
            umma_commit(&bar_pempty[slot]);
            if (++slot == ring) { slot = 0; pphase ^= 1; }
          }
        }
        umma_commit(&bar_tfull[acc]);
        acc ^= 1;
        if (acc == 0) accphase ^= 1;
      }
      if (P.prof) {
        long long* o = P.prof + (long long)blockIdx.x * 8;
        o[0] = clock64() - c_start; o[1] = c_tempty; o[2] = 0; o[3] = c_pfull;
      }
    }
  } else if (warp >= 2 && warp <= 5) {
    if constexpr (NB) halo_epilogue_nb<MT>(P, tmem_base, bar_tfull, bar_tempty, s_bias, &s_nbtab, warp & 3, lane);
    else halo_epilogue<MT>(P, tmem_base, bar_tfull, bar_tempty, s_stage[warp & 3], s_bias, warp & 3, lane);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, P.tmem_cols);
}

// ------------------------------------------------------------------------------------------------------------------
// Stride-2 forward (3x3x3, pad 1, Cin = 32, Cout = 32 / 64: the first down-sampling conv, 128^3 -> 64^3).
//
// The parity-class form in conv_tc.cu fetches one 128-voxel box per tap (27 boxes of 64-byte rows per tile) and is bound
// by the TMA row rate (0.28 ms, 210 TFLOP/s for 32 -> 64 at 2 x 128^3).  Here every input plane is fetched ONCE, split
// by the parity of (h, w) into four dense sub-planes (one strided tensor map each), so that an in-plane tap is again
// just a descriptor start inside a sub-plane:
//     input h = 2 oy + ty - 1:  ty = 1 -> even rows, line oy;  ty = 0 / 2 -> odd rows (first line = 2 oy0 - 1), line oy / oy + 1
//     (same along w); a sub-plane is [16 | 17 lines][8 | 9 voxels][32 ch], 8-row-group pitch = its line pitch.
// Depth: input plane a of a work item (d_in = 2 od0 - 1 + a, a = 0 .. 2 mt) is odd-numbered (a = 2t + 1) -> centre tap of
// output plane t, or even-numbered (a = 2t) -> tap 0 of output plane t AND tap 2 of output plane t - 1: those two are
// folded into one MMA of N' = 2N whose B operand is the weight rows [tz = 2 | tz = 0] and whose accumulator columns are
// the adjacent TMEM blocks of planes t - 1 and t (as in conv_halo_fold_kernel).  All 27 weight tiles stay resident.
// Per 128-voxel output tile: 18 MMAs of N' = 128 + 18 of N = 64 (2,016 cycles) against 1,122 TMA rows (~4,150 cycles).
// ------------------------------------------------------------------------------------------------------------------
constexpr int S2_KC = 32, S2_ROWB = S2_KC * 2, S2_MT = 4, S2_RING = 3;
constexpr int S2_NH0 = TILE_H, S2_NH1 = TILE_H + 1, S2_NW0 = TILE_W, S2_NW1 = TILE_W + 1;
constexpr int s2_align1k(int v) { return (v + 1023) & ~1023; }
constexpr int S2_SZ00 = S2_NH0 * S2_NW0 * S2_ROWB, S2_SZ01 = S2_NH0 * S2_NW1 * S2_ROWB,
              S2_SZ10 = S2_NH1 * S2_NW0 * S2_ROWB, S2_SZ11 = S2_NH1 * S2_NW1 * S2_ROWB;
constexpr int S2_OFF00 = 0, S2_OFF01 = s2_align1k(S2_OFF00 + S2_SZ00), S2_OFF10 = s2_align1k(S2_OFF01 + S2_SZ01),
              S2_OFF11 = s2_align1k(S2_OFF10 + S2_SZ10), S2_SLOT_BYTES = s2_align1k(S2_OFF11 + S2_SZ11);
constexpr int S2_PLANE_TX = S2_SZ00 + S2_SZ01 + S2_SZ10 + S2_SZ11;

struct alignas(64) S2Maps {
  CUtensorMap a[4];   // [h parity * 2 + w parity]: (C, W/2, H/2, D, B) with doubled W / H strides
  CUtensorMap b;      // weights [27 * N][32] box (32, N)
};

__global__ void __launch_bounds__(kThreads, 1) conv_halo_s2_kernel(const __grid_constant__ S2Maps maps,
                                                                   const __grid_constant__ HaloParams P) {
  constexpr int MT = S2_MT;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar_pfull[S2_RING], bar_pempty[S2_RING], bar_wres, bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(16) uint8_t s_stage[4][2048];
  __shared__ __align__(16) float s_bias[64];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int N = P.n_tile;
  const int ntile_bytes = N * S2_ROWB;                      // one tap: N rows x 64 B
  uint8_t* smem_we = smem;                                  // 9 x [tz = 2 | tz = 0]
  uint8_t* smem_wo = smem + (size_t)18 * ntile_bytes;       // 9 x [tz = 1]
  uint8_t* smem_p = smem + (size_t)27 * ntile_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S2_RING; ++s) { mbar_init(&bar_pfull[s], 1); mbar_init(&bar_pempty[s], 1); }
    mbar_init(&bar_wres, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&bar_tfull[a], 1); mbar_init(&bar_tempty[a], 4); }
    fence_barrier_init();
  }
  halo_stage_bias(P, s_bias);
  if (warp == 1) tmem_alloc(&s_tmem_base, P.tmem_cols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp == 0) {
    if (elect_one_sync()) {
      // ================= plane producer: four parity sub-planes per input plane, one barrier =================
      int slot = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < P.total_items; item += gridDim.x) {
        int n0, b, d0, h0, w0;
        halo_decode(P, MT, item, n0, b, d0, h0, w0);
        const int mt = min(MT, P.D - d0);
        for (int a = 0; a <= 2 * mt; ++a) {
          mbar_wait(&bar_pempty[slot], phase ^ 1, 71);
          mbar_arrive_expect_tx(&bar_pfull[slot], (uint32_t)S2_PLANE_TX);
          uint8_t* dst = smem_p + (size_t)slot * S2_SLOT_BYTES;
          const int din = 2 * d0 - 1 + a;
          tma_load_5d(&maps.a[0], dst + S2_OFF00, &bar_pfull[slot], 0, w0, h0, din, b);
          tma_load_5d(&maps.a[1], dst + S2_OFF01, &bar_pfull[slot], 0, w0 - 1, h0, din, b);
          tma_load_5d(&maps.a[2], dst + S2_OFF10, &bar_pfull[slot], 0, w0, h0 - 1, din, b);
          tma_load_5d(&maps.a[3], dst + S2_OFF11, &bar_pfull[slot], 0, w0 - 1, h0 - 1, din, b);
          if (++slot == S2_RING) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 6) {
    if (elect_one_sync()) {
      // ================= weights: 27 tiles, once =================
      mbar_arrive_expect_tx(&bar_wres, (uint32_t)(27 * ntile_bytes));
      for (int oyx = 0; oyx < 9; ++oyx) {
        tma_load_2d(&maps.b, smem_we + (size_t)(2 * oyx) * ntile_bytes, &bar_wres, 0, P.wrow[18 + oyx]);
        tma_load_2d(&maps.b, smem_we + (size_t)(2 * oyx + 1) * ntile_bytes, &bar_wres, 0, P.wrow[oyx]);
        tma_load_2d(&maps.b, smem_wo + (size_t)oyx * ntile_bytes, &bar_wres, 0, P.wrow[9 + oyx]);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      // ================= MMA issuer =================
      const int total_items = P.total_items, gstride = gridDim.x, D = P.D;
      const uint32_t a_hi0 = (uint32_t)(make_smem_desc(0, 16, S2_NW0 * S2_ROWB, kLayoutSw64) >> 32);
      const uint32_t a_hi1 = (uint32_t)(make_smem_desc(0, 16, S2_NW1 * S2_ROWB, kLayoutSw64) >> 32);
      const uint32_t b_hi = (uint32_t)(make_smem_desc(0, 16, 8 * S2_ROWB, kLayoutSw64) >> 32);
      const uint32_t p_base = (smem_u32(smem_p) >> 4) | (1u << 16);
      const uint32_t we_base = (smem_u32(smem_we) >> 4) | (1u << 16);
      const uint32_t wo_base = (smem_u32(smem_wo) >> 4) | (1u << 16);
      const uint32_t ntile16 = (uint32_t)ntile_bytes >> 4;
      const uint32_t idesc1 = make_idesc_bf16(128, N, 0, 0), idesc2 = make_idesc_bf16(128, 2 * N, 0, 0);
      int acc = 0, slot = 0;
      uint32_t accphase = 0, pphase = 0;
      mbar_wait(&bar_wres, 0, 72);
      tcgen05_fence_after();
      for (int item = blockIdx.x; item < total_items; item += gstride) {
        const int d0 = ((item / (P.tiles_w * P.tiles_h)) % P.dgroups) * MT;   // num_n_tiles == 1
        const int mt = min(MT, D - d0);
        mbar_wait(&bar_tempty[acc], accphase ^ 1, 73);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * MT * N);
        for (int a = 0; a <= 2 * mt; ++a) {
          // which output planes this input plane feeds, and with which weight rows
          uint32_t d_col, wlo, wstep, idesc;
          bool fresh_hi = false, both = false;
          if (a & 1) {                                   // centre tap of plane t = (a - 1) / 2
            d_col = d_tmem + (uint32_t)((a >> 1) * N);
            wlo = wo_base; wstep = ntile16; idesc = idesc1;
          } else {
            const int t_hi = a >> 1, t_lo = t_hi - 1;    // tap 0 of t_hi, tap 2 of t_lo
            wstep = 2 * ntile16;
            if (t_lo >= 0 && t_hi < mt) { d_col = d_tmem + (uint32_t)(t_lo * N); wlo = we_base; idesc = idesc2; both = true; fresh_hi = true; }
            else if (t_lo < 0) { d_col = d_tmem; wlo = we_base + ntile16; idesc = idesc1; fresh_hi = true; }
            else { d_col = d_tmem + (uint32_t)(t_lo * N); wlo = we_base; idesc = idesc1; }
          }
          mbar_wait(&bar_pfull[slot], pphase, 74);
          tcgen05_fence_after();
          const uint32_t plo = p_base + (uint32_t)slot * (uint32_t)(S2_SLOT_BYTES >> 4);
#pragma unroll
          for (int oyx = 0; oyx < 9; ++oyx) {
            const int ty = oyx / 3, tx = oyx - ty * 3;
            const int hp = (ty != 1), wp = (tx != 1), dy = (ty == 2), dx = (tx == 2);
            const int sub_off = hp ? (wp ? S2_OFF11 : S2_OFF10) : (wp ? S2_OFF01 : S2_OFF00);
            const int nw = wp ? S2_NW1 : S2_NW0;
            const uint32_t alo = plo + (uint32_t)((sub_off + (dy * nw + dx) * S2_ROWB) >> 4);
            const uint64_t adesc = ((uint64_t)(wp ? a_hi1 : a_hi0) << 32) | (uint64_t)alo;
            const uint64_t bdesc = ((uint64_t)b_hi << 32) | (uint64_t)(wlo + (uint32_t)oyx * wstep);
#pragma unroll
            for (int k = 0; k < S2_KC / 16; ++k) {
              if (oyx == 0 && k == 0 && fresh_hi) {
                // output plane t_hi receives its first contribution here: start it from zero
                if (both) {
                  umma_bf16(d_col + (uint32_t)N, adesc, bdesc + (uint64_t)ntile16, idesc1, 0u);
                  umma_bf16(d_col, adesc, bdesc, idesc1, 1u);
                } else {
                  umma_bf16(d_col, adesc, bdesc, idesc1, 0u);
                }
              } else {
                umma_bf16(d_col, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
              }
            }
          }
          umma_commit(&bar_pempty[slot]);
          if (++slot == S2_RING) { slot = 0; pphase ^= 1; }
        }
        umma_commit(&bar_tfull[acc]);
        acc ^= 1;
        if (acc == 0) accphase ^= 1;
      }
    }
  } else if (warp >= 2 && warp <= 5) {
    halo_epilogue<MT>(P, tmem_base, bar_tfull, bar_tempty, s_stage[warp & 3], s_bias, warp & 3, lane);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, P.tmem_cols);
}

int pick_n_tile(int N) {
  if (N % 32) return 0;
  if (N <= 256) return N;
  for (int t = 256; t >= 32; t -= 32)
    if (N % t == 0) return t;
  return 0;
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace

long long* g_halo_prof = nullptr;
int g_plan_nt = 0, g_plan_mt = 0, g_plan_wst = 0;   // debug overrides of the tap-major kernel's plan (0 = automatic)

bool tc_halo_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MVD_NO_HALO");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

// narrow layers whose 27 weight tiles fit in shared memory next to a short plane ring run the depth-folded kernel
bool tc_halo_fold_eligible(int N, int K, int D) {
  static int fold_enabled = -1;
  if (fold_enabled < 0) {
    const char* e = getenv("MVD_NO_HALO_FOLD");
    fold_enabled = (e && e[0] == '1') ? 0 : 1;
  }
  if (!fold_enabled || !(N == 32 || N == 64) || D < 2 || K % 32) return false;
  const int kc = (K % 64 == 0) ? 64 : 32;
  const size_t plane_bytes_f = (size_t)((PLANE_ROWS * kc * 2 + 1023) & ~1023);
  return (size_t)27 * N * K * 2 + 3 * plane_bytes_f <= kFoldBudget;
}

int tc_halo_conv(const bf16* src, int lds, int K, bf16* dst, int ldd, int N, const bf16* w, const int wrow[27],
                 const float* bias, int accumulate, double* stats, int B, int D, int H, int W, cudaStream_t st,
                 const char* who, const mvd_norm_bwd_stats_args* nb) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) { set_error("%s: no cuTensorMapEncodeTiled", who); return MVD_ERR_CUDA; }
  const int kc = (K % 64 == 0) ? 64 : 32;
  HaloMaps maps;
  HaloParams P;
  memset(&P, 0, sizeof(P));
  P.n_tile = pick_n_tile(N);
  if (P.n_tile == 0 || K % 32 || N > kMaxN) { set_error("%s: unsupported channel counts", who); return MVD_ERR_UNSUPPORTED; }
  // Low-resolution layers have too few (brick, N-tile) work items for 148 SMs with the widest N tile: pick the tile
  // width that minimises  waves x planes-per-item x cycles-per-MMA  (MMA M128 x n x K16 = max(n/2, 32 + n/4) cycles),
  // keeping the widest tile unless a narrower one is predicted >= 20 % faster.  Fused statistics need n_tile == N.
  int mt_cap = 4;
  if (!stats && N > 64) {   // N <= 64 belongs to the depth-folded kernel below
    // per channel chunk: MMA cycles vs TMA box rows (3.7 cycles each) of one work item; whole launch = waves x that
    auto predicted = [&](int nt, int mt) -> double {
      const long long items = (long long)B * cdiv(D, mt) * cdiv(H, TILE_H) * cdiv(W, TILE_W) * (N / nt);
      const long long waves = (items + num_sms() - 1) / num_sms();
      const int mma = (nt / 2 > 32 + nt / 4) ? nt / 2 : 32 + nt / 4;
      const double t_mma = 27.0 * mt * (kc / 16) * mma;
      const double t_tma = 3.7 * (mt + 2) * PLANE_ROWS + 1.8 * 27.0 * nt;   // dense weight rows stream faster
      // + a fixed ~8k cycles per wave: pipeline fill, first-plane latency, epilogue drain of a short item
      return (double)waves * ((t_mma > t_tma ? t_mma : t_tma) * (K / kc) + 8000.0);
    };
    auto mt_max = [&](int nt) {
      int mt = 512 / (2 * nt);
      if (mt > 4) mt = 4;
      if (mt > D) mt = D;
      if (mt == 3) mt = 2;
      return mt < 1 ? 1 : mt;
    };
    double best = predicted(P.n_tile, mt_max(P.n_tile));
    mt_cap = mt_max(P.n_tile);
    const int nt0 = P.n_tile;
    for (int nt = nt0; nt >= 32; nt -= 32) {
      if (N % nt) continue;
      for (int mt = mt_max(nt); mt >= 1; mt >>= 1) {
        if (nt == nt0 && mt == mt_max(nt0)) continue;
        const double c = predicted(nt, mt);
        if (c < 0.8 * best) { best = c; P.n_tile = nt; mt_cap = mt; }
      }
    }
  }
  if (!stats && N > 64 && g_plan_nt > 0 && N % g_plan_nt == 0 && g_plan_nt % 32 == 0 && g_plan_nt <= 256) {
    P.n_tile = g_plan_nt;
    mt_cap = g_plan_mt > 0 ? g_plan_mt : 4;
  }
  if (stats && N != 32 && N != 64) { set_error("%s: fused InstanceNorm sums need N = 32 or 64", who); return MVD_ERR_UNSUPPORTED; }
  if (nb && (stats || !tc_halo_fold_eligible(N, K, D) || N != 32 || nb->ldy % 8 || !nb->y || !nb->stats || !nb->bstats)) {
    set_error("%s: the backward statistics of the norm in front are built for the depth-folded kernel with N = 32", who);
    return MVD_ERR_UNSUPPORTED;
  }
  {
    const long long ld = lds;
    cuuint64_t gdim[5] = {(cuuint64_t)K, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
    cuuint64_t gstr[4] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * W * 2, (cuuint64_t)ld * W * H * 2,
                          (cuuint64_t)ld * W * H * D * 2};
    cuuint32_t box[5] = {(cuuint32_t)kc, HALO_W, HALO_H, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&maps.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)src, gdim, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     tc_l2_promotion(kc * 2, ld * 2), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("%s: cuTensorMapEncodeTiled(planes) failed (%d)", who, (int)r); return MVD_ERR_CUDA; }
  }
  if (!tc_encode_w_map(&maps.b, w, (long long)27 * N, K, P.n_tile, kc)) {
    set_error("%s: cuTensorMapEncodeTiled(weights) failed", who);
    return MVD_ERR_CUDA;
  }
  // narrow layers whose weights fit in shared memory: depth-folded kernel (see conv_halo_fold_kernel)
  {
    const int rowb_f = kc * 2;
    const int plane_bytes_f = (PLANE_ROWS * rowb_f + 1023) & ~1023;
    const size_t w_res = (size_t)27 * N * K * 2;
    const size_t budget_f = kFoldBudget;
    if (tc_halo_fold_eligible(N, K, D)) {
      int MTf = 512 / (2 * N);                 // 8 (N = 32) or 4 (N = 64)
      while (MTf > D) MTf >>= 1;
      int ringf = (int)((budget_f - w_res) / plane_bytes_f);
      if (ringf > 6) ringf = 6;
      P.MT = MTf; P.ring = ringf; P.wstages = 0;
      P.B = B; P.D = D; P.H = H; P.W = W;
      P.tiles_w = cdiv(W, TILE_W); P.tiles_h = cdiv(H, TILE_H); P.dgroups = cdiv(D, MTf);
      P.num_n_tiles = 1;
      P.total_items = B * P.dgroups * P.tiles_h * P.tiles_w;
      P.kchunks = K / kc;
      P.idesc = 0;
      uint32_t colsf = 32;
      while ((int)colsf < 2 * MTf * N) colsf <<= 1;
      P.tmem_cols = colsf;
      P.out = dst;
      P.sw = ldd; P.sh = (long long)ldd * W; P.sd = P.sh * H; P.sb = P.sd * D;
      P.bias = bias; P.accumulate = accumulate;
      P.prof = g_halo_prof;
      P.stats = stats; P.Ntot = N;
      P.lane_own = (N == 32);      // narrow layers: the tensor pipe's operand reads own the shared-memory bandwidth
      if (nb) {
        P.nb_y = (const bf16*)nb->y; P.nb_ld = nb->ldy; P.nb_stats = nb->stats; P.nb_gamma = nb->gamma;
        P.nb_beta = nb->beta; P.nb_eps = nb->eps; P.nb_slope = nb->slope; P.nb_out = nb->bstats;
      }
      for (int i = 0; i < 27; ++i) P.wrow[i] = wrow[i];
      const size_t smemf = w_res + (size_t)ringf * plane_bytes_f + 1024;
      int gridf = num_sms();
      if (gridf > P.total_items) gridf = P.total_items;
      void (*kf)(const HaloMaps, const HaloParams) = nullptr;
      int kif = 0;
#define FOLD_PICK(KCV, MTV, IDX)                       \
  if (kc == KCV && MTf == MTV) {                       \
    kf = conv_halo_fold_kernel<KCV, MTV>;              \
    kif = IDX;                                         \
  }
      FOLD_PICK(64, 2, 0) FOLD_PICK(64, 4, 1) FOLD_PICK(64, 8, 2) FOLD_PICK(32, 2, 3) FOLD_PICK(32, 4, 4) FOLD_PICK(32, 8, 5)
#undef FOLD_PICK
#define FOLD_PICK_NB(KCV, MTV, IDX)                    \
  if (nb && kc == KCV && MTf == MTV) {                 \
    kf = conv_halo_fold_kernel<KCV, MTV, true>;        \
    kif = IDX;                                         \
  }
      // N = 32 -> MT = 8 (4 / 2 for short volumes)
      FOLD_PICK_NB(64, 2, 6) FOLD_PICK_NB(64, 4, 7) FOLD_PICK_NB(64, 8, 8) FOLD_PICK_NB(32, 2, 9) FOLD_PICK_NB(32, 4, 10)
      FOLD_PICK_NB(32, 8, 11)
#undef FOLD_PICK_NB
      if (kf) {
        static bool fattr[12] = {};
        if (!fattr[kif]) {
          cudaError_t e = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 217 * 1024);
          if (e != cudaSuccess) {
            (void)cudaGetLastError();
            set_error("%s: cudaFuncSetAttribute(fold): %s", who, cudaGetErrorString(e));
            return MVD_ERR_CUDA;
          }
          fattr[kif] = true;
        }
        kf<<<gridf, kThreads, smemf, st>>>(maps, P);
        MVD_LAUNCH_CHECK(who);
        return MVD_OK;
      }
    }
  }
  // MT: as many output planes per item as TMEM (2 x MT x n_tile <= 512 columns) and shared memory allow, at most 4
  int MT = 512 / (2 * P.n_tile);
  if (MT > 4) MT = 4;
  if (MT > mt_cap) MT = mt_cap;
  if (MT > D) MT = D;
  if (MT == 3) MT = 2;
  if (MT < 1) MT = 1;
  const int rowb = kc * 2;
  const int plane_bytes = (PLANE_ROWS * rowb + 1023) & ~1023;
  const int w_bytes = P.n_tile * rowb;
  const int budget = 198 * 1024;
  int ring, wstages;
  for (;;) {
    ring = MT + 3;
    if (ring > kMaxRing) ring = kMaxRing;
    wstages = (budget - ring * plane_bytes) / w_bytes;
    if (wstages >= 3 || MT == 1) break;
    MT >>= 1;
  }
  if (wstages > kMaxWStages) wstages = kMaxWStages;
  if (g_plan_wst > 0 && wstages > g_plan_wst) wstages = g_plan_wst;
  if (wstages < 2) {
    ring = MT + 2;
    wstages = (budget - ring * plane_bytes) / w_bytes;
    if (wstages > kMaxWStages) wstages = kMaxWStages;
    if (wstages < 2) { set_error("%s: tile does not fit shared memory", who); return MVD_ERR_UNSUPPORTED; }
  }
  P.MT = MT; P.ring = ring; P.wstages = wstages;
  P.B = B; P.D = D; P.H = H; P.W = W;
  P.tiles_w = cdiv(W, TILE_W); P.tiles_h = cdiv(H, TILE_H); P.dgroups = cdiv(D, MT);
  P.num_n_tiles = N / P.n_tile;
  P.total_items = B * P.dgroups * P.tiles_h * P.tiles_w * P.num_n_tiles;
  P.kchunks = K / kc;
  P.idesc = make_idesc_bf16(128, P.n_tile, 0, 0);
  uint32_t cols = 32;
  while ((int)cols < 2 * MT * P.n_tile) cols <<= 1;
  P.tmem_cols = cols;
  P.out = dst;
  P.sw = ldd; P.sh = (long long)ldd * W; P.sd = P.sh * H; P.sb = P.sd * D;
  P.bias = bias; P.accumulate = accumulate;
  P.prof = g_halo_prof;
  P.stats = stats; P.Ntot = N;
  for (int i = 0; i < 27; ++i) P.wrow[i] = wrow[i];
  const size_t smem = (size_t)ring * plane_bytes + (size_t)wstages * w_bytes + 1024;
  int grid = num_sms();
  if (grid > P.total_items) grid = P.total_items;
  void (*kern)(const HaloMaps, const HaloParams) = nullptr;
  int ki = 0;
#define HALO_PICK(KCV, MTV, IDX)                         \
  if (kc == KCV && MT == MTV) {                          \
    kern = conv_halo_kernel<KCV, MTV>;                   \
    ki = IDX;                                            \
  }
  HALO_PICK(64, 1, 0) HALO_PICK(64, 2, 1) HALO_PICK(64, 4, 2) HALO_PICK(32, 1, 3) HALO_PICK(32, 2, 4) HALO_PICK(32, 4, 5)
#undef HALO_PICK
  if (!kern) { set_error("%s: no halo kernel for kc=%d MT=%d", who, kc, MT); return MVD_ERR_UNSUPPORTED; }
  static bool attr_done[6] = {false, false, false, false, false, false};
  if (!attr_done[ki]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(e));
      return MVD_ERR_CUDA;
    }
    attr_done[ki] = true;
  }
  kern<<<grid, kThreads, smem, st>>>(maps, P);
  MVD_LAUNCH_CHECK(who);
  return MVD_OK;
}

// ---- stride-2 forward (conv_halo_s2_kernel) ----------------------------------------------------------------------
bool tc_halo_s2_fprop_supported(const mvd_conv3d_args* a) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("MVD_NO_S2HALO");
    enabled = (e && e[0] == '1') ? 0 : 1;
  }
  if (!enabled || !tc_halo_enabled()) return false;
  if (!(a->kd == 3 && a->kh == 3 && a->kw == 3 && a->sd == 2 && a->sh == 2 && a->sw == 2 && a->pd == 1 && a->ph == 1 &&
        a->pw == 1))
    return false;
  if (a->Cin != S2_KC || (a->Cout != 32 && a->Cout != 64)) return false;
  if (a->Di < 2 || a->Hi < 2 || a->Wi < 2) return false;   // every parity lattice must be non-empty
  if (a->ldx % 8 || a->ldy % 8 || ((uintptr_t)a->x & 15) || ((uintptr_t)a->y & 15) || ((uintptr_t)a->w & 15)) return false;
  return get_encode_tiled() != nullptr;
}

int tc_halo_s2_fprop(const mvd_conv3d_args* a, cudaStream_t st) {
  const char* who = "conv3d_fprop(tcgen05 stride-2 halo)";
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) { set_error("%s: no cuTensorMapEncodeTiled", who); return MVD_ERR_CUDA; }
  S2Maps maps;
  HaloParams P;
  memset(&P, 0, sizeof(P));
  const int N = a->Cout;
  const long long ld = a->ldx;
  for (int hp = 0; hp < 2; ++hp)
    for (int wp = 0; wp < 2; ++wp) {
      // even lattice: h = 2i; odd lattice: h = 2i + 1 (the kernel asks for line oy0 - 1 = input row 2 oy0 - 1)
      const bf16* base = (const bf16*)a->x + ((long long)hp * a->Wi + wp) * ld;
      cuuint64_t gdim[5] = {(cuuint64_t)S2_KC, (cuuint64_t)((a->Wi + 1 - wp) / 2), (cuuint64_t)((a->Hi + 1 - hp) / 2),
                            (cuuint64_t)a->Di, (cuuint64_t)a->B};
      cuuint64_t gstr[4] = {(cuuint64_t)ld * 4, (cuuint64_t)ld * a->Wi * 4, (cuuint64_t)ld * a->Wi * a->Hi * 2,
                            (cuuint64_t)ld * a->Wi * a->Hi * a->Di * 2};
      cuuint32_t box[5] = {(cuuint32_t)S2_KC, (cuuint32_t)(wp ? S2_NW1 : S2_NW0), (cuuint32_t)(hp ? S2_NH1 : S2_NH0), 1, 1};
      cuuint32_t es[5] = {1, 1, 1, 1, 1};
      CUresult r = enc(&maps.a[hp * 2 + wp], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)base, gdim, gstr, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, tc_l2_promotion(S2_ROWB, ld * 4),
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("%s: cuTensorMapEncodeTiled(planes) failed (%d)", who, (int)r); return MVD_ERR_CUDA; }
    }
  if (!tc_encode_w_map(&maps.b, (const bf16*)a->w, (long long)27 * N, S2_KC, N, S2_KC)) {
    set_error("%s: cuTensorMapEncodeTiled(weights) failed", who);
    return MVD_ERR_CUDA;
  }
  P.MT = S2_MT; P.ring = S2_RING; P.wstages = 0;
  P.n_tile = N; P.num_n_tiles = 1; P.kchunks = 1;
  P.B = a->B; P.D = a->Do; P.H = a->Ho; P.W = a->Wo;
  P.tiles_w = cdiv(a->Wo, TILE_W); P.tiles_h = cdiv(a->Ho, TILE_H); P.dgroups = cdiv(a->Do, S2_MT);
  P.total_items = a->B * P.dgroups * P.tiles_h * P.tiles_w;
  uint32_t cols = 32;
  while ((int)cols < 2 * S2_MT * N) cols <<= 1;
  P.tmem_cols = cols;
  P.out = (bf16*)a->y;
  P.sw = a->ldy; P.sh = (long long)a->ldy * a->Wo; P.sd = P.sh * a->Ho; P.sb = P.sd * a->Do;
  P.bias = a->bias; P.accumulate = 0;
  P.stats = a->stats; P.Ntot = N;
  for (int i = 0; i < 27; ++i) P.wrow[i] = i * N;
  const size_t smem = (size_t)27 * N * S2_ROWB + (size_t)S2_RING * S2_SLOT_BYTES + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_s2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         27 * 64 * S2_ROWB + S2_RING * S2_SLOT_BYTES + 1024);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(e));
      return MVD_ERR_CUDA;
    }
    attr_done = true;
  }
  int grid = num_sms();
  if (grid > P.total_items) grid = P.total_items;
  conv_halo_s2_kernel<<<grid, kThreads, smem, st>>>(maps, P);
  MVD_LAUNCH_CHECK(who);
  return MVD_OK;
}

}  // namespace mvd

// debug hook: per-CTA cycle counters of the halo kernel's MMA issuer ([grid][8] int64, device memory) or NULL
extern "C" void mvd_debug_set_halo_prof(long long* buf) { mvd::g_halo_prof = buf; }
// debug hook: force the tap-major halo kernel's N-tile width / planes per item / weight-ring depth (0 = automatic)
extern "C" void mvd_debug_set_halo_plan(int n_tile, int mt, int wstages) {
  mvd::g_plan_nt = n_tile; mvd::g_plan_mt = mt; mvd::g_plan_wst = wstages;
}
