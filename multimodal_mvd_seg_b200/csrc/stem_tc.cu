// stem_tc.cu -- the network's first convolution (Cin = 1 or 2 modalities -> 32 features, 3x3x3, stride 1, pad 1) as a
// tcgen05 GEMM whose im2col A tile is built IN SHARED MEMORY: no X_col matrix ever reaches HBM.
//
// K = 27*Cin is too thin for an implicit GEMM over taps, so the stem is an explicit [voxels x KPAD] x [KPAD x 32] GEMM
// with KPAD = 32*Cin columns (k = tap*Cin + ci, zero padded).  Round 1 materialised X_col ([V][KPAD] bf16 = 537 MB at
// 2 x 128^3, written once and read by fprop and wgrad: 0.68 ms per step).  Here four producer warps build every
// 128-voxel brick's tile straight into the UMMA shared-memory layout:
//   * the brick's input halo (3 planes x 18 h x 10 w voxels, 2*Cin bytes each) is staged with plain loads;
//   * thread r assembles row r (one voxel): KPAD/8 16-byte chunks gathered from the halo (tap decode is compile-time),
//     stored at  tile + r*ROWB + ((chunk ^ swz(r)) << 4)  -- the 128B / 64B swizzle the tensor core applies as a pure
//     function of the shared-memory address (profiles/r1_umma_descriptor_probe.txt), i.e. what TMA would have written;
//   * fence.proxy.async makes the generic-proxy stores visible to the tensor core, then the row's thread arrives on the
//     tile's mbarrier.
// fprop: A = tile (K-major), B = the 32 x KPAD weight matrix (resident), 2-4 MMAs per brick, epilogue with bias,
//        InstanceNorm partial sums and coalesced bf16 stores (tc_epilogue.cuh).
// wgrad: A = the same tile read MN-major (rows = the reduction index), B = the dy brick via TMA, accumulator
//        [KPAD (of 128 rows)][32] kept in TMEM for the CTA's whole voxel range, fp32 atomics into dw_col at the end.
#include "conv_common.cuh"
#include "tc_common.cuh"
#include "tc_epilogue.cuh"

namespace mvd {
namespace {

using namespace tc;

// warps 0..7: two groups of four tile producers (alternating bricks); warp 8: MMA issuer (+ TMA);
// fprop: warps 9..16 = two sets of four epilogue warps (set = accumulator buffer = brick parity; 17 warps leave 96
// registers per thread, so the epilogue keeps only two running sums per lane); wgrad: warps 9..12 run the final epilogue
constexpr int kIssuerWarp = 8;
constexpr int kThreadsFprop = 17 * 32, kThreadsWgrad = 13 * 32;
constexpr int kProducers = 128;   // threads per producer group = rows of a tile
constexpr int TILE_W = 8, TILE_H = 16, HALO_W = 10, HALO_H = 18, HALO_VOX = 3 * HALO_H * HALO_W;   // 540
// Shared-memory pitch of a halo line, in voxels.  A warp's 32 rows read 4 h-lines x 8 consecutive voxels per tap; with
// 32-bit voxels (CIN = 2) and the natural pitch of 10 the fourth line lands on the banks of the first (ncu: 6.4 M of 12 M
// shared-load wavefronts were bank conflicts); a pitch of 24 words puts the four lines on banks 0 / 24 / 16 / 8.
template <int CIN> struct HaloPitch {
  static constexpr int value = (CIN == 2) ? 24 : HALO_W;
  static constexpr int vox = 3 * HALO_H * value;      // voxels of one staged halo in shared memory
};
constexpr int NOUT = 32;          // output features of the stem
constexpr int kStages = 4;
constexpr bool kLaneOwnStores = true;

struct StemParams {
  const bf16* x;                  // dense NDHWC input, CIN channels
  int B, D, H, W;
  int tiles_w, tiles_h, total_tiles;
  // fprop
  bf16* y; long long ysb, ysd, ysh, ysw;
  const float* bias;
  double* stats;
  // wgrad
  float* dw_col;                  // [32][CIN][27] fp32 (torch layout), accumulated with atomics (zeroed by the entry point)
};

struct alignas(64) StemMaps {
  CUtensorMap w;    // fprop: weights [32 rows][KPAD] box (KPAD, 32)
  CUtensorMap dy;   // wgrad: dy (C = 32, W, H, D, B) box (32, 8, 16, 1, 1), SWIZZLE_64B
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void producer_bar(int group) {
  asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory");
}

__device__ __forceinline__ void decode_tile(const StemParams& P, int tile, int& b, int& d, int& h0, int& w0) {
  w0 = (tile % P.tiles_w) * TILE_W; tile /= P.tiles_w;
  h0 = (tile % P.tiles_h) * TILE_H; tile /= P.tiles_h;
  d = tile % P.D;
  b = tile / P.D;
}

// producer warps.  Software-pipelined: the halo of brick i+1 is fetched into REGISTERS (5 voxels per thread) while brick
// i's tile is assembled from shared memory, so the global-load latency is off the per-tile critical path.
constexpr int kHaloPerThread = (HALO_VOX + kProducers - 1) / kProducers;   // 5

template <int CIN>
struct HaloRegs { uint32_t v[kHaloPerThread]; };

template <int CIN>
__device__ __forceinline__ void halo_fetch(const StemParams& P, int tile, int t, HaloRegs<CIN>& r) {
  int b, d, h0, w0;
  decode_tile(P, tile, b, d, h0, w0);
#pragma unroll
  for (int u = 0; u < kHaloPerThread; ++u) {
    const int i = t + u * kProducers;
    uint32_t val = 0u;
    if (i < HALO_VOX) {
      const int pd = i / (HALO_H * HALO_W), rem = i - pd * (HALO_H * HALO_W);
      const int ph = rem / HALO_W, pw = rem - ph * HALO_W;
      const int z = d + pd - 1, yy = h0 + ph - 1, xx = w0 + pw - 1;
      if (z >= 0 && z < P.D && yy >= 0 && yy < P.H && xx >= 0 && xx < P.W) {
        const long long g = ((((long long)b * P.D + z) * P.H + yy) * P.W + xx) * CIN;
        if (CIN == 2) val = __ldg(reinterpret_cast<const uint32_t*>(P.x + g));
        else val = __ldg(reinterpret_cast<const unsigned short*>(P.x + g));
      }
    }
    r.v[u] = val;
  }
}

template <int CIN>
__device__ __forceinline__ void halo_store(bf16* halo, int t, const HaloRegs<CIN>& r) {
#pragma unroll
  for (int u = 0; u < kHaloPerThread; ++u) {
    const int i = t + u * kProducers;
    if (i < HALO_VOX) {
      const int line = i / HALO_W, pw = i - line * HALO_W;          // line = pd * HALO_H + ph
      const int o = line * HaloPitch<CIN>::value + pw;
      if (CIN == 2) reinterpret_cast<uint32_t*>(halo)[o] = r.v[u];
      else reinterpret_cast<unsigned short*>(halo)[o] = (unsigned short)r.v[u];
    }
  }
}

// thread t assembles row t (one voxel) of the [128][KPAD] tile from the staged halo
template <int CIN>
__device__ __forceinline__ void build_row(const bf16* halo, uint8_t* tile, int t /* 0..127 */) {
  constexpr int KPAD = 32 * CIN, ROWB = KPAD * 2, CHUNKS = ROWB / 16;
  const int hh = t >> 3, ww = t & 7;
  const int swz = (CIN == 2) ? (t & 7) : ((t >> 1) & 3);      // SWIZZLE_128B / SWIZZLE_64B chunk XOR of row t
  uint8_t* row = tile + t * ROWB;
#pragma unroll
  for (int j = 0; j < CHUNKS; ++j) {
    uint32_t wv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {       // 32-bit word e of chunk j = k values 8j + 2e, 8j + 2e + 1
      if (CIN == 2) {
        const int tap = 4 * j + e;      // one tap = both channels = one 32-bit halo word
        uint32_t v = 0u;
        if (tap < 27) {
          const int td = tap / 9, th = (tap / 3) % 3, tw = tap % 3;
          v = reinterpret_cast<const uint32_t*>(halo)[(td * HALO_H + hh + th) * HaloPitch<CIN>::value + ww + tw];
        }
        wv[e] = v;
      } else {
        uint32_t v = 0u;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int tap = 8 * j + 2 * e + s;
          if (tap < 27) {
            const int td = tap / 9, th = (tap / 3) % 3, tw = tap % 3;
            const uint32_t u = reinterpret_cast<const unsigned short*>(halo)[(td * HALO_H + hh + th) * HaloPitch<CIN>::value + ww + tw];
            v |= u << (16 * s);
          }
        }
        wv[e] = v;
      }
    }
    *reinterpret_cast<uint4*>(row + ((j ^ swz) << 4)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
  }
  fence_proxy_async();
}

// the producer loop shared by fprop and wgrad: group g (warps 4g..4g+3) builds the bricks with local index = g mod 2
// into ring stage (local index % kStages); one mbarrier arrival per warp
template <int CIN>
__device__ __forceinline__ void produce_tiles(const StemParams& P, bf16 (*s_halo)[HaloPitch<CIN>::vox * CIN], uint8_t* smem_a,
                                              int tile_bytes, uint64_t* bar_full, uint64_t* bar_empty) {
  const int group = threadIdx.x >> 7, t = threadIdx.x & 127, lane = threadIdx.x & 31;
  bf16 (*halo2)[HaloPitch<CIN>::vox * CIN] = s_halo + 2 * group;
  HaloRegs<CIN> regs;
  const int first = blockIdx.x + group * gridDim.x, step = 2 * gridDim.x;
  if (first < P.total_tiles) halo_fetch<CIN>(P, first, t, regs);
  int it = group, k = 0;                      // it: local brick index of this CTA, k: this group's iteration
  for (int tile = first; tile < P.total_tiles; tile += step, it += 2, ++k) {
    bf16* halo = halo2[k & 1];
    halo_store<CIN>(halo, t, regs);
    if (tile + step < P.total_tiles) halo_fetch<CIN>(P, tile + step, t, regs);   // in flight during barrier + build
    producer_bar(group);
    const int stage = it % kStages;
    const uint32_t phase = (uint32_t)(it / kStages) & 1u;
    mbar_wait(&bar_empty[stage], phase ^ 1, 71);
    build_row<CIN>(halo, smem_a + (size_t)stage * tile_bytes, t);
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar_full[stage]);
  }
}

// ------------------------------------------------------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(kThreadsFprop, 1) stem_fprop_kernel(const __grid_constant__ StemMaps maps,
                                                                 const __grid_constant__ StemParams P) {
  constexpr int KPAD = 32 * CIN, ROWB = KPAD * 2, TILE_BYTES = 128 * ROWB;
  constexpr uint64_t LAYOUT = (CIN == 2) ? kLayoutSw128 : kLayoutSw64;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar_full[kStages], bar_empty[kStages], bar_w, bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(16) uint8_t s_stage[8][2048];
  __shared__ __align__(16) float s_bias[NOUT];
  __shared__ __align__(16) bf16 s_halo[4][HaloPitch<CIN>::vox * CIN];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem;                         // [32][KPAD]
  uint8_t* smem_a = smem + 4096;                  // kStages tiles
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&bar_full[s], 4); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_w, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&bar_tfull[a], 1); mbar_init(&bar_tempty[a], 4); }
    fence_barrier_init();
  }
  if (threadIdx.x < NOUT) s_bias[threadIdx.x] = P.bias ? round_bf(__ldg(P.bias + threadIdx.x)) : 0.f;
  if (warp == kIssuerWarp) tmem_alloc(&s_tmem_base, 64);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp < kIssuerWarp) {
    // ================= tile producers =================
    produce_tiles<CIN>(P, s_halo, smem_a, TILE_BYTES, bar_full, bar_empty);
  } else if (warp == kIssuerWarp) {
    if (elect_one_sync()) {
      // ================= weights (once) + MMA issuer =================
      mbar_arrive_expect_tx(&bar_w, (uint32_t)(NOUT * ROWB));
      tma_load_2d(&maps.w, smem_w, &bar_w, 0, 0);
      const uint32_t idesc = make_idesc_bf16(128, NOUT, 0, 0);
      const uint32_t hi = (uint32_t)(make_smem_desc(0, 16, 8 * ROWB, LAYOUT) >> 32);
      const uint32_t a_base = (smem_u32(smem_a) >> 4) | (1u << 16);
      const uint32_t w_lo = (smem_u32(smem_w) >> 4) | (1u << 16);
      mbar_wait(&bar_w, 0, 72);
      int stage = 0, acc = 0;
      uint32_t phase = 0, accphase = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        mbar_wait(&bar_tempty[acc], accphase ^ 1, 73);
        mbar_wait(&bar_full[stage], phase, 74);
        tcgen05_fence_after();
        const uint32_t a_lo = a_base + (uint32_t)stage * (uint32_t)(TILE_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < KPAD / 16; ++k)
          umma_bf16(tmem_base + (uint32_t)(acc * NOUT), ((uint64_t)hi << 32) | (a_lo + 2u * k),
                    ((uint64_t)hi << 32) | (w_lo + 2u * k), idesc, k ? 1u : 0u);
        umma_commit(&bar_empty[stage]);
        umma_commit(&bar_tfull[acc]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
        acc ^= 1;
        if (acc == 0) accphase ^= 1;
      }
    }
  } else {
    // ================= epilogue: set s = (warp - 9) / 4 takes the bricks whose local index has parity s ==========
    const int q = warp & 3, set = (warp - kIssuerWarp - 1) >> 2;
    uint8_t* stage_buf = s_stage[warp - kIssuerWarp - 1];
    const int acc = set;
    uint32_t accphase = 0;
    // InstanceNorm sums: 17 warps leave 96 registers per thread (5 warps on one scheduler partition), too few for the
    // per-lane partial arrays of LaneStats -- so each tile is reduced across the warp's rows right away (lane l ends
    // up with channel l) and only two running sums per lane are carried
    float ssum = 0.f, ssq = 0.f;
    int sb = -1;
    const bool do_stats = P.stats != nullptr;
    auto flush = [&]() {
      double* dst = P.stats + ((long long)sb * NOUT + lane) * 2;
      atomicAdd(dst, (double)ssum);
      atomicAdd(dst + 1, (double)ssq);
      ssum = ssq = 0.f;
    };
    for (int tile = blockIdx.x + set * gridDim.x; tile < P.total_tiles; tile += 2 * gridDim.x) {
      int b, d, h0, w0;
      decode_tile(P, tile, b, d, h0, w0);
      if (do_stats && b != sb) {
        if (sb >= 0) flush();
        sb = b;
      }
      mbar_wait(&bar_tfull[acc], accphase, 75);
      accphase ^= 1;
      tcgen05_fence_after();
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NOUT), v);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[acc]);    // the accumulator is in registers: release it early
      const int rr0 = q * 32 + lane;
      const bool ok = (h0 + (rr0 >> 3) < P.H) && (w0 + (rr0 & 7) < P.W);
      uint32_t w2[16];
      {
        LaneStats<0> none;
        epilogue_chunk<0>(v, s_bias, none, 0, ok, w2);
      }
      if (do_stats) {   // two passes over one 32-entry array (register budget)
        float a[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          a[2 * j] = ok ? __uint_as_float(w2[j] << 16) : 0.f;
          a[2 * j + 1] = ok ? __uint_as_float(w2[j] & 0xffff0000u) : 0.f;
        }
        ssum += warp_column_sum32(a, lane);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float lo = ok ? __uint_as_float(w2[j] << 16) : 0.f, hi = ok ? __uint_as_float(w2[j] & 0xffff0000u) : 0.f;
          a[2 * j] = lo * lo;
          a[2 * j + 1] = hi * hi;
        }
        ssq += warp_column_sum32(a, lane);
      }
      bf16* base = P.y + (long long)b * P.ysb + (long long)d * P.ysd;
      if (kLaneOwnStores) {   // no shared-memory transpose next to the tile producers' traffic (tc_epilogue.cuh)
        if (ok) store_row_lane_own(base + (long long)(h0 + (rr0 >> 3)) * P.ysh + (long long)(w0 + (rr0 & 7)) * P.ysw, w2, false);
      } else {
        store_rows_coalesced_packed(stage_buf, lane, w2, [&](int R) -> bf16* {
          const int rr = q * 32 + R;
          const int h = h0 + (rr >> 3), w = w0 + (rr & 7);
          return (h < P.H && w < P.W) ? base + (long long)h * P.ysh + (long long)w * P.ysw : nullptr;
        }, false);
      }
    }
    if (do_stats && sb >= 0) flush();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kIssuerWarp) tmem_dealloc(tmem_base, 64);
}

// ------------------------------------------------------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(kThreadsWgrad, 1) stem_wgrad_kernel(const __grid_constant__ StemMaps maps,
                                                                 const __grid_constant__ StemParams P) {
  constexpr int KPAD = 32 * CIN, ROWB = KPAD * 2, TILE_BYTES = 128 * ROWB;
  constexpr uint64_t A_LAYOUT = (CIN == 2) ? kLayoutSw128 : kLayoutSw64;
  constexpr int BRICK_BYTES = 128 * 64;           // dy brick [128 v][32 co] bf16
  constexpr int kBRing = 3;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar_full[kStages], bar_empty[kStages], bar_bfull[kBRing], bar_bempty[kBRing], bar_done;
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(16) bf16 s_halo[4][HaloPitch<CIN>::vox * CIN];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem;                                   // kBRing dy bricks
  uint8_t* smem_a = smem + kBRing * BRICK_BYTES;            // kStages tiles (+ slack read by the junk M rows)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&bar_full[s], 4); mbar_init(&bar_empty[s], 1); }
    for (int s = 0; s < kBRing; ++s) { mbar_init(&bar_bfull[s], 1); mbar_init(&bar_bempty[s], 1); }
    mbar_init(&bar_done, 1);
    fence_barrier_init();
  }
  if (warp == kIssuerWarp) tmem_alloc(&s_tmem_base, 32);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  const bool has_work = (int)blockIdx.x < P.total_tiles;

  if (warp < kIssuerWarp) {
    produce_tiles<CIN>(P, s_halo, smem_a, TILE_BYTES, bar_full, bar_empty);
  } else if (warp == kIssuerWarp) {
    if (has_work && elect_one_sync()) {
      // ================= dy bricks by TMA + MMA issuer (one thread: loads run one brick ahead) =================
      // A: MN-major (rows = voxels = the reduction index, KPAD contiguous), M = 128 = KPAD valid rows + junk blocks
      // (LBO = one tile: the junk blocks read the following tiles / slack, their accumulator rows are never used)
      const uint32_t idesc = make_idesc_bf16(128, NOUT, 1, 1);
      const uint32_t a_hi = (uint32_t)(make_smem_desc(0, TILE_BYTES, 8 * ROWB, A_LAYOUT) >> 32);
      const uint32_t b_hi = (uint32_t)(make_smem_desc(0, BRICK_BYTES, 8 * 64, kLayoutSw64) >> 32);
      const uint32_t a_base = (smem_u32(smem_a) >> 4) | ((uint32_t)(TILE_BYTES >> 4) << 16);
      const uint32_t b_base = (smem_u32(smem_b) >> 4) | ((uint32_t)(BRICK_BYTES >> 4) << 16);
      auto load_brick = [&](int tile, int slot, uint32_t ph) {
        int b, d, h0, w0;
        decode_tile(P, tile, b, d, h0, w0);
        mbar_wait(&bar_bempty[slot], ph ^ 1, 82);
        mbar_arrive_expect_tx(&bar_bfull[slot], (uint32_t)BRICK_BYTES);
        tma_load_5d(&maps.dy, smem_b + slot * BRICK_BYTES, &bar_bfull[slot], 0, w0, h0, d, b);
      };
      int stage = 0, bs = 0, ls = 0;
      uint32_t phase = 0, bph = 0, lph = 0;
      int next = blockIdx.x;
      // prefetch up to kBRing - 1 bricks
      for (int i = 0; i < kBRing - 1 && next < P.total_tiles; ++i, next += gridDim.x) {
        load_brick(next, ls, lph);
        if (++ls == kBRing) { ls = 0; lph ^= 1; }
      }
      bool first = true;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        if (next < P.total_tiles) {
          load_brick(next, ls, lph);
          if (++ls == kBRing) { ls = 0; lph ^= 1; }
          next += gridDim.x;
        }
        mbar_wait(&bar_full[stage], phase, 83);
        mbar_wait(&bar_bfull[bs], bph, 84);
        tcgen05_fence_after();
        const uint32_t a_lo = a_base + (uint32_t)stage * (uint32_t)(TILE_BYTES >> 4);
        const uint32_t b_lo = b_base + (uint32_t)bs * (uint32_t)(BRICK_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < 8; ++k)   // 16 voxels (two 8-row groups) per MMA
          umma_bf16(tmem_base, ((uint64_t)a_hi << 32) | (a_lo + (uint32_t)(k * ((2 * 8 * ROWB) >> 4))),
                    ((uint64_t)b_hi << 32) | (b_lo + (uint32_t)(k * ((2 * 8 * 64) >> 4))), idesc,
                    (first && k == 0) ? 0u : 1u);
        first = false;
        umma_commit(&bar_empty[stage]);
        umma_commit(&bar_bempty[bs]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
        if (++bs == kBRing) { bs = 0; bph ^= 1; }
      }
      umma_commit(&bar_done);
    }
  } else if (has_work) {
    // ================= epilogue: accumulator rows 0..KPAD-1 -> dw_col[co][k] =================
    const int q = warp & 3;
    mbar_wait(&bar_done, 0, 85);
    tcgen05_fence_after();
    const int m = q * 32 + lane;
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16), v);
    tmem_ld_wait();
    if (m < 27 * CIN) {      // row m = tap * CIN + ci  ->  torch layout dw[co][ci][tap]
      const int tap = m / CIN, ci = m - tap * CIN;
#pragma unroll
      for (int e = 0; e < 32; ++e) atomicAdd(&P.dw_col[(e * CIN + ci) * 27 + tap], __uint_as_float(v[e]));
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == kIssuerWarp) tmem_dealloc(tmem_base, 32);
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

bool fill_common(StemParams& P, const void* x, int B, int D, int H, int W) {
  memset(&P, 0, sizeof(P));
  P.x = (const bf16*)x;
  P.B = B; P.D = D; P.H = H; P.W = W;
  P.tiles_w = cdiv(W, TILE_W); P.tiles_h = cdiv(H, TILE_H);
  const long long tiles = (long long)B * D * P.tiles_h * P.tiles_w;
  if (tiles >= (1LL << 31)) return false;
  P.total_tiles = (int)tiles;
  return true;
}

template <typename K>
int set_smem_attr(K kern, size_t smem, bool* done, const char* who) {
  if (*done) return MVD_OK;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(e));
    return MVD_ERR_CUDA;
  }
  *done = true;
  return MVD_OK;
}

}  // namespace
}  // namespace mvd

using namespace mvd;

extern "C" {

int mvd_stem_conv_fprop(const void* x, int B, int D, int H, int W, int Cin, const void* wcol, const float* bias,
                        void* y, int ldy, double* stats, mvd_stream_t stream) {
  const char* who = "stem_conv_fprop";
  MVD_REQUIRE(x && wcol && y && B > 0 && D > 0 && H > 0 && W > 0, "%s: bad arguments", who);
  MVD_REQUIRE(Cin == 1 || Cin == 2, "%s: built for 1 or 2 input modalities (got %d)", who, Cin);
  MVD_REQUIRE(ldy >= NOUT && ldy % 8 == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)wcol & 15) == 0 &&
                  ((uintptr_t)x & 3) == 0, "%s: alignment / pitch", who);
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) { set_error("%s: no cuTensorMapEncodeTiled", who); return MVD_ERR_CUDA; }
  const int kpad = 32 * Cin;
  StemMaps maps;
  StemParams P;
  MVD_REQUIRE(fill_common(P, x, B, D, H, W), "%s: volume too large", who);
  if (!tc_encode_w_map(&maps.w, (const bf16*)wcol, NOUT, kpad, NOUT, kpad)) {
    set_error("%s: cuTensorMapEncodeTiled(weights) failed", who);
    return MVD_ERR_CUDA;
  }
  maps.dy = maps.w;
  P.y = (bf16*)y;
  P.ysw = ldy; P.ysh = (long long)ldy * W; P.ysd = P.ysh * H; P.ysb = P.ysd * D;
  P.bias = bias; P.stats = stats;
  const size_t smem = 4096 + (size_t)kStages * 128 * kpad * 2 + 1024;
  int grid = num_sms();
  if (grid > P.total_tiles) grid = P.total_tiles;
  static bool a1 = false, a2 = false;
  int rc;
  if (Cin == 1) {
    if ((rc = set_smem_attr(stem_fprop_kernel<1>, smem, &a1, who))) return rc;
    stem_fprop_kernel<1><<<grid, kThreadsFprop, smem, (cudaStream_t)stream>>>(maps, P);
  } else {
    if ((rc = set_smem_attr(stem_fprop_kernel<2>, smem, &a2, who))) return rc;
    stem_fprop_kernel<2><<<grid, kThreadsFprop, smem, (cudaStream_t)stream>>>(maps, P);
  }
  MVD_LAUNCH_CHECK(who);
  return MVD_OK;
}

int mvd_stem_conv_wgrad(const void* x, int B, int D, int H, int W, int Cin, const void* dy, int lddy, float* dw_col,
                        mvd_stream_t stream) {
  const char* who = "stem_conv_wgrad";
  MVD_REQUIRE(x && dy && dw_col && B > 0 && D > 0 && H > 0 && W > 0, "%s: bad arguments", who);
  MVD_REQUIRE(Cin == 1 || Cin == 2, "%s: built for 1 or 2 input modalities (got %d)", who, Cin);
  MVD_REQUIRE(lddy >= NOUT && lddy % 8 == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)x & 3) == 0,
              "%s: alignment / pitch", who);
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) { set_error("%s: no cuTensorMapEncodeTiled", who); return MVD_ERR_CUDA; }
  const int kpad = 32 * Cin;
  StemMaps maps;
  StemParams P;
  MVD_REQUIRE(fill_common(P, x, B, D, H, W), "%s: volume too large", who);
  {
    const int dims[4] = {W, H, D, B};
    const long long ld = lddy;
    const long long strides[4] = {ld, ld * W, ld * W * H, ld * W * H * D};
    if (!tc_encode_act_map(&maps.dy, (const bf16*)dy, NOUT, lddy, dims, strides, 32)) {
      set_error("%s: cuTensorMapEncodeTiled(dy) failed", who);
      return MVD_ERR_CUDA;
    }
    maps.w = maps.dy;
  }
  P.dw_col = dw_col;
  MVD_CUDA(cudaMemsetAsync(dw_col, 0, sizeof(float) * NOUT * Cin * 27, (cudaStream_t)stream));
  // tiles + slack for the junk M blocks of the last stage (M = 128 rows = 128 / KPAD tile-sized blocks)
  const size_t tile_bytes = (size_t)128 * kpad * 2;
  const size_t smem = (size_t)3 * 128 * 64 + (size_t)(kStages + 128 / kpad) * tile_bytes + 1024;
  int grid = num_sms();
  if (grid > P.total_tiles) grid = P.total_tiles;
  static bool a1 = false, a2 = false;
  int rc;
  if (Cin == 1) {
    if ((rc = set_smem_attr(stem_wgrad_kernel<1>, smem, &a1, who))) return rc;
    stem_wgrad_kernel<1><<<grid, kThreadsWgrad, smem, (cudaStream_t)stream>>>(maps, P);
  } else {
    if ((rc = set_smem_attr(stem_wgrad_kernel<2>, smem, &a2, who))) return rc;
    stem_wgrad_kernel<2><<<grid, kThreadsWgrad, smem, (cudaStream_t)stream>>>(maps, P);
  }
  MVD_LAUNCH_CHECK(who);
  return MVD_OK;
}

}  // extern "C"
