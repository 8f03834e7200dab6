// conv_tc_subpixel.cu -- tcgen05 data gradient of the stride-2 3x3x3 convolution (pad 1) at the top of the encoder
// (32 -> 64 channels, 128^3 -> 64^3: dx has 32 channels at full resolution), as a SUB-PIXEL convolution over the dy
// lattice.
//
//   dx[2a + r] = sum over shifts s <= r (component-wise, s in {0,1}^3) of  W[tap(r, s)]^T dy[a + s]
//   per axis: (r=0, s=0) -> tap 1;  (r=1, s=0) -> tap 2;  (r=1, s=1) -> tap 0          (27 (r, s) pairs = 27 taps)
//
// The parity-class form (conv_tc.cu) runs 8 separate strided GEMMs that fetch every dy brick once per tap through TMA and
// store 64-byte rows at a 128-byte pitch: 0.33-0.49 ms for this layer (120-180 TFLOP/s).  Here a CTA takes a brick of
// 128 dy anchors (16 h x 8 w of plane d) and produces ALL EIGHT parity classes of the 2x2x2 dx block under it:
//   * accumulator = [128 anchors][8 classes x 32 channels] = 256 TMEM columns (double-buffered: 512);
//   * for each shift s the A operand is the dy brick displaced by s -- planes d and d+1 of a halo'd plane ring
//     ([18][10][64] boxes with origin (h0, w0)), in-plane shifts by moving the descriptor start (address-based swizzle);
//   * the taps that share a shift are folded into the N axis: s = 000 feeds all 8 classes with ONE N = 256 MMA,
//     s = 100 four classes (N = 128), s = 010 two runs of two (N = 64), ... 14 MMAs per K step instead of 27 N = 32 ones
//     (696 instead of 1080 cycles); the 27 weight tiles (110 KB) stay resident in shared memory in exactly that order;
//   * planes stream along d (each fetched once per run of DR anchors), 8 epilogue warps (two per TMEM lane quarter)
//     convert and store, optionally ACCUMULATING into dx (the skip-connection gradient, old values prefetched).
// Coverage: k = 3, s = 2, p = 1 on all axes, Cin = 32 (dx channels), Cout = 64; everything else stays on conv_tc.cu.
#include "conv_common.cuh"
#include "tc_common.cuh"
#include "tc_epilogue.cuh"

namespace mvd {
namespace {

using namespace tc;

constexpr int kThreads = 352;                 // warp 0 planes, 1 MMA issuer, 2 weights, 3..10 epilogue
constexpr int TILE_W = 8, TILE_H = 16, HALO_W = 10, HALO_H = 18, PLANE_ROWS = HALO_W * HALO_H;
constexpr int KC = 64, ROWB = KC * 2;
constexpr int PLANE_TX = PLANE_ROWS * ROWB;                       // 23040
constexpr int PLANE_BYTES = (PLANE_TX + 1023) & ~1023;            // 23552
constexpr int NCH = 32;                                           // dx channels = columns per class
constexpr int WBLOCK_BYTES = NCH * ROWB;                          // one tap: 32 rows x 128 B = 4 KB
constexpr int W_BYTES = 27 * WBLOCK_BYTES;                        // 110592
constexpr int kRing = 4;

struct alignas(64) SpMaps {
  CUtensorMap a;   // dy: (C, W, H, D, B) box (64, 10, 18, 1, 1)
  CUtensorMap b;   // weights wd [27 * 32 rows][64] box (64, 32)
};

struct SpParams {
  int B, Do, Ho, Wo;        // dy lattice
  int Di, Hi, Wi;           // dx extents
  int tiles_w, tiles_h, druns, DR, total_items;
  bf16* out;
  long long sb, sd, sh, sw; // dx element strides
  int accumulate;
  int wtap[27];             // tap index held by the j-th 32-row weight block (shift-major order, see below)
};

// MMA schedule of one K step: {plane (0: d, 1: d+1), in-plane offset (sh, sw), first weight block, blocks, first class}
struct SpMma { int plane, sh, sw, blk, n, cls; };
__device__ constexpr SpMma kSched[14] = {
    {0, 0, 0, 0, 8, 0},                                       // s = 000: all classes, N = 256 (initialises the tile)
    {1, 0, 0, 8, 4, 4},                                       // s = 100: rd = 1 -> classes 4..7
    {0, 1, 0, 12, 2, 2},  {0, 1, 0, 14, 2, 6},                // s = 010: rh = 1 -> {2,3}, {6,7}
    {0, 0, 1, 16, 1, 1},  {0, 0, 1, 17, 1, 3},  {0, 0, 1, 18, 1, 5},  {0, 0, 1, 19, 1, 7},   // s = 001: rw = 1
    {1, 1, 0, 20, 2, 6},                                      // s = 110: {6,7}
    {1, 0, 1, 22, 1, 5},  {1, 0, 1, 23, 1, 7},                // s = 101: {5,7}
    {0, 1, 1, 24, 1, 3},  {0, 1, 1, 25, 1, 7},                // s = 011: {3,7}
    {1, 1, 1, 26, 1, 7}};                                     // s = 111: {7}

__global__ void __launch_bounds__(kThreads, 1) conv_subpixel_dgrad_kernel(const __grid_constant__ SpMaps maps,
                                                                           const __grid_constant__ SpParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar_pfull[kRing], bar_pempty[kRing], bar_wres, bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(16) uint8_t s_stage[8][2048];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem;
  uint8_t* smem_p = smem + W_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRing; ++s) { mbar_init(&bar_pfull[s], 1); mbar_init(&bar_pempty[s], 1); }
    mbar_init(&bar_wres, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&bar_tfull[a], 1); mbar_init(&bar_tempty[a], 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&s_tmem_base, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  // item -> (sample, brick, run of DR anchor planes)
  auto decode = [&](int item, int& b, int& h0, int& w0, int& d_lo, int& d_hi) {
    const int dr = item % P.druns;
    int m = item / P.druns;
    w0 = (m % P.tiles_w) * TILE_W; m /= P.tiles_w;
    h0 = (m % P.tiles_h) * TILE_H; m /= P.tiles_h;
    b = m;
    d_lo = dr * P.DR;
    d_hi = min(d_lo + P.DR, P.Do);
  };

  if (warp == 0) {
    if (elect_one_sync()) {
      // ================= plane producer: dy planes d_lo .. d_hi (the last one only as the "d + 1" operand) ==========
      int slot = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < P.total_items; item += gridDim.x) {
        int b, h0, w0, d_lo, d_hi;
        decode(item, b, h0, w0, d_lo, d_hi);
        for (int p = d_lo; p <= d_hi; ++p) {
          mbar_wait(&bar_pempty[slot], phase ^ 1, 61);
          mbar_arrive_expect_tx(&bar_pfull[slot], (uint32_t)PLANE_TX);
          tma_load_5d(&maps.a, smem_p + (size_t)slot * PLANE_BYTES, &bar_pfull[slot], 0, w0, h0, p, b);
          if (++slot == kRing) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    if (elect_one_sync()) {
      // ================= weights: 27 tap tiles in schedule order, once =================
      mbar_arrive_expect_tx(&bar_wres, (uint32_t)W_BYTES);
      for (int j = 0; j < 27; ++j)
        tma_load_2d(&maps.b, smem_w + (size_t)j * WBLOCK_BYTES, &bar_wres, 0, P.wtap[j] * NCH);
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      // ================= MMA issuer =================
      constexpr uint32_t A_SBO = HALO_W * ROWB, B_SBO = 8 * ROWB;
      const uint32_t a_hi = (uint32_t)(make_smem_desc(0, 16, A_SBO, kLayoutSw128) >> 32);
      const uint32_t b_hi = (uint32_t)(make_smem_desc(0, 16, B_SBO, kLayoutSw128) >> 32);
      const uint32_t p_base = (smem_u32(smem_p) >> 4) | (1u << 16);
      const uint32_t w_base = (smem_u32(smem_w) >> 4) | (1u << 16);
      const int total_items = P.total_items, gstride = gridDim.x;
      int acc = 0, slot = 0;
      uint32_t accphase = 0, pphase = 0;
      mbar_wait(&bar_wres, 0, 62);
      tcgen05_fence_after();
      for (int item = blockIdx.x; item < total_items; item += gstride) {
        int b, h0, w0, d_lo, d_hi;
        decode(item, b, h0, w0, d_lo, d_hi);
        mbar_wait(&bar_pfull[slot], pphase, 63);         // plane d_lo
        for (int d = d_lo; d < d_hi; ++d) {
          int slot1 = slot + 1;
          uint32_t ph1 = pphase;
          if (slot1 == kRing) { slot1 = 0; ph1 ^= 1; }
          mbar_wait(&bar_pfull[slot1], ph1, 64);         // plane d + 1
          mbar_wait(&bar_tempty[acc], accphase ^ 1, 65);
          tcgen05_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
          const uint32_t plo0 = p_base + (uint32_t)slot * (uint32_t)(PLANE_BYTES >> 4);
          const uint32_t plo1 = p_base + (uint32_t)slot1 * (uint32_t)(PLANE_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) {
#pragma unroll
            for (int i = 0; i < 14; ++i) {
              const SpMma e = kSched[i];
              const uint32_t alo = (e.plane ? plo1 : plo0) + (uint32_t)(((e.sh * HALO_W + e.sw) * ROWB) >> 4) + 2u * k;
              const uint32_t blo = w_base + (uint32_t)((e.blk * WBLOCK_BYTES) >> 4) + 2u * k;
              umma_bf16(d_tmem + (uint32_t)(e.cls * NCH), ((uint64_t)a_hi << 32) | alo, ((uint64_t)b_hi << 32) | blo,
                        make_idesc_bf16(128, e.n * NCH, 0, 0), (i == 0 && k == 0) ? 0u : 1u);
            }
          }
          umma_commit(&bar_tfull[acc]);
          umma_commit(&bar_pempty[slot]);                 // plane d is done; plane d + 1 becomes the next anchor's plane
          acc ^= 1;
          if (acc == 0) accphase ^= 1;
          slot = slot1;
          pphase = ph1;
        }
        umma_commit(&bar_pempty[slot]);                   // the run's last plane (d_hi) was only ever a "d + 1" operand
        if (++slot == kRing) { slot = 0; pphase ^= 1; }
      }
    }
  } else if (warp >= 3) {
    // ================= epilogue: 8 warps, two per TMEM lane quarter, alternating classes =================
    const int q = warp & 3, half = (warp - 3) >> 2;
    uint8_t* stage = s_stage[warp - 3];
    int acc = 0;
    uint32_t accphase = 0;
    LaneStats<0> nostats;
    const bool accum = P.accumulate != 0;
    for (int item = blockIdx.x; item < P.total_items; item += gridDim.x) {
      int b, h0, w0, d_lo, d_hi;
      decode(item, b, h0, w0, d_lo, d_hi);
      for (int d = d_lo; d < d_hi; ++d) {
        // destination of channel 0 for row R of this warp's 32 anchors and parity class c
        auto row_ptr = [&](int R, int c) -> bf16* {
          const int rr = q * 32 + R;
          const int z = 2 * d + (c >> 2), y = 2 * (h0 + (rr >> 3)) + ((c >> 1) & 1), x = 2 * (w0 + (rr & 7)) + (c & 1);
          return (z < P.Di && y < P.Hi && x < P.Wi && h0 + (rr >> 3) < P.Ho && w0 + (rr & 7) < P.Wo)
                     ? P.out + (long long)b * P.sb + (long long)z * P.sd + (long long)y * P.sh + (long long)x * P.sw
                     : nullptr;
        };
        bf16x8 old[4];
        bool has[4] = {false, false, false, false};
        if (accum) prefetch_rows(lane, [&](int R) { return row_ptr(R, half); }, old, has);
        mbar_wait(&bar_tfull[acc], accphase, 66);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
        for (int c = half; c < 8; c += 2) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + (uint32_t)(c * NCH), v);
          tmem_ld_wait();
          uint32_t w2[16];
          epilogue_chunk<0>(v, nullptr, nostats, 0, true, w2);
          if (accum) {
            bf16x8 cur[4];
            bool chas[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { cur[i] = old[i]; chas[i] = has[i]; }
            if (c + 2 < 8) prefetch_rows(lane, [&](int R) { return row_ptr(R, c + 2); }, old, has);
            store_rows_accumulate_packed(stage, lane, w2, [&](int R) { return row_ptr(R, c); }, cur, chas);
          } else {
            store_rows_coalesced_packed(stage, lane, w2, [&](int R) { return row_ptr(R, c); }, false);
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_tempty[acc]);
        acc ^= 1;
        if (acc == 0) accphase ^= 1;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace

bool tc_subpixel_dgrad_supported(const mvd_conv3d_args* a) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("MVD_NO_SUBPIXEL");
    enabled = (e && e[0] == '1') ? 0 : 1;
  }
  if (!enabled) return false;
  if (!(a->kd == 3 && a->kh == 3 && a->kw == 3 && a->sd == 2 && a->sh == 2 && a->sw == 2 && a->pd == 1 && a->ph == 1 &&
        a->pw == 1))
    return false;
  if (a->Cin != NCH || a->Cout != KC || a->bias) return false;
  if (a->ldx % 8 || a->ldy % 8 || ((uintptr_t)a->x & 15) || ((uintptr_t)a->y & 15) || ((uintptr_t)a->w & 15)) return false;
  return get_encode_tiled() != nullptr;
}

// a->w is the dgrad packing [tap][Cin][Cout]
int tc_subpixel_dgrad(const mvd_conv3d_args* a, cudaStream_t st) {
  PFN_encodeTiled enc = get_encode_tiled();
  const char* who = "conv3d_dgrad(tcgen05 sub-pixel)";
  if (!enc) { set_error("%s: no cuTensorMapEncodeTiled", who); return MVD_ERR_CUDA; }
  SpMaps maps;
  SpParams P;
  memset(&P, 0, sizeof(P));
  {
    const long long ld = a->ldy;
    cuuint64_t gdim[5] = {(cuuint64_t)a->Cout, (cuuint64_t)a->Wo, (cuuint64_t)a->Ho, (cuuint64_t)a->Do, (cuuint64_t)a->B};
    cuuint64_t gstr[4] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * a->Wo * 2, (cuuint64_t)ld * a->Wo * a->Ho * 2,
                          (cuuint64_t)ld * a->Wo * a->Ho * a->Do * 2};
    cuuint32_t box[5] = {KC, HALO_W, HALO_H, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&maps.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)a->y, gdim, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, tc_l2_promotion(ROWB, ld * 2),
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("%s: cuTensorMapEncodeTiled(planes) failed (%d)", who, (int)r); return MVD_ERR_CUDA; }
  }
  if (!tc_encode_w_map(&maps.b, (const bf16*)a->w, (long long)27 * NCH, KC, NCH, KC)) {
    set_error("%s: cuTensorMapEncodeTiled(weights) failed", who);
    return MVD_ERR_CUDA;
  }
  // weight blocks in schedule order: for every shift s (000, 100, 010, 001, 110, 101, 011, 111) the classes r >= s in
  // ascending class index c = rd*4 + rh*2 + rw; per axis tap = 1 (r=0), 2 (r=1, s=0), 0 (r=1, s=1)
  static const int shifts[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};
  int nb = 0;
  for (int si = 0; si < 8; ++si)
    for (int c = 0; c < 8; ++c) {
      const int r[3] = {c >> 2, (c >> 1) & 1, c & 1};
      bool ok = true;
      int t[3];
      for (int ax = 0; ax < 3; ++ax) {
        if (shifts[si][ax] > r[ax]) ok = false;
        t[ax] = (r[ax] == 0) ? 1 : (shifts[si][ax] ? 0 : 2);
      }
      if (ok) P.wtap[nb++] = (t[0] * 3 + t[1]) * 3 + t[2];
    }
  if (nb != 27) { set_error("%s: internal schedule error", who); return MVD_ERR_CUDA; }
  P.B = a->B; P.Do = a->Do; P.Ho = a->Ho; P.Wo = a->Wo;
  P.Di = a->Di; P.Hi = a->Hi; P.Wi = a->Wi;
  P.tiles_w = cdiv(a->Wo, TILE_W); P.tiles_h = cdiv(a->Ho, TILE_H);
  P.DR = a->Do < 4 ? a->Do : 4;
  P.druns = cdiv(a->Do, P.DR);
  P.total_items = a->B * P.tiles_h * P.tiles_w * P.druns;
  P.out = (bf16*)a->x;
  P.sw = a->ldx; P.sh = (long long)a->ldx * a->Wi; P.sd = P.sh * a->Hi; P.sb = P.sd * a->Di;
  P.accumulate = a->accumulate;
  const size_t smem = (size_t)W_BYTES + (size_t)kRing * PLANE_BYTES + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_subpixel_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(e));
      return MVD_ERR_CUDA;
    }
    attr_done = true;
  }
  int grid = num_sms();
  if (grid > P.total_items) grid = P.total_items;
  conv_subpixel_dgrad_kernel<<<grid, kThreads, smem, st>>>(maps, P);
  MVD_LAUNCH_CHECK(who);
  return MVD_OK;
}

}  // namespace mvd
