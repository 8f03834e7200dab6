// conv_tc_wgrad_halo.cu -- tcgen05 weight gradient for 3x3x3 / stride 1 / pad 1 convolutions with a SLIDING WINDOW of
// halo'd input planes in shared memory.
//
//   dw[co][ci][tap] = sum_v y[v, co] * x[v + tap - 1, ci]
//
// The tap-by-tap wgrad (conv_tc_wgrad.cu) fetches 27 shifted x bricks per 128-voxel brick of y and is bound by the TMA
// request rate.  Here a CTA owns one "set" = (32 input channels, 32 output channels) and walks work items
// (sample, 16h x 8w column, run of d planes): per y plane it loads ONE new x plane [18 h][10 w][32 ch] (SWIZZLE_64B)
// and ONE y brick [128 v][32 co]; the planes d-1, d, d+1 stay resident.  All 27 taps read the planes through shifted
// MN-major descriptors (address-based swizzle, scripts/umma_probe.py):
//   A (M = 128 = 4 blocks of 32 channels): block b = tap ox = b of the current (oz, oy): LBO = one voxel row (64 B);
//       K rows = voxels, 8-row groups = h lines, SBO = plane row pitch (10 voxels = 640 B); ox = 3 is a junk block.
//   B (N = 32): the y brick, MN-major, dense.
//   one MMA (K = 16 voxels = 2 h lines) per (oz, oy, k-step): 9 accumulators x 32 columns stay in TMEM for the whole
//   kernel; one epilogue of fp32 atomics into dw.
// TMA rows per y brick: 27*128 + 128  ->  180 + 128.
#include "conv_common.cuh"
#include "tc_common.cuh"

namespace mvd {
namespace {

using namespace tc;

constexpr int kThreads = 192;
constexpr int TILE_W = 8, TILE_H = 16, HALO_W = 10, HALO_H = 18;
constexpr int ROWB = 64;                                   // 32 bf16 channels
constexpr int PLANE_TX = HALO_W * HALO_H * ROWB;           // 11520
constexpr int PLANE_BYTES = 12 * 1024;                     // slot pitch
constexpr int BRICK_BYTES = 128 * ROWB;                    // 8 KB
constexpr int kPlaneRing = 6, kBrickRing = 3;

struct alignas(64) WhMaps {
  CUtensorMap x;   // (C, W, H, D, B) box (32, 10, 18, 1, 1)
  CUtensorMap y;   // (C, W, H, D, B) box (32, 8, 16, 1, 1)
};

struct WhParams {
  int B, D, H, W, tiles_w, tiles_h;
  int Cin, Cout, cblocks, nblocks;      // 32-channel blocks
  int dchunk, dchunks;                  // d planes per work item
  int items_per_set, ctas_per_set;
  float* dw;
};

__global__ void __launch_bounds__(kThreads, 1) wgrad_halo_kernel(const __grid_constant__ WhMaps maps,
                                                                 const __grid_constant__ WhParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar_pfull[kPlaneRing], bar_pempty[kPlaneRing], bar_bfull[kBrickRing], bar_bempty[kBrickRing],
      bar_done;
  __shared__ uint32_t s_tmem_base;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_p = smem;
  uint8_t* smem_b = smem + kPlaneRing * PLANE_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int set = blockIdx.x / P.ctas_per_set, rank = blockIdx.x % P.ctas_per_set;
  const int cb = set / P.nblocks, nb = set % P.nblocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kPlaneRing; ++s) { mbar_init(&bar_pfull[s], 1); mbar_init(&bar_pempty[s], 1); }
    for (int s = 0; s < kBrickRing; ++s) { mbar_init(&bar_bfull[s], 1); mbar_init(&bar_bempty[s], 1); }
    mbar_init(&bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&s_tmem_base, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  // work item -> (sample, column, d range)
  auto decode = [&](int item, int& b, int& h0, int& w0, int& d_lo, int& d_hi) {
    const int dc = item % P.dchunks;
    int m = item / P.dchunks;
    w0 = (m % P.tiles_w) * TILE_W; m /= P.tiles_w;
    h0 = (m % P.tiles_h) * TILE_H; m /= P.tiles_h;
    b = m;
    d_lo = dc * P.dchunk;
    d_hi = d_lo + P.dchunk;
    if (d_hi > P.D) d_hi = P.D;
  };
  const bool has_work = rank < P.items_per_set;

  if (warp == 0) {
    if (has_work && elect_one_sync()) {
      // ================= producer: x planes and y bricks =================
      int ps = 0, bs = 0;
      uint32_t pph = 0, bph = 0;
      for (int item = rank; item < P.items_per_set; item += P.ctas_per_set) {
        int b, h0, w0, d_lo, d_hi;
        decode(item, b, h0, w0, d_lo, d_hi);
        // planes d_lo-1 .. d_hi ; brick d follows plane d+1 so that the consumer never waits on an unissued load
        for (int p = d_lo - 1; p <= d_hi; ++p) {
          mbar_wait(&bar_pempty[ps], pph ^ 1, 51);
          mbar_arrive_expect_tx(&bar_pfull[ps], (uint32_t)PLANE_TX);
          tma_load_5d(&maps.x, smem_p + ps * PLANE_BYTES, &bar_pfull[ps], cb * 32, w0 - 1, h0 - 1, p, b);
          if (++ps == kPlaneRing) { ps = 0; pph ^= 1; }
          const int d = p - 1;
          if (d >= d_lo) {
            mbar_wait(&bar_bempty[bs], bph ^ 1, 52);
            mbar_arrive_expect_tx(&bar_bfull[bs], (uint32_t)BRICK_BYTES);
            tma_load_5d(&maps.y, smem_b + bs * BRICK_BYTES, &bar_bfull[bs], nb * 32, w0, h0, d, b);
            if (++bs == kBrickRing) { bs = 0; bph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (has_work && elect_one_sync()) {
      // ================= MMA issuer =================
      const uint32_t idesc = make_idesc_bf16(128, 32, 1, 1);
      // A: MN-major SWIZZLE_64B, LBO = 64 B (next ox tap), SBO = 640 B (next h line)
      const uint32_t a_hi = (uint32_t)(make_smem_desc(0, ROWB, HALO_W * ROWB, kLayoutSw64) >> 32);
      const uint32_t a_lo_fields = ((uint32_t)(ROWB >> 4) << 16);
      // B: MN-major SWIZZLE_64B dense brick, SBO = 512 B
      const uint32_t b_hi = (uint32_t)(make_smem_desc(0, BRICK_BYTES, 8 * ROWB, kLayoutSw64) >> 32);
      const uint32_t b_lo_fields = ((uint32_t)(BRICK_BYTES >> 4) << 16);
      const uint32_t p_base = smem_u32(smem_p) >> 4, b_base = smem_u32(smem_b) >> 4;
      int ps = 0, bs = 0;             // ring position of the OLDEST live plane / of the current brick
      uint32_t pph = 0, bph = 0;
      bool first = true;
      for (int item = rank; item < P.items_per_set; item += P.ctas_per_set) {
        int b, h0, w0, d_lo, d_hi;
        decode(item, b, h0, w0, d_lo, d_hi);
        // the first two planes of the item (d_lo-1, d_lo)
        int s0 = ps, s1 = ps + 1;
        uint32_t ph0 = pph, ph1 = pph;
        if (s1 >= kPlaneRing) { s1 -= kPlaneRing; ph1 ^= 1; }
        mbar_wait(&bar_pfull[s0], ph0, 53);
        mbar_wait(&bar_pfull[s1], ph1, 54);
        for (int d = d_lo; d < d_hi; ++d) {
          int s2 = s1 + 1;
          uint32_t ph2 = ph1;
          if (s2 >= kPlaneRing) { s2 -= kPlaneRing; ph2 ^= 1; }
          mbar_wait(&bar_pfull[s2], ph2, 55);
          mbar_wait(&bar_bfull[bs], bph, 56);
          tcgen05_fence_after();
          const uint32_t blo = (b_base + (uint32_t)bs * (BRICK_BYTES >> 4)) | b_lo_fields;
          const int slots[3] = {s0, s1, s2};
#pragma unroll
          for (int oz = 0; oz < 3; ++oz) {
            const uint32_t plo = p_base + (uint32_t)slots[oz] * (PLANE_BYTES >> 4);
#pragma unroll
            for (int oy = 0; oy < 3; ++oy) {
              const uint32_t d_tmem = tmem_base + (uint32_t)((oz * 3 + oy) * 32);
#pragma unroll
              for (int j = 0; j < 8; ++j) {   // 16 voxels = h lines 2j, 2j+1 of the brick
                const uint32_t alo = ((plo + (uint32_t)(((2 * j + oy) * HALO_W * ROWB) >> 4)) & 0x3FFF) | a_lo_fields;
                const uint64_t adesc = ((uint64_t)a_hi << 32) | (uint64_t)alo;
                const uint64_t bdesc = ((uint64_t)b_hi << 32) | (uint64_t)(blo + (uint32_t)(j * ((2 * 8 * ROWB) >> 4)));
                umma_bf16(d_tmem, adesc, bdesc, idesc, (first && j == 0) ? 0u : 1u);
              }
            }
          }
          first = false;
          umma_commit(&bar_pempty[s0]);    // plane d-1 is no longer needed
          umma_commit(&bar_bempty[bs]);
          if (++bs == kBrickRing) { bs = 0; bph ^= 1; }
          s0 = s1; ph0 = ph1; s1 = s2; ph1 = ph2;
          if (++ps == kPlaneRing) { ps = 0; pph ^= 1; }
        }
        // the last two planes of the item are released without further use
        umma_commit(&bar_pempty[s0]);
        umma_commit(&bar_pempty[s1]);
        ps += 2;
        if (ps >= kPlaneRing) { ps -= kPlaneRing; pph ^= 1; }
      }
      umma_commit(&bar_done);
    }
  } else if (has_work) {
    // ================= epilogue (warps 2..5) =================
    const int q = warp & 3;
    mbar_wait(&bar_done, 0, 57);
    tcgen05_fence_after();
    const int m = q * 32 + lane;          // row = ox * 32 + channel
    const int ox = m >> 5, ci = cb * 32 + (m & 31);
    for (int g = 0; g < 9; ++g) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 32), v);
      tmem_ld_wait();
      if (ox < 3) {
        const int tap = g * 3 + ox;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int co = nb * 32 + e;
          atomicAdd(&P.dw[((long long)co * P.Cin + ci) * 27 + tap], __uint_as_float(v[e]));
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

bool encode5(CUtensorMap* m, const bf16* base, int C, long long ld, int W, int H, int D, int B, int bw, int bh) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
  cuuint64_t gstr[4] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * W * 2, (cuuint64_t)ld * W * H * 2,
                        (cuuint64_t)ld * W * H * D * 2};
  cuuint32_t box[5] = {32, (cuuint32_t)bw, (cuuint32_t)bh, 1, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ==
         CUDA_SUCCESS;
}

}  // namespace

bool tc_wgrad_halo_supported(const mvd_conv3d_args* a) {
  if (!(a->kd == 3 && a->kh == 3 && a->kw == 3 && a->sd == 1 && a->sh == 1 && a->sw == 1 && a->pd == 1 && a->ph == 1 &&
        a->pw == 1))
    return false;
  if (a->Cin % 32 || a->Cout % 32 || a->ldx % 8 || a->ldy % 8) return false;
  if (((uintptr_t)a->x & 15) || ((uintptr_t)a->y & 15)) return false;
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("MVD_NO_WGRAD_HALO");
    enabled = (e && e[0] == '1') ? 0 : 1;
  }
  // measured (microbench, round 1): the sliding-window kernel wins for Cout == 32 and for small volumes; for Cout >= 64
  // on >= 32^3 x 2 voxels the tap-by-tap kernel with N = 64/128 MMAs is on par or slightly ahead
  const long long vox = (long long)a->B * a->Do * a->Ho * a->Wo;
  if (a->Cout >= 64 && vox >= 65536) return false;
  return enabled == 1 && get_encode_tiled() != nullptr;
}

int tc_wgrad_halo(const mvd_conv3d_args* a, cudaStream_t st) {
  MVD_CUDA(cudaMemsetAsync(a->dw, 0, sizeof(float) * (size_t)a->Cout * a->Cin * 27, st));
  WhMaps maps;
  WhParams P;
  memset(&P, 0, sizeof(P));
  if (!encode5(&maps.x, (const bf16*)a->x, a->Cin, a->ldx, a->Wi, a->Hi, a->Di, a->B, HALO_W, HALO_H) ||
      !encode5(&maps.y, (const bf16*)a->y, a->Cout, a->ldy, a->Wo, a->Ho, a->Do, a->B, TILE_W, TILE_H)) {
    set_error("conv3d_wgrad(tcgen05 halo): cuTensorMapEncodeTiled failed");
    return MVD_ERR_CUDA;
  }
  P.B = a->B; P.D = a->Do; P.H = a->Ho; P.W = a->Wo;
  P.tiles_w = cdiv(P.W, TILE_W); P.tiles_h = cdiv(P.H, TILE_H);
  P.Cin = a->Cin; P.Cout = a->Cout; P.cblocks = a->Cin / 32; P.nblocks = a->Cout / 32;
  const int sets = P.cblocks * P.nblocks;
  int ctas_per_set = num_sms() / sets;
  if (ctas_per_set < 1) ctas_per_set = 1;
  const int columns = P.B * P.tiles_h * P.tiles_w;
  // split the depth so that every CTA of a set gets >= ~4 work items, but keep runs >= 8 planes (2 halo planes per run)
  int dchunks = cdiv(ctas_per_set * 4, columns);
  if (dchunks < 1) dchunks = 1;
  int dchunk = cdiv(P.D, dchunks);
  if (dchunk < 8) dchunk = (P.D < 8) ? P.D : 8;
  dchunks = cdiv(P.D, dchunk);
  P.dchunk = dchunk; P.dchunks = dchunks;
  P.items_per_set = columns * dchunks;
  if (ctas_per_set > P.items_per_set) ctas_per_set = P.items_per_set;
  P.ctas_per_set = ctas_per_set;
  P.dw = a->dw;
  const size_t smem = (size_t)kPlaneRing * PLANE_BYTES + (size_t)kBrickRing * BRICK_BYTES + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("conv3d_wgrad(tcgen05 halo): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return MVD_ERR_CUDA;
    }
    attr_done = true;
  }
  wgrad_halo_kernel<<<sets * ctas_per_set, kThreads, smem, st>>>(maps, P);
  MVD_LAUNCH_CHECK("conv3d_wgrad(tcgen05 halo)");
  if (a->dbias)
    return mvd_channel_sum(a->y, a->ldy, (long long)a->B * a->Do * a->Ho * a->Wo, a->Cout, a->dbias, (mvd_stream_t)st);
  return MVD_OK;
}

}  // namespace mvd
