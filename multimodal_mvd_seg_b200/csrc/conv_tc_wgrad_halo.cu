// conv_tc_wgrad_halo.cu -- tcgen05 weight gradient for 3x3x3 / stride 1 / pad 1 convolutions with a SLIDING WINDOW of
// halo'd input planes in shared memory and the three depth taps folded into the MMA's N axis.
//
//   dw[co][ci][tap] = sum_v y[v, co] * x[v + tap - 1, ci]
//
// The tap-by-tap wgrad (conv_tc_wgrad.cu) fetches 27 shifted x bricks per 128-voxel brick of y and is bound by the TMA
// request rate.  Here a CTA owns one "set" = (32 input channels, 32 output channels) and walks work items
// (sample, 16h x 8w column, run of d planes).  The loop is anchored on the x plane p ([18 h][10 w][32 ch], SWIZZLE_64B,
// loaded once): it meets the y bricks of planes p-1, p, p+1 (depth taps oz = 2, 1, 0), which sit in three consecutive
// slots of a (mirrored) brick ring, so one MMA sees them as a single MN-major B operand of N = 96:
//   A (M = 128 = 4 blocks of 32 channels): block = in-plane tap ox (LBO = one voxel row = 64 B; block 3 is junk);
//       K rows = voxels, 8-row groups = h lines, SBO = plane row pitch (10 voxels = 640 B); oy by the start address.
//   B (N = 96 = 3 bricks x 32 co): LBO = brick pitch, SBO = 512 B.
//   one MMA (K = 16 voxels = 2 h lines) per (oy, k-step): 3 accumulators x 96 columns stay in TMEM for the whole
//   kernel (zeroed once by the epilogue warps, every MMA accumulates); one epilogue of 16-byte vector reductions into
//   the [tap][ci][co] scratch (see conv_tc_wgrad.cu).
// A tcgen05.mma M128 x N x K16 costs max(N/2, 32 + N/4) cycles: 24 MMAs of 56 cycles per y brick here instead of 72 of
// 40 in the unfolded form.  At the ends of a run the operand narrows to the bricks that exist (N = 64 / 32).
// Mirrored ring: brick c lives in slot c % R and, when c % R < 2, also in slot R + c % R, so the triple starting at any
// slot is contiguous.  TMA rows per y brick: 27*128 + 128  ->  180 + 128 (+ 43 for the mirror copies).
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
constexpr int kPlaneRing = 4, kBrickRing = 6;              // + 2 mirror slots behind the brick ring

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
  long long slice_stride;   // deterministic mode: CTA `rank` of every set stores into slice `rank`; 0 = one slice, red.add
};

__device__ __forceinline__ void tmem_st_zero_32x32b_x32(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(kThreads, 1) wgrad_halo_kernel(const __grid_constant__ WhMaps maps,
                                                                 const __grid_constant__ WhParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar_pfull[kPlaneRing], bar_pempty[kPlaneRing], bar_bfull[kBrickRing], bar_bempty[kBrickRing],
      bar_zero, bar_done;
  __shared__ uint32_t s_tmem_base;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem;                                            // kBrickRing + 2 bricks
  uint8_t* smem_p = smem + (kBrickRing + 2) * BRICK_BYTES;           // kPlaneRing planes
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int set = blockIdx.x / P.ctas_per_set, rank = blockIdx.x % P.ctas_per_set;
  const int cb = set / P.nblocks, nb = set % P.nblocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kPlaneRing; ++s) { mbar_init(&bar_pfull[s], 1); mbar_init(&bar_pempty[s], 1); }
    for (int s = 0; s < kBrickRing; ++s) { mbar_init(&bar_bfull[s], 1); mbar_init(&bar_bempty[s], 1); }
    mbar_init(&bar_zero, 4);
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
      // ================= producer: y bricks (mirrored ring) and x planes =================
      int ps = 0, bs = 0;
      uint32_t pph = 0, bph = 0;
      for (int item = rank; item < P.items_per_set; item += P.ctas_per_set) {
        int b, h0, w0, d_lo, d_hi;
        decode(item, b, h0, w0, d_lo, d_hi);
        // anchor p needs plane p and the bricks up to p + 1: brick p + 1 is issued right before plane p
        for (int p = d_lo - 1; p <= d_hi; ++p) {
          const int q = p + 1;
          if (q < d_hi) {
            const bool mirror = bs < 2;
            mbar_wait(&bar_bempty[bs], bph ^ 1, 52);
            mbar_arrive_expect_tx(&bar_bfull[bs], (uint32_t)(mirror ? 2 * BRICK_BYTES : BRICK_BYTES));
            tma_load_5d(&maps.y, smem_b + bs * BRICK_BYTES, &bar_bfull[bs], nb * 32, w0, h0, q, b);
            if (mirror)
              tma_load_5d(&maps.y, smem_b + (kBrickRing + bs) * BRICK_BYTES, &bar_bfull[bs], nb * 32, w0, h0, q, b);
            if (++bs == kBrickRing) { bs = 0; bph ^= 1; }
          }
          mbar_wait(&bar_pempty[ps], pph ^ 1, 51);
          mbar_arrive_expect_tx(&bar_pfull[ps], (uint32_t)PLANE_TX);
          tma_load_5d(&maps.x, smem_p + ps * PLANE_BYTES, &bar_pfull[ps], cb * 32, w0 - 1, h0 - 1, p, b);
          if (++ps == kPlaneRing) { ps = 0; pph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (has_work && elect_one_sync()) {
      // ================= MMA issuer =================
      const uint32_t idesc1 = make_idesc_bf16(128, 32, 1, 1), idesc2 = make_idesc_bf16(128, 64, 1, 1),
                     idesc3 = make_idesc_bf16(128, 96, 1, 1);
      // A: MN-major SWIZZLE_64B, LBO = 64 B (next ox tap), SBO = 640 B (next h line)
      const uint32_t a_hi = (uint32_t)(make_smem_desc(0, ROWB, HALO_W * ROWB, kLayoutSw64) >> 32);
      const uint32_t a_lo_fields = ((uint32_t)(ROWB >> 4) << 16);
      // B: MN-major SWIZZLE_64B, N blocks = consecutive bricks (LBO = brick pitch), SBO = 512 B
      const uint32_t b_hi = (uint32_t)(make_smem_desc(0, BRICK_BYTES, 8 * ROWB, kLayoutSw64) >> 32);
      const uint32_t b_lo_fields = ((uint32_t)(BRICK_BYTES >> 4) << 16);
      const uint32_t p_base = smem_u32(smem_p) >> 4, b_base = smem_u32(smem_b) >> 4;
      int ps = 0;                     // ring slot of the anchor plane
      uint32_t pph = 0;
      int bnew = 0;                   // ring slot / parity of the next brick to arrive
      uint32_t bnew_ph = 0;
      mbar_wait(&bar_zero, 0, 58);    // accumulators zeroed by the epilogue warps
      tcgen05_fence_after();
      for (int item = rank; item < P.items_per_set; item += P.ctas_per_set) {
        int b, h0, w0, d_lo, d_hi;
        decode(item, b, h0, w0, d_lo, d_hi);
        for (int p = d_lo - 1; p <= d_hi; ++p) {
          // bricks of this item that meet plane p: q in [max(p-1, d_lo), min(p+1, d_hi-1)]; block j = q - (p - 1)
          const int q_lo = max(p - 1, d_lo), q_hi = min(p + 1, d_hi - 1);
          const int j_lo = q_lo - (p - 1), cnt = q_hi - q_lo + 1;
          if (p + 1 < d_hi) {          // brick p+1 arrives with this anchor
            mbar_wait(&bar_bfull[bnew], bnew_ph, 56);
            if (++bnew == kBrickRing) { bnew = 0; bnew_ph ^= 1; }
          }
          mbar_wait(&bar_pfull[ps], pph, 55);
          tcgen05_fence_after();
          // slot of brick p+1 is (bnew - 1) (just advanced) when it exists; base slot of the triple = slot(p+1) - 2
          // in ring arithmetic.  Track it through the slot of brick q_hi instead:
          //   slot(q_hi) = bnew - 1 - ((p + 1 < d_hi) ? 0 : (p + 1 - q_hi) - 1)  -- all bricks up to q_hi have arrived,
          //   and bnew - 1 is the newest arrived brick, which is min(p + 1, d_hi - 1) = q_hi.
          int s_hi = bnew - 1;
          if (s_hi < 0) s_hi += kBrickRing;
          int s_base = s_hi - (q_hi - (p - 1));      // slot of block 0 of the triple (may be negative -> wrap)
          if (s_base < 0) s_base += kBrickRing;
          const uint32_t blo0 = (b_base + (uint32_t)(s_base + j_lo) * (BRICK_BYTES >> 4)) | b_lo_fields;
          const uint32_t plo = p_base + (uint32_t)ps * (PLANE_BYTES >> 4);
          const uint32_t idesc = (cnt == 3) ? idesc3 : ((cnt == 2) ? idesc2 : idesc1);
          const uint32_t dcol = tmem_base + (uint32_t)(32 * j_lo);
#pragma unroll
          for (int oy = 0; oy < 3; ++oy) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {   // 16 voxels = h lines 2j, 2j+1 of the brick
              const uint32_t alo = ((plo + (uint32_t)(((2 * j + oy) * HALO_W * ROWB) >> 4)) & 0x3FFF) | a_lo_fields;
              const uint64_t adesc = ((uint64_t)a_hi << 32) | (uint64_t)alo;
              const uint64_t bdesc = ((uint64_t)b_hi << 32) | (uint64_t)(blo0 + (uint32_t)(j * ((2 * 8 * ROWB) >> 4)));
              umma_bf16(dcol + (uint32_t)(oy * 96), adesc, bdesc, idesc, 1u);
            }
          }
          umma_commit(&bar_pempty[ps]);
          if (++ps == kPlaneRing) { ps = 0; pph ^= 1; }
          // brick p-1 (block 0) has now met its last plane
          if (p - 1 >= d_lo) umma_commit(&bar_bempty[s_base]);
        }
        // brick d_hi-1 met plane d_hi last: released by the p = d_hi iteration above (p - 1 = d_hi - 1)
      }
      umma_commit(&bar_done);
    }
  } else if (has_work) {
    // ================= epilogue (warps 2..5) =================
    const int q = warp & 3;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int c = 0; c < 288; c += 32) tmem_st_zero_32x32b_x32(tlane + (uint32_t)c);
    tmem_st_wait();
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar_zero);
    mbar_wait(&bar_done, 0, 57);
    tcgen05_fence_after();
    const int m = q * 32 + lane;          // row = ox * 32 + channel
    const int ox = m >> 5, ci = cb * 32 + (m & 31);
    for (int g = 0; g < 9; ++g) {         // column group = (oy, block j): depth tap oz = 2 - j
      const int oy = g / 3, oz = 2 - (g - oy * 3);
      uint32_t v[32];
      tmem_ld_32x32b_x32(tlane + (uint32_t)(g * 32), v);
      tmem_ld_wait();
      if (ox < 3) {   // scratch [tap][ci][co]: 128 contiguous bytes per row -> 8 vector reductions
        const int tap = (oz * 3 + oy) * 3 + ox;
        float* dst = P.dw + (long long)rank * P.slice_stride + ((long long)tap * P.Cin + ci) * P.Cout + nb * 32;
        if (P.slice_stride) {
#pragma unroll
          for (int e = 0; e < 32; e += 4) *reinterpret_cast<uint4*>(dst + e) = make_uint4(v[e], v[e + 1], v[e + 2], v[e + 3]);
        } else {
#pragma unroll
          for (int e = 0; e < 32; e += 4) red_add_v4(dst + e, v[e], v[e + 1], v[e + 2], v[e + 3]);
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
             CU_TENSOR_MAP_SWIZZLE_64B, tc_l2_promotion(64, ld * 2), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ==
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
  static int all_sizes = -1;   // MVD_WGRAD_HALO_NARROW=1 restores the round-1 rule (Cout == 32 or small volumes only)
  if (all_sizes < 0) {
    const char* e = getenv("MVD_WGRAD_HALO_NARROW");
    all_sizes = (e && e[0] == '1') ? 0 : 1;
  }
  const long long vox = (long long)a->B * a->Do * a->Ho * a->Wo;
  if (!all_sizes && a->Cout >= 64 && vox >= 65536) return false;
  return enabled == 1 && get_encode_tiled() != nullptr;
}

bool wgrad_deterministic();   // conv_tc_wgrad.cu

// work decomposition shared by the launch and by the workspace query: sets of (32 ci, 32 co), CTAs per set, depth chunks
static void wgrad_halo_plan(const mvd_conv3d_args* a, WhParams& P) {
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
}

int tc_wgrad_halo_splits(const mvd_conv3d_args* a) {
  WhParams P;
  memset(&P, 0, sizeof(P));
  wgrad_halo_plan(a, P);
  return P.ctas_per_set;
}

int tc_wgrad_halo(const mvd_conv3d_args* a, cudaStream_t st) {
  float* scratch = nullptr;
  if (int rc0 = tc_wgrad_begin(a, st, &scratch)) return rc0;
  WhMaps maps;
  WhParams P;
  memset(&P, 0, sizeof(P));
  if (!encode5(&maps.x, (const bf16*)a->x, a->Cin, a->ldx, a->Wi, a->Hi, a->Di, a->B, HALO_W, HALO_H) ||
      !encode5(&maps.y, (const bf16*)a->y, a->Cout, a->ldy, a->Wo, a->Ho, a->Do, a->B, TILE_W, TILE_H)) {
    set_error("conv3d_wgrad(tcgen05 halo): cuTensorMapEncodeTiled failed");
    return MVD_ERR_CUDA;
  }
  wgrad_halo_plan(a, P);
  const int sets = P.cblocks * P.nblocks;
  const int ctas_per_set = P.ctas_per_set;
  P.slice_stride = wgrad_deterministic() ? (long long)a->Cout * a->Cin * 27 : 0;
  P.dw = scratch;
  const size_t smem = (size_t)kPlaneRing * PLANE_BYTES + (size_t)(kBrickRing + 2) * BRICK_BYTES + 1024;
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
  return tc_wgrad_finish(a, st);
}

}  // namespace mvd
