// conv_tc_wgrad.cu -- tcgen05 weight gradient:  dw[co][ci][tap] = sum_v y[v, co] * x[src(v, tap), ci].
//
// GEMM view (per "slot" = (tap, block of SW input channels)):
//     D[(slot, ci), co] += sum_{v in 128-voxel brick} Xslot[v, ci] * Y[v, co]
//   A = shifted x bricks, "MN-major" (channels contiguous, the reduction index v strides by one row), M = 128 rows =
//       2 slots of 64 channels (SWIZZLE_128B boxes) or 4 slots of 32 channels (SWIZZLE_64B boxes) side by side (LBO =
//       one box), so narrow layers still fill the 128 TMEM lanes;
//   B = the y brick [128 v][n_tile co], MN-major as well, n_tile <= 128 (1..2 boxes of 64 or 1..4 boxes of 32 channels);
//   K = 128 voxels per brick = 8 tcgen05.mma (K = 16) per slot group, start address advanced by 2 row groups per step.
// All accumulators of a CTA (G slot groups x n_tile columns <= 512) stay in TMEM over the CTA's whole voxel range
// (split-K across CTAs); a single epilogue adds them with 16-byte vector reductions (red.global.add.v4.f32) into a
// scratch laid out [tap][ci][co] (co = accumulator column, contiguous), which wgrad_finish_kernel then transposes into
// the torch layout dw[co][ci][tap].  (Scalar atomics straight into dw are stride-27 scattered: one L2 sector each;
// they took 60-90 % of the time of the low-resolution layers.)
// Work decomposition: unit = (set of G slot groups, N tile); grid = units x ksplit.
// Pipeline: warp 0 lane 0 TMA producer (B ring of 2 bricks + A ring of slot groups), warp 1 lane 0 MMA issuer,
// warps 2..5 epilogue.
#include "conv_common.cuh"
#include "tc_common.cuh"

namespace mvd {
namespace {

using namespace tc;

constexpr int kThreads = 192;
constexpr int kMaxTaps = 27;
constexpr int kMaxMaps = 8;
constexpr int TILE_W = 8, TILE_H = 16;

struct WgTap { int map, dz, dy, dx; };

struct alignas(64) WgMaps {
  CUtensorMap a[kMaxMaps];
  CUtensorMap b;
};

struct WgParams {
  int B, Dt, Ht, Wt, tiles_w, tiles_h, num_v_tiles;   // lattice of y (conv output) voxels
  int Cin, Cout, taps, cblocks;                        // cblocks = Cin / SW
  int total_slots, total_groups, G, num_sets;          // slot groups per set
  int n_tile, num_n_tiles, ksplit;
  int a_stages;
  uint32_t idesc, tmem_cols;
  float* dw;                                           // scratch [slices][taps][Cin][Cout]
  long long slice_stride;                              // 0: all splits reduce into slice 0 with red.global.add (atomics)
  WgTap tap[kMaxTaps];
};

template <int SW, int BW>
__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgMaps maps,
                                                               const __grid_constant__ WgParams P) {
  constexpr int SPG = 128 / SW;                 // slots per group
  constexpr int A_BOX = 128 * SW * 2;           // one slot brick
  constexpr int A_STAGE = SPG * A_BOX;          // 32 KB
  constexpr int B_BOX = 128 * BW * 2;
  constexpr uint64_t A_LAYOUT = (SW == 64) ? kLayoutSw128 : kLayoutSw64;
  constexpr uint64_t B_LAYOUT = (BW == 64) ? kLayoutSw128 : kLayoutSw64;
  constexpr uint32_t A_SBO = 8 * SW * 2, B_SBO = 8 * BW * 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar_afull[6], bar_aempty[6], bar_bfull[2], bar_bempty[2], bar_done;
  __shared__ uint32_t s_tmem_base;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_boxes = P.n_tile / BW;
  const int b_bytes = b_boxes * B_BOX;
  uint8_t* smem_b = smem;                         // 2 bricks of y
  uint8_t* smem_a = smem + 2 * b_bytes;           // a_stages slot groups
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int unit = blockIdx.x / P.ksplit, split = blockIdx.x % P.ksplit;
  const int set = unit / P.num_n_tiles, ntile = unit % P.num_n_tiles;
  const int n0 = ntile * P.n_tile;
  const int g_begin = set * P.G;
  int g_end = g_begin + P.G;
  if (g_end > P.total_groups) g_end = P.total_groups;
  const int ng = g_end - g_begin;
  const int vt_begin = (int)((long long)P.num_v_tiles * split / P.ksplit);
  const int vt_end = (int)((long long)P.num_v_tiles * (split + 1) / P.ksplit);

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.a_stages; ++s) { mbar_init(&bar_afull[s], 1); mbar_init(&bar_aempty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&bar_bfull[s], 1); mbar_init(&bar_bempty[s], 1); }
    mbar_init(&bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&s_tmem_base, P.tmem_cols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  auto decode_vt = [&](int vt, int& b, int& d, int& h0, int& w0) {
    w0 = (vt % P.tiles_w) * TILE_W;
    vt /= P.tiles_w;
    h0 = (vt % P.tiles_h) * TILE_H;
    vt /= P.tiles_h;
    d = vt % P.Dt;
    b = vt / P.Dt;
  };

  if (warp == 0) {
    if (ng > 0 && elect_one_sync()) {
      // ================= TMA producer =================
      int as = 0, bs = 0;
      uint32_t aphase = 0, bphase = 0;
      for (int vt = vt_begin; vt < vt_end; ++vt) {
        int b, d, h0, w0;
        decode_vt(vt, b, d, h0, w0);
        mbar_wait(&bar_bempty[bs], bphase ^ 1, 11);
        mbar_arrive_expect_tx(&bar_bfull[bs], (uint32_t)b_bytes);
        for (int j = 0; j < b_boxes; ++j)
          tma_load_5d(&maps.b, smem_b + bs * b_bytes + j * B_BOX, &bar_bfull[bs], n0 + j * BW, w0, h0, d, b);
        if (++bs == 2) { bs = 0; bphase ^= 1; }
        for (int g = g_begin; g < g_end; ++g) {
          mbar_wait(&bar_aempty[as], aphase ^ 1, 12);
          int nslot = P.total_slots - g * SPG;
          if (nslot > SPG) nslot = SPG;
          mbar_arrive_expect_tx(&bar_afull[as], (uint32_t)(nslot * A_BOX));
          for (int j = 0; j < nslot; ++j) {
            const int slot = g * SPG + j;
            const int tp = slot / P.cblocks, cb = slot - tp * P.cblocks;
            const WgTap t = P.tap[tp];
            tma_load_5d(&maps.a[t.map], smem_a + as * A_STAGE + j * A_BOX, &bar_afull[as], cb * SW, w0 + t.dx,
                        h0 + t.dy, d + t.dz, b);
          }
          if (++as == P.a_stages) { as = 0; aphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (ng > 0 && elect_one_sync()) {
      // ================= MMA issuer =================
      const int a_stages = P.a_stages, n_tile = P.n_tile;
      const uint32_t idesc = P.idesc;
      const uint32_t a_hi = (uint32_t)(make_smem_desc(0, A_BOX, A_SBO, A_LAYOUT) >> 32);
      const uint32_t b_hi = (uint32_t)(make_smem_desc(0, B_BOX, B_SBO, B_LAYOUT) >> 32);
      const uint32_t a_lo0 = (smem_u32(smem_a) >> 4) | ((uint32_t)(A_BOX >> 4) << 16);
      const uint32_t b_lo0 = (smem_u32(smem_b) >> 4) | ((uint32_t)(B_BOX >> 4) << 16);
      const uint32_t b_step = (uint32_t)b_bytes >> 4;
      int as = 0, bs = 0;
      uint32_t aphase = 0, bphase = 0;
      for (int vt = vt_begin; vt < vt_end; ++vt) {
        mbar_wait(&bar_bfull[bs], bphase, 13);
        tcgen05_fence_after();
        const uint64_t bdesc0 = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo0 + (uint32_t)bs * b_step);
        const uint32_t first = (vt > vt_begin) ? 1u : 0u;
        for (int g = 0; g < ng; ++g) {
          mbar_wait(&bar_afull[as], aphase, 14);
          tcgen05_fence_after();
          const uint64_t adesc0 = ((uint64_t)a_hi << 32) | (uint64_t)(a_lo0 + (uint32_t)as * (uint32_t)(A_STAGE >> 4));
          const uint32_t d_tmem = tmem_base + (uint32_t)(g * n_tile);
#pragma unroll
          for (int k = 0; k < 8; ++k)   // 16 voxels (two 8-row groups) per MMA
            umma_bf16(d_tmem, adesc0 + (uint64_t)(k * ((2 * A_SBO) >> 4)), bdesc0 + (uint64_t)(k * ((2 * B_SBO) >> 4)),
                      idesc, (k > 0) ? 1u : first);
          umma_commit(&bar_aempty[as]);
          if (++as == a_stages) { as = 0; aphase ^= 1; }
        }
        umma_commit(&bar_bempty[bs]);
        if (++bs == 2) { bs = 0; bphase ^= 1; }
      }
      umma_commit(&bar_done);
    }
  } else if (ng > 0 && vt_end > vt_begin) {
    // ================= epilogue (warps 2..5) =================
    const int q = warp & 3;
    mbar_wait(&bar_done, 0, 15);
    tcgen05_fence_after();
    const int m = q * 32 + lane;            // accumulator row = (slot in group, channel in slot)
    const int j = m / SW, cj = m - j * SW;
    for (int g = 0; g < ng; ++g) {
      const int slot = (g_begin + g) * SPG + j;
      const bool valid = slot < P.total_slots;
      const int tp = valid ? slot / P.cblocks : 0;
      const int ci = valid ? (slot - tp * P.cblocks) * SW + cj : 0;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * P.n_tile);
      for (int c = 0; c < P.n_tile; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + (uint32_t)c, v);
        tmem_ld_wait();
        if (valid) {   // scratch [tap][ci][co]: this row's 32 columns are 128 contiguous bytes
          float* dst = P.dw + (long long)split * P.slice_stride + ((long long)tp * P.Cin + ci) * P.Cout + n0 + c;
          if (P.slice_stride) {          // deterministic: plain stores into this split's own slice
#pragma unroll
            for (int e = 0; e < 32; e += 4) *reinterpret_cast<uint4*>(dst + e) = make_uint4(v[e], v[e + 1], v[e + 2], v[e + 3]);
          } else {                       // 8 vector reductions
#pragma unroll
            for (int e = 0; e < 32; e += 4) red_add_v4(dst + e, v[e], v[e + 1], v[e + 2], v[e + 3]);
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, P.tmem_cols);
}

// scratch [taps][Cin][Cout] -> dw [Cout][Cin][taps]; block = (ci, 64 output channels), transposed through smem
// deterministic mode, first level: the splits' slices are added in kSliceGroups interleaved groups (fixed order inside a
// group), one thread per element and group, so that the 16 MB of partials of a 148-way split are swept by the whole
// machine instead of by the finishing kernel's Cin x Cout/64 blocks; the groups land behind the slices
constexpr int kSliceGroups = 8;
__global__ void __launch_bounds__(256) wgrad_slice_reduce_kernel(const float* __restrict__ scratch, long long E,
                                                                 int nslices, float* __restrict__ groups) {
  const int g = blockIdx.y;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int sl = g; sl < nslices; sl += kSliceGroups) a += scratch[(long long)sl * E + e];
    groups[(long long)g * E + e] = a;
  }
}

// (nslices > 1: the deterministic two-stage reduction -- the slices are added here in a fixed order)
__global__ void __launch_bounds__(128) wgrad_finish_kernel(const float* __restrict__ scratch, float* __restrict__ dw,
                                                           int taps, int Cin, int Cout, int nslices,
                                                           long long slice_stride) {
  __shared__ float tile[27][65];
  const int ci = blockIdx.x, co0 = blockIdx.y * 64;
  const int nco = min(64, Cout - co0);
  for (int i = threadIdx.x; i < taps * 64; i += 128) {
    const int t = i >> 6, c = i & 63;
    if (c < nco) {
      const float* p = scratch + ((long long)t * Cin + ci) * Cout + co0 + c;
      float a = 0.f;
      for (int sl = 0; sl < nslices; ++sl) a += p[(long long)sl * slice_stride];
      tile[t][c] = a;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nco * taps; i += 128) {
    const int c = i / taps, t = i - c * taps;
    dw[((long long)(co0 + c) * Cin + ci) * taps + t] = tile[t][c];
  }
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline void floordivmod(int v, int s, int& q, int& r) {
  q = (v >= 0) ? v / s : -((-v + s - 1) / s);
  r = v - q * s;
}

int pick_wg_n_tile(int N) {
  if (N % 32) return 0;
  if (N <= 128) return N;
  for (int t = 128; t >= 32; t -= 32)
    if (N % t == 0) return t;
  return 0;
}

template <int SW, int BW>
int launch_wg(const WgMaps& maps, const WgParams& P, size_t smem, int grid, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<SW, BW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("conv3d_wgrad(tcgen05): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return MVD_ERR_CUDA;
    }
    attr_done = true;
  }
  wgrad_tc_kernel<SW, BW><<<grid, kThreads, smem, st>>>(maps, P);
  MVD_LAUNCH_CHECK("conv3d_wgrad(tcgen05)");
  return MVD_OK;
}

}  // namespace

bool tc_wgrad_supported(const mvd_conv3d_args* a) {
  if (a->Cin % 32 || a->Cout % 32) return false;
  if (pick_wg_n_tile(a->Cout) == 0) return false;
  if (a->ldx % 8 || a->ldy % 8) return false;
  if (((uintptr_t)a->x & 15) || ((uintptr_t)a->y & 15) || ((uintptr_t)a->dw & 3)) return false;
  if (a->kd * a->kh * a->kw > kMaxTaps || a->sd * a->sh * a->sw > kMaxMaps) return false;
  return get_encode_tiled() != nullptr;
}

// ---- reduction mode ---------------------------------------------------------------------------------------------
// default: every split of the voxel range reduces into ONE scratch with red.global.add.v4.f32 (order-dependent last bits).
// deterministic (mvd_set_deterministic(1) or MVD_DETERMINISTIC=1): every split writes its partial dw into its OWN slice
// of the workspace with plain stores and the slices are added in a fixed order (two levels: 8 interleaved groups, then
// the finishing kernel) -> dw is reproducible bit for bit.  Measured in the cfg-2 step on the B200: 9.72 ms vs 9.61 ms
// per step (+1.2 %: 15 more launches and up to 16 MB of partials per layer), which is why it is opt-in.
static int g_deterministic = -1;
bool wgrad_deterministic() {
  if (g_deterministic < 0) {
    const char* e = getenv("MVD_DETERMINISTIC");
    g_deterministic = (e && e[0] == '1') ? 1 : 0;
  }
  return g_deterministic == 1;
}
void set_wgrad_deterministic(int on) { g_deterministic = on ? 1 : 0; }

// number of splits of the voxel range each kernel uses for this layer (= slices in deterministic mode)
static int wgrad_tc_splits(const mvd_conv3d_args* a) {
  const int SW = (a->Cin % 64 == 0) ? 64 : 32;
  const int taps = a->kd * a->kh * a->kw;
  const int total_groups = cdiv(taps * (a->Cin / SW), 128 / SW);
  const int n_tile = pick_wg_n_tile(a->Cout);
  const int gmax = 512 / n_tile;
  const int units = cdiv(total_groups, gmax) * (a->Cout / n_tile);
  const int v_tiles = a->B * a->Do * cdiv(a->Ho, TILE_H) * cdiv(a->Wo, TILE_W);
  int ksplit = cdiv(num_sms(), units);
  if (ksplit > v_tiles) ksplit = v_tiles;
  return ksplit < 1 ? 1 : ksplit;
}
int tc_wgrad_halo_splits(const mvd_conv3d_args* a);   // conv_tc_wgrad_halo.cu

size_t tc_wgrad_workspace_bytes(const mvd_conv3d_args* a) {
  const size_t one = sizeof(float) * (size_t)a->Cout * a->Cin * a->kd * a->kh * a->kw;
  if (!wgrad_deterministic()) return one;
  const int splits = tc_wgrad_halo_supported(a) ? tc_wgrad_halo_splits(a) : wgrad_tc_splits(a);
  return one * (size_t)(splits + (splits > kSliceGroups ? kSliceGroups : 0));     // slices (+ the first-level group sums)
}
static int wgrad_slices(const mvd_conv3d_args* a) {
  return tc_wgrad_halo_supported(a) ? tc_wgrad_halo_splits(a) : wgrad_tc_splits(a);
}

int tc_wgrad_begin(const mvd_conv3d_args* a, cudaStream_t st, float** scratch) {
  const size_t need = tc_wgrad_workspace_bytes(a);
  if (!a->workspace || a->workspace_bytes < need || ((uintptr_t)a->workspace & 15)) {
    set_error("conv3d_wgrad(tcgen05): needs a 16-byte aligned workspace of %zu bytes (mvd_conv3d_workspace_bytes)", need);
    return MVD_ERR_INVALID;
  }
  // atomics accumulate into a zeroed scratch; in deterministic mode only the tap-by-tap kernel can leave holes (splits or
  // slot groups without work), the halo kernel overwrites every element of every slice
  if (!wgrad_deterministic())
    MVD_CUDA(cudaMemsetAsync(a->workspace, 0, need, st));
  else if (!tc_wgrad_halo_supported(a))
    MVD_CUDA(cudaMemsetAsync(a->workspace, 0, sizeof(float) * (size_t)a->Cout * a->Cin * a->kd * a->kh * a->kw * wgrad_slices(a), st));
  *scratch = (float*)a->workspace;
  return MVD_OK;
}

int tc_wgrad_finish(const mvd_conv3d_args* a, cudaStream_t st) {
  const int taps = a->kd * a->kh * a->kw;
  dim3 grid(a->Cin, (a->Cout + 63) / 64);
  const long long one = (long long)a->Cout * a->Cin * taps;
  int nslices = wgrad_deterministic() ? wgrad_slices(a) : 1;
  const float* src = (const float*)a->workspace;
  if (nslices > kSliceGroups) {      // first level of the fixed-order reduction
    float* groups = (float*)a->workspace + (long long)nslices * one;
    wgrad_slice_reduce_kernel<<<dim3((unsigned)grid_for(one, 256, num_sms() * 2), kSliceGroups), 256, 0, st>>>(src, one, nslices,
                                                                                                          groups);
    MVD_LAUNCH_CHECK("conv3d_wgrad(slice reduce)");
    src = groups;
    nslices = kSliceGroups;
  }
  wgrad_finish_kernel<<<grid, 128, 0, st>>>(src, a->dw, taps, a->Cin, a->Cout, nslices, one);
  MVD_LAUNCH_CHECK("conv3d_wgrad(finish)");
  if (a->dbias)
    return mvd_channel_sum(a->y, a->ldy, (long long)a->B * a->Do * a->Ho * a->Wo, a->Cout, a->dbias, (mvd_stream_t)st);
  return MVD_OK;
}

int tc_wgrad(const mvd_conv3d_args* a, cudaStream_t st) {
  if (tc_wgrad_halo_supported(a)) return tc_wgrad_halo(a, st);
  const int SW = (a->Cin % 64 == 0) ? 64 : 32;
  const int BW = (a->Cout % 64 == 0) ? 64 : 32;
  const int taps = a->kd * a->kh * a->kw;
  float* scratch = nullptr;
  if (int rc0 = tc_wgrad_begin(a, st, &scratch)) return rc0;
  WgMaps maps;
  WgParams P;
  memset(&P, 0, sizeof(P));
  const bf16* x = (const bf16*)a->x;
  const long long ld = a->ldx;
  bool ok = true;
  for (int rd = 0; rd < a->sd && ok; ++rd)
    for (int rh = 0; rh < a->sh && ok; ++rh)
      for (int rw = 0; rw < a->sw && ok; ++rw) {
        const int mi = (rd * a->sh + rh) * a->sw + rw;
        int dims[4] = {cdiv(a->Wi - rw, a->sw), cdiv(a->Hi - rh, a->sh), cdiv(a->Di - rd, a->sd), a->B};
        const bf16* base = x + ((long long)rd * a->Hi * a->Wi + (long long)rh * a->Wi + rw) * ld;
        if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) { dims[0] = dims[1] = dims[2] = 1; base = x; }
        const long long strides[4] = {ld * a->sw, ld * a->Wi * a->sh, ld * a->Wi * a->Hi * a->sd,
                                      ld * a->Wi * a->Hi * a->Di};
        ok = tc_encode_act_map(&maps.a[mi], base, a->Cin, a->ldx, dims, strides, SW);
      }
  for (int i = a->sd * a->sh * a->sw; i < kMaxMaps; ++i) maps.a[i] = maps.a[0];
  {
    const long long ldy = a->ldy;
    const int dims[4] = {a->Wo, a->Ho, a->Do, a->B};
    const long long strides[4] = {ldy, ldy * a->Wo, ldy * a->Wo * a->Ho, ldy * a->Wo * a->Ho * a->Do};
    ok = ok && tc_encode_act_map(&maps.b, (const bf16*)a->y, a->Cout, a->ldy, dims, strides, BW);
  }
  if (!ok) { set_error("conv3d_wgrad(tcgen05): cuTensorMapEncodeTiled failed"); return MVD_ERR_CUDA; }
  int nt = 0;
  for (int td = 0; td < a->kd; ++td)
    for (int th = 0; th < a->kh; ++th)
      for (int tw = 0; tw < a->kw; ++tw) {
        int qd, rd, qh, rh, qw, rw;
        floordivmod(td - a->pd, a->sd, qd, rd);
        floordivmod(th - a->ph, a->sh, qh, rh);
        floordivmod(tw - a->pw, a->sw, qw, rw);
        WgTap& t = P.tap[nt++];
        t.map = (rd * a->sh + rh) * a->sw + rw;
        t.dz = qd; t.dy = qh; t.dx = qw;
      }
  P.B = a->B; P.Dt = a->Do; P.Ht = a->Ho; P.Wt = a->Wo;
  P.tiles_w = cdiv(P.Wt, TILE_W); P.tiles_h = cdiv(P.Ht, TILE_H);
  P.num_v_tiles = P.B * P.Dt * P.tiles_h * P.tiles_w;
  P.Cin = a->Cin; P.Cout = a->Cout; P.taps = taps; P.cblocks = a->Cin / SW;
  const int spg = 128 / SW;
  P.total_slots = taps * P.cblocks;
  P.total_groups = cdiv(P.total_slots, spg);
  P.n_tile = pick_wg_n_tile(a->Cout);
  P.num_n_tiles = a->Cout / P.n_tile;
  const int gmax = 512 / P.n_tile;
  P.num_sets = cdiv(P.total_groups, gmax);
  P.G = cdiv(P.total_groups, P.num_sets);
  uint32_t cols = 32;
  while ((int)cols < P.G * P.n_tile) cols <<= 1;
  P.tmem_cols = cols;
  P.idesc = make_idesc_bf16(128, P.n_tile, 1, 1);
  const int units = P.num_sets * P.num_n_tiles;
  int ksplit = cdiv(num_sms(), units);
  if (ksplit > P.num_v_tiles) ksplit = P.num_v_tiles;
  if (ksplit < 1) ksplit = 1;
  P.ksplit = ksplit;
  P.dw = scratch;
  P.slice_stride = wgrad_deterministic() ? (long long)a->Cout * a->Cin * taps : 0;
  const int b_bytes = P.n_tile * 128 * 2;
  const int a_stage = 32 * 1024;
  int a_stages = (196 * 1024 - 2 * b_bytes) / a_stage;
  if (a_stages > 6) a_stages = 6;
  P.a_stages = a_stages;
  const size_t smem = (size_t)2 * b_bytes + (size_t)a_stages * a_stage + 1024;
  const int grid = units * ksplit;
  int rc;
  if (SW == 64 && BW == 64) rc = launch_wg<64, 64>(maps, P, smem, grid, st);
  else if (SW == 64) rc = launch_wg<64, 32>(maps, P, smem, grid, st);
  else if (BW == 64) rc = launch_wg<32, 64>(maps, P, smem, grid, st);
  else rc = launch_wg<32, 32>(maps, P, smem, grid, st);
  if (rc) return rc;
  return tc_wgrad_finish(a, st);
}

}  // namespace mvd
