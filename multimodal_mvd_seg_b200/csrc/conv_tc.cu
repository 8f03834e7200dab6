// conv_tc.cu -- tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (fprop and dgrad of Conv3d, and through the
// adjoint formulation ConvTranspose3d fprop / dgrad).
//
// GEMM view:  D[m, n] = sum_{tap} sum_{k} A_tap[m, k] * W[tap][n][k]
//   m  : 128 voxels of the produced tensor = an 8 (w) x 16 (h) x 1 (d) brick of one sample (UMMA M = 128)
//   n  : produced channels, tile = UMMA N (<= 256, multiple of 32)
//   k  : gathered channels, KC = 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B) per pipeline stage
//   tap: a (tensor-map, dz, dy, dx) shift of the gathered tensor.  Strided convolutions are expressed on the s^3
//        parity sub-lattices of the strided side (each a plain strided 5-D tensor map), so every tap is an ordinary
//        box load; out-of-volume rows are zero-filled by TMA, which implements the conv padding.
// Pipeline (one persistent CTA per SM, 192 threads):
//   warp 0 lane 0 : TMA producer    -- per stage one 5-D box load of A (128 x KC) and one 2-D box load of W (N x KC)
//   warp 1 lane 0 : MMA issuer      -- KC/16 tcgen05.mma per stage into one of two TMEM accumulators; tcgen05.commit
//                                      releases the stage and, after the last stage, publishes the accumulator
//   warps 2..5    : epilogue        -- tcgen05.ld 32 columns at a time, + bias, bf16, 16-byte global stores
// smem full/empty mbarriers between producer and issuer, tmem full/empty mbarriers between issuer and epilogue.
#include <mutex>
#include <type_traits>
#include "conv_common.cuh"
#include "tc_common.cuh"
#include "tc_epilogue.cuh"

namespace mvd {

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  });
  return fn;
}

static bool tc_splitk_wanted(const mvd_conv3d_args* a, int pass);

namespace {

using namespace tc;

constexpr int kMaxTaps = 27;
constexpr int kMaxMaps = 8;
constexpr int kMaxClasses = 8;
constexpr int TILE_W = 8, TILE_H = 16;
constexpr int kThreads = 320;   // warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int kMaxBias = 1024;   // produced channels a bias vector is staged for

struct TcTap {
  int map;          // which A tensor map
  int dz, dy, dx;   // shift (in that map's lattice coordinates) relative to the tile origin
  int wrow;         // first row of this tap's [N][K] weight block in the 2-D weight map
};

// A launch covers up to 8 "classes": independent produced lattices that share the gathered tensor, the weights and
// the output strides (the parity classes of a strided dgrad).  Plain convolutions have one class.
struct TcClass {
  int tap_begin, ntaps;
  int Dt, Ht, Wt, tiles_w, tiles_h, tiles_d;
  int tile_begin;                // first (m) tile index of this class
  long long out_off;             // element offset of the class lattice origin inside `out`
};

struct alignas(64) TcMaps {
  CUtensorMap a[kMaxMaps];
  CUtensorMap b;
};

struct TcParams {
  int B;
  int nclasses, num_m_tiles, num_n_tiles;
  int n_tile;                    // UMMA N
  int kchunks;
  int stages;
  uint32_t idesc;
  uint32_t tmem_cols;
  bf16* out;                     // element (b,d,h,w,n) of a class at out + off + b*sb + d*sd + h*sh + w*sw + n
  long long sb, sd, sh, sw;
  const float* bias;             // per produced channel, may be null
  int accumulate;
  double* stats;                 // optional [B][Ntot][2] (sum, sumsq) of the bf16-rounded output
  int Ntot;
  int nbias;                     // channels the bias vector covers (scatter_c in scatter mode, else the produced channels)
  // scatter mode (transposed conv, kernel == stride): the N axis is (parity, channel); column n goes to channel
  // n % scatter_c of the voxel displaced by par_off[n / scatter_c]
  int scatter_c;
  long long par_off[8];
  int lane_own;                  // scatter mode without accumulate: each lane stores its own 64-byte row (tc_epilogue.cuh)
  // split-K (small produced lattices: 8^3 / 4^3 layers whose few tiles would leave most SMs idle while every tile
  // re-streams the whole weight set): a work unit is (tile, slice s of the tap x channel-chunk loop); the epilogue adds
  // its fp32 partial into `scratch` (same element offsets as `out`) with vector reductions, splitk_finish_kernel then
  // applies bias / rounding / accumulate
  // shape of the 128-voxel brick (M tile) in the produced lattice: (8 w, 16 h, 1 d, 1 sample) for the large layers; small
  // lattices (8^3, 4^3, 5x5x6 ...) use bricks that reach over planes and samples -- (8,8,2,1), (4,4,4,2) -- so that all
  // 128 MMA rows are real voxels and the layer's weights are streamed once per 128 voxels, not once per half-empty plane
  int bw, bh, bd, bb, tiles_b;
  int ksplit;
  float* scratch;              // [ksplit][out-shaped fp32]: slice s holds the partial of K slice s (plain stores, no atomics:
  long long slice_stride;      //  the finishing kernel adds the slices in a fixed order -> bit-reproducible results)
  TcClass cls[kMaxClasses];
  TcTap taps[kMaxTaps];
};

template <int KC>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ TcMaps maps,
                                                              const __grid_constant__ TcParams P) {
  constexpr int A_BYTES = 128 * KC * 2;
  constexpr uint64_t LAYOUT = (KC == 64) ? kLayoutSw128 : kLayoutSw64;
  constexpr uint32_t SBO = 8 * KC * 2;  // 8 rows of KC bf16
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar_full[8], bar_empty[8], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(16) uint8_t s_stage[8][2048];   // per epilogue warp: 32 rows x 64 B transpose buffer
  __shared__ __align__(16) float s_bias[kMaxBias];     // bias rounded to bf16 (zero when absent), indexed by channel

  // dynamic smem may only be 16-byte aligned by the runtime: align by hand (host adds 1 KB of slack)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int b_bytes = P.n_tile * KC * 2;
  const int stage_bytes = A_BYTES + b_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stages = P.stages;
  const int ksplit = P.ksplit;
  const int total_tiles = P.num_m_tiles * P.num_n_tiles * ksplit;   // work units: (tile, K slice)

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bar_tfull[a], 1);
      mbar_init(&bar_tempty[a], 8);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < P.nbias; i += blockDim.x) s_bias[i] = P.bias ? round_bf(__ldg(P.bias + i)) : 0.f;
  if (warp == 1) tmem_alloc(&s_tmem_base, P.tmem_cols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  // work unit -> (class, n0, brick origin b, d, h0, w0)
  const bool generic_brick = !(P.bw == 8 && P.bh == 16);
  auto decode_tile = [&](int unit, int& c, int& n0, int& b, int& d, int& h0, int& w0) {
    const int tile = unit / ksplit;
    const int nt = tile % P.num_n_tiles;
    int m = tile / P.num_n_tiles;
    n0 = nt * P.n_tile;
    c = 0;
    for (int i = 1; i < P.nclasses; ++i)
      if (m >= P.cls[i].tile_begin) c = i;
    m -= P.cls[c].tile_begin;
    const int tw = P.cls[c].tiles_w, th = P.cls[c].tiles_h, td = P.cls[c].tiles_d;
    w0 = (m % tw) * P.bw;
    m /= tw;
    h0 = (m % th) * P.bh;
    m /= th;
    d = (m % td) * P.bd;
    b = (m / td) * P.bb;
  };
  // row of the brick (0..127) -> voxel offsets inside it
  auto row_coords = [&](int rr, int& rw, int& rh, int& rd, int& rb) {
    if (!generic_brick) { rw = rr & 7; rh = rr >> 3; rd = 0; rb = 0; return; }
    rw = rr % P.bw; rr /= P.bw;
    rh = rr % P.bh; rr /= P.bh;
    rd = rr % P.bd;
    rb = rr / P.bd;
  };

  if (warp == 0) {
    if (elect_one_sync()) {
      // ================= TMA producer =================
      tma_prefetch_desc(&maps.b);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int c, n0, b, d, h0, w0;
        decode_tile(tile, c, n0, b, d, h0, w0);
        const int t0 = P.cls[c].tap_begin;
        const int kit = P.cls[c].ntaps * P.kchunks, sl = tile % ksplit;
        const int it0 = (int)((long long)kit * sl / ksplit), it1 = (int)((long long)kit * (sl + 1) / ksplit);
        int t = t0 + it0 / P.kchunks, kc = it0 % P.kchunks;
        for (int it = it0; it < it1; ++it) {
          const TcTap tap = P.taps[t];
          mbar_wait(&bar_empty[stage], phase ^ 1, 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          mbar_arrive_expect_tx(&bar_full[stage], (uint32_t)stage_bytes);
          tma_load_5d(&maps.a[tap.map], sa, &bar_full[stage], kc * KC, w0 + tap.dx, h0 + tap.dy, d + tap.dz, b);
          tma_load_2d(&maps.b, sa + A_BYTES, &bar_full[stage], kc * KC, tap.wrow + n0);
          if (++stage == stages) { stage = 0; phase ^= 1; }
          if (++kc == P.kchunks) { kc = 0; ++t; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      // ================= MMA issuer =================
      const int n_tile = P.n_tile, kchunks = P.kchunks, num_n_tiles = P.num_n_tiles, nclasses = P.nclasses;
      const uint32_t idesc = P.idesc;
      const uint32_t hi = (uint32_t)(make_smem_desc(0, 16, SBO, LAYOUT) >> 32);
      const uint32_t lo0 = (smem_u32(smem) >> 4) | (1u << 16);
      const uint32_t st_step = (uint32_t)stage_bytes >> 4;
      int stage = 0, acc = 0;
      uint32_t phase = 0, accphase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int c = 0;
        const int m = (tile / ksplit) / num_n_tiles;
        for (int i = 1; i < nclasses; ++i)
          if (m >= P.cls[i].tile_begin) c = i;
        const int kit = P.cls[c].ntaps * kchunks, sl = tile % ksplit;
        const int kiters = (int)((long long)kit * (sl + 1) / ksplit) - (int)((long long)kit * sl / ksplit);
        mbar_wait(&bar_tempty[acc], accphase ^ 1, 2);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * n_tile);
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(&bar_full[stage], phase, 3);
          tcgen05_fence_after();
          const uint32_t alo = lo0 + (uint32_t)stage * st_step;
          const uint64_t adesc = ((uint64_t)hi << 32) | (uint64_t)alo;
          const uint64_t bdesc = ((uint64_t)hi << 32) | (uint64_t)(alo + (uint32_t)(A_BYTES >> 4));
#pragma unroll
          for (int k = 0; k < KC / 16; ++k)  // advance 32 bytes (16 bf16) along K inside the swizzle atom
            umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it | k) ? 1u : 0u);
          umma_commit(&bar_empty[stage]);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&bar_tfull[acc]);
        acc ^= 1;
        if (acc == 0) accphase ^= 1;
      }
    }
  } else {
    // ================= epilogue (warps 2..5) =================
    // bias from shared memory, one packed conversion per column pair, per-lane InstanceNorm partials (tc_epilogue.cuh);
    // SC = number of 32-column chunks with fused statistics (host: n_tile == Ntot == 32 * SC, no scatter)
    auto run_epilogue = [&](auto sc_tag) {
      constexpr int SC = decltype(sc_tag)::value;
      const int q = warp & 3;  // the TMEM lane quarter this warp may access
      const int half = (warp - 2) >> 2;   // the two warps of a quarter take alternating 32-column chunks
      uint8_t* stage = s_stage[warp - 2];
      int acc = 0;
      uint32_t accphase = 0;
      LaneStats<SC> st;
      st.reset(-1);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int c, n0, b, d, h0, w0;
        decode_tile(tile, c, n0, b, d, h0, w0);
        if (SC > 0 && b != st.b) {
          if (st.b >= 0) st.flush(P.stats, P.Ntot, lane);
          st.reset(b);
        }
        const int Dt = P.cls[c].Dt, Ht = P.cls[c].Ht, Wt = P.cls[c].Wt;
        bf16* tile_base = P.out + P.cls[c].out_off + (long long)b * P.sb + (long long)d * P.sd;
        const long long sh = P.sh, sw = P.sw;
        // destination of channel 0 of the 32-column group starting at produced column n0 + cc, for row R of this warp
        auto col_offset = [&](int cc) -> long long {
          int n = n0 + cc;
          if (P.scatter_c) {        // (parity, channel): 32-column groups never straddle a parity (scatter_c % 32 == 0)
            const int par = n / P.scatter_c;
            return P.par_off[par] + (n - par * P.scatter_c);
          }
          return n;
        };
        // element offset of row rr of the brick relative to tile_base, or -1 when that voxel is outside the lattice
        auto row_off = [&](int rr) -> long long {
          int rw, rh, rd, rb;
          row_coords(rr, rw, rh, rd, rb);
          const int h = h0 + rh, w = w0 + rw;
          if (h >= Ht || w >= Wt || d + rd >= Dt || b + rb >= P.B) return -1;
          return (long long)rb * P.sb + (long long)rd * P.sd + (long long)h * sh + (long long)w * sw;
        };
        auto row_ptr_at = [&](int R, long long col_off) -> bf16* {
          const long long o = row_off(q * 32 + R);
          return o >= 0 ? tile_base + o + col_off : nullptr;
        };
        const bool accum = P.accumulate != 0 && ksplit == 1;
        bf16x8 old[4];
        bool has[4] = {false, false, false, false};
        if (accum && 32 * half < P.n_tile) {   // old values of the first chunk fly while the MMAs of this tile finish
          const long long co0 = col_offset(32 * half);
          prefetch_rows(lane, [&](int R) { return row_ptr_at(R, co0); }, old, has);
        }
        mbar_wait(&bar_tfull[acc], accphase, 4);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * P.n_tile);
        const int rr0 = q * 32 + lane;
        const long long own_off = row_off(rr0);
        const bool ok = own_off >= 0;
        for (int cc = 32 * half; cc < P.n_tile; cc += 64) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + (uint32_t)cc, v);
          tmem_ld_wait();
          if (ksplit > 1) {      // fp32 partial of this K slice -> its own scratch slice (8 x 16-byte stores per row chunk)
            if (ok) {
              const int sl = tile % ksplit;
              const int kit = P.cls[c].ntaps * P.kchunks;
              const bool empty = (int)((long long)kit * (sl + 1) / ksplit) == (int)((long long)kit * sl / ksplit);
              float* dst = P.scratch + (long long)sl * P.slice_stride + P.cls[c].out_off + (long long)b * P.sb +
                           (long long)d * P.sd + own_off + n0 + cc;
#pragma unroll
              for (int e = 0; e < 32; e += 4)
                *reinterpret_cast<uint4*>(dst + e) = empty ? make_uint4(0u, 0u, 0u, 0u) : make_uint4(v[e], v[e + 1], v[e + 2], v[e + 3]);
            }
            continue;
          }
          const long long col_off = col_offset(cc);
          int n = n0 + cc;          // bias index: channel within the parity in scatter mode
          if (P.scatter_c) n -= (n / P.scatter_c) * P.scatter_c;
          uint32_t w2[16];
          epilogue_chunk<SC>(v, P.bias ? s_bias + n : nullptr, st, cc, ok, w2);   // no bias: no shared-memory reads
          if (accum) {
            bf16x8 cur[4];
            bool chas[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { cur[i] = old[i]; chas[i] = has[i]; }
            if (cc + 64 < P.n_tile) {   // next chunk's old values
              const long long con = col_offset(cc + 64);
              prefetch_rows(lane, [&](int R) { return row_ptr_at(R, con); }, old, has);
            }
            store_rows_accumulate_packed(stage, lane, w2, [&](int R) { return row_ptr_at(R, col_off); }, cur, chas);
          } else if (P.lane_own) {
            // up-sampling scatter: a row is its own 64-byte segment of the output either way (stride-2 positions), and
            // with 256 columns per tile the eight epilogue warps, not the tensor pipe, set this kernel's pace
            if (ok) store_row_lane_own(tile_base + own_off + col_off, w2, false);
          } else {
            store_rows_coalesced_packed(stage, lane, w2, [&](int R) { return row_ptr_at(R, col_off); }, false);
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_tempty[acc]);
        acc ^= 1;
        if (acc == 0) accphase ^= 1;
      }
      if (SC > 0 && st.b >= 0) st.flush(P.stats, P.Ntot, lane);
    };
    if (!P.stats) run_epilogue(std::integral_constant<int, 0>{});
    else if (P.n_tile <= 32) run_epilogue(std::integral_constant<int, 1>{});
    else run_epilogue(std::integral_constant<int, 2>{});
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, P.tmem_cols);
}

// split-K tail: out[v][n] = bf16(scratch[v][n] + bias[n]) (+ out[v][n] when accumulating), voxels dense with pitch ld
__global__ void __launch_bounds__(256) splitk_finish_kernel(const float* __restrict__ scratch, long long slice_stride,
                                                            int nslices, bf16* __restrict__ out,
                                                            const float* __restrict__ bias, long long nvox, int N,
                                                            int ld, int accumulate) {
  const int CG = N >> 3;
  const long long total = nvox * CG;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long v = i / CG;
    const int cg = (int)(i - v * CG);
    const long long off = v * ld + cg * 8;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int sl = 0; sl < nslices; ++sl) {      // fixed order: reproducible bit for bit
      const float* p = scratch + (long long)sl * slice_stride + off;
      const float4 a0 = *reinterpret_cast<const float4*>(p), a1 = *reinterpret_cast<const float4*>(p + 4);
      f[0] += a0.x; f[1] += a0.y; f[2] += a0.z; f[3] += a0.w; f[4] += a1.x; f[5] += a1.y; f[6] += a1.z; f[7] += a1.w;
    }
    if (bias) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += round_bf(__ldg(bias + cg * 8 + j));
    }
    bf16x8 r = pack8(f);
    if (accumulate) {      // same two roundings as the in-kernel accumulate path: bf16(new), then bf16(new + old)
      float n[8], o[8];
      unpack8(r, n);
      unpack8(ldg16(out + off), o);
#pragma unroll
      for (int j = 0; j < 8; ++j) n[j] += o[j];
      r = pack8(n);
    }
    stg16(out + off, r);
  }
}

// -----------------------------------------------------------------------------------------------------------------
// host side
// -----------------------------------------------------------------------------------------------------------------
int pick_n_tile(int N) {
  if (N % 32) return 0;
  if (N <= 256) return N;
  for (int t = 256; t >= 32; t -= 32)
    if (N % t == 0) return t;
  return 0;
}

}  // namespace

bool tc_encode_act_map(CUtensorMap* m, const bf16* base, int C, int ld, const int dims[4] /*W,H,D,B extents*/,
                       const long long strides_el[4] /*W,H,D,B strides in elements*/, int kc, const int* brick) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)dims[0], (cuuint64_t)dims[1], (cuuint64_t)dims[2], (cuuint64_t)dims[3]};
  cuuint64_t gstr[4] = {(cuuint64_t)strides_el[0] * 2, (cuuint64_t)strides_el[1] * 2, (cuuint64_t)strides_el[2] * 2,
                        (cuuint64_t)strides_el[3] * 2};
  cuuint32_t box[5] = {(cuuint32_t)kc, TILE_W, TILE_H, 1, 1};
  if (brick) { box[1] = brick[0]; box[2] = brick[1]; box[3] = brick[2]; box[4] = brick[3]; }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)base, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   tc_l2_promotion(kc * 2, strides_el[0] * 2), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool tc_encode_w_map(CUtensorMap* m, const bf16* w, long long rows, int K, int n_tile, int kc) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)n_tile};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

namespace {

// shape coverage shared by fprop (K = Cin, N = Cout) and dgrad (K = Cout, N = Cin)
bool tc_shape_ok(int K, int N, int ldk, int ldn, const void* pk, const void* pn, const void* w, int taps) {
  if (K % 32 || pick_n_tile(N) == 0) return false;
  if (ldk % 8 || ldn % 8) return false;
  if (((uintptr_t)pk & 15) || ((uintptr_t)pn & 15) || ((uintptr_t)w & 15)) return false;
  if (taps > kMaxTaps) return false;
  return true;
}

// brick (M tile) shape for a produced lattice [B][D][H][W]: the default 8 x 16 plane brick, or -- for small lattices --
// a brick that reaches over planes / samples so that its 128 rows are real voxels
static void pick_brick(int B, int D, int H, int W, long long nvox_limit, int brick[4]) {
  brick[0] = 8; brick[1] = 16; brick[2] = 1; brick[3] = 1;
  if ((long long)B * D * H * W > nvox_limit || H > 8) return;
  const int bw = W > 4 ? 8 : 4;
  const int bh = H > 4 ? 8 : 4;
  int rest = 128 / (bw * bh);            // 2, 4 or 8 rows of (d, b) left
  int bd = 1;
  while (bd < rest && bd < D) bd <<= 1;
  brick[0] = bw; brick[1] = bh; brick[2] = bd; brick[3] = rest / bd;
}

// produced lattices this small go through the split-K form when the caller supplies a workspace (conv3d_workspace_bytes)
constexpr long long kSplitKMaxVoxels = 4096;
constexpr int kSplitKMaxSlices = 16;      // workspace is sized for this many K slices

struct SplitK {
  float* scratch = nullptr;     // caller's workspace
  size_t bytes = 0;             // its size: bounds the number of K slices
  long long nvox = 0;           // voxels of the produced lattice (all samples)
  int ld = 0, N = 0;            // pitch / channels of the produced tensor
};

int launch_tc(TcMaps& maps, TcParams& P, int kc, cudaStream_t st, const char* who, const SplitK* sk = nullptr) {
  if (P.nbias > kMaxBias) { set_error("%s: more than %d produced channels", who, kMaxBias); return MVD_ERR_UNSUPPORTED; }
  P.ksplit = 1;
  P.scratch = nullptr;
  P.slice_stride = 0;
  if (sk && sk->scratch && !P.scatter_c && !P.stats && sk->nvox <= kSplitKMaxVoxels) {
    const int tiles = P.num_m_tiles * P.num_n_tiles;
    int max_kit = 0;
    for (int c = 0; c < P.nclasses; ++c) max_kit = P.cls[c].ntaps * P.kchunks > max_kit ? P.cls[c].ntaps * P.kchunks : max_kit;
    const long long slice = sk->nvox * sk->ld;                 // floats per slice: the produced tensor's own offsets
    int S = num_sms() / (tiles > 0 ? tiles : 1);
    if (S > max_kit / 2) S = max_kit / 2;                      // at least two pipeline stages of work per unit
    const long long fit = (long long)(sk->bytes / sizeof(float)) / slice;
    if (S > fit) S = (int)fit;
    if (S > 1) {
      P.ksplit = S;
      P.scratch = sk->scratch;
      P.slice_stride = slice;
    }
  }
  float* const bias_keep = const_cast<float*>(P.bias);
  const int acc_keep = P.accumulate;
  if (P.ksplit > 1) P.stats = nullptr;        // statistics of a split layer: the caller's streaming pass
  const int a_bytes = 128 * kc * 2, b_bytes = P.n_tile * kc * 2;
  const int stage_bytes = a_bytes + b_bytes;
  int stages = (192 * 1024) / stage_bytes;   // + 21 KB of static shared memory (epilogue stages, bias)
  if (stages > 8) stages = 8;
  if (stages < 2) { set_error("%s: tile does not fit shared memory", who); return MVD_ERR_UNSUPPORTED; }
  P.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  uint32_t cols = 32;
  while ((int)cols < 2 * P.n_tile) cols <<= 1;
  P.tmem_cols = cols;
  P.idesc = make_idesc_bf16(128, P.n_tile, 0, 0);
  static bool attr_done[2] = {false, false};
  const int ki = (kc == 64) ? 0 : 1;
  if (!attr_done[ki]) {
    cudaError_t e = (kc == 64)
                        ? cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024)
                        : cudaFuncSetAttribute(conv_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(e));
      return MVD_ERR_CUDA;
    }
    attr_done[ki] = true;
  }
  const int total = P.num_m_tiles * P.num_n_tiles * P.ksplit;
  int grid = num_sms();
  if (grid > total) grid = total;
  if (kc == 64) conv_tc_kernel<64><<<grid, kThreads, smem, st>>>(maps, P);
  else conv_tc_kernel<32><<<grid, kThreads, smem, st>>>(maps, P);
  MVD_LAUNCH_CHECK(who);
  if (P.ksplit > 1) {
    const long long vecs = sk->nvox * (sk->N / 8);
    splitk_finish_kernel<<<grid_for(vecs, 256, num_sms() * 4), 256, 0, st>>>(sk->scratch, P.slice_stride, P.ksplit, P.out,
                                                                            bias_keep, sk->nvox, sk->N, sk->ld, acc_keep);
    MVD_LAUNCH_CHECK(who);
  }
  return MVD_OK;
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// appends a class covering the lattice (Dt, Ht, Wt); returns false when the lattice is empty
bool add_class(TcParams& P, int Dt, int Ht, int Wt, int tap_begin, int ntaps, long long out_off) {
  if (Dt <= 0 || Ht <= 0 || Wt <= 0) return false;
  TcClass& c = P.cls[P.nclasses++];
  c.tap_begin = tap_begin; c.ntaps = ntaps;
  c.Dt = Dt; c.Ht = Ht; c.Wt = Wt;
  c.tiles_w = cdiv(Wt, P.bw); c.tiles_h = cdiv(Ht, P.bh); c.tiles_d = cdiv(Dt, P.bd);
  P.tiles_b = cdiv(P.B, P.bb);
  c.tile_begin = P.num_m_tiles;
  c.out_off = out_off;
  P.num_m_tiles += P.tiles_b * c.tiles_d * c.tiles_h * c.tiles_w;
  return true;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// fprop: produced = y (conv output lattice), gathered = x.  Input coordinate o*s - p + t = s*(o + q) + r with
// r = (t - p) mod s, q = floor((t - p)/s): tap -> (parity map r, shift q).
// ---------------------------------------------------------------------------------------------------------------
// split-K is taken when the produced lattice is small, the reduction is long enough to slice, no statistics are fused
// into the epilogue (Cout 32 / 64) and the caller handed over the workspace mvd_conv3d_workspace_bytes asked for
static size_t splitk_bytes(const mvd_conv3d_args* a, int pass) {
  const long long nvox = pass == 0 ? (long long)a->B * a->Do * a->Ho * a->Wo : (long long)a->B * a->Di * a->Hi * a->Wi;
  const int ld = pass == 0 ? a->ldy : a->ldx, K = pass == 0 ? a->Cin : a->Cout, N = pass == 0 ? a->Cout : a->Cin;
  const int taps = a->kd * a->kh * a->kw;
  const bool scatter = pass == 1 && a->kd == a->sd && a->kh == a->sh && a->kw == a->sw && a->pd == 0 && a->ph == 0 && a->pw == 0;
  if (nvox > kSplitKMaxVoxels || (long long)taps * K < 1024 || scatter || N % 8) return 0;
  if (pass == 0 && a->stats && (a->Cout == 32 || a->Cout == 64)) return 0;
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("MVD_NO_SPLITK");
    enabled = (e && e[0] == '1') ? 0 : 1;
  }
  return enabled ? (size_t)kSplitKMaxSlices * nvox * ld * sizeof(float) : 0;
}
size_t tc_splitk_workspace_bytes(const mvd_conv3d_args* a, int pass) {
  if (pass == 0 ? !tc_fprop_supported(a) : !tc_dgrad_supported(a)) return 0;
  if (pass == 1 && tc_subpixel_dgrad_supported(a)) return 0;
  if (pass == 0 && tc_halo_s2_fprop_supported(a)) return 0;
  return splitk_bytes(a, pass);
}
static bool tc_splitk_wanted(const mvd_conv3d_args* a, int pass) {
  const size_t need = splitk_bytes(a, pass);
  return need > 0 && a->workspace && a->workspace_bytes >= need / kSplitKMaxSlices * 2 && (((uintptr_t)a->workspace) & 15) == 0;
}

bool tc_fprop_supported(const mvd_conv3d_args* a) {
  const int taps = a->kd * a->kh * a->kw;
  if (!tc_shape_ok(a->Cin, a->Cout, a->ldx, a->ldy, a->x, a->y, a->w, taps)) return false;
  if (a->sd * a->sh * a->sw > kMaxMaps) return false;
  return get_encode_tiled() != nullptr;
}

static inline void floordivmod(int v, int s, int& q, int& r) {
  q = (v >= 0) ? v / s : -((-v + s - 1) / s);
  r = v - q * s;
}

static bool is_k3s1p1(const mvd_conv3d_args* a) {
  return a->kd == 3 && a->kh == 3 && a->kw == 3 && a->sd == 1 && a->sh == 1 && a->sw == 1 && a->pd == 1 &&
         a->ph == 1 && a->pw == 1;
}

int tc_fprop(const mvd_conv3d_args* a, cudaStream_t st) {
  if (a->accumulate && ((is_k3s1p1(a) && tc_halo_enabled() && !tc_splitk_wanted(a, 0)) || tc_halo_s2_fprop_supported(a))) {
    set_error("conv3d_fprop(tcgen05): accumulate is built for the tap-by-tap kernel (up-convolution adjoints) only");
    return MVD_ERR_UNSUPPORTED;
  }
  if (is_k3s1p1(a) && tc_halo_enabled() && !tc_splitk_wanted(a, 0)) {
    int wrow[27];
    for (int i = 0; i < 27; ++i) wrow[i] = i * a->Cout;
    return tc_halo_conv((const bf16*)a->x, a->ldx, a->Cin, (bf16*)a->y, a->ldy, a->Cout, (const bf16*)a->w, wrow,
                        a->bias, 0, a->stats, a->B, a->Do, a->Ho, a->Wo, st, "conv3d_fprop(tcgen05 halo)");
  }
  if (tc_halo_s2_fprop_supported(a)) return tc_halo_s2_fprop(a, st);
  const int kc = (a->Cin % 64 == 0) ? 64 : 32;
  TcMaps maps;
  TcParams P;
  memset(&P, 0, sizeof(P));
  int brick[4];
  pick_brick(a->B, a->Do, a->Ho, a->Wo, (a->stats && (a->Cout == 32 || a->Cout == 64)) ? 0 : kSplitKMaxVoxels, brick);
  P.bw = brick[0]; P.bh = brick[1]; P.bd = brick[2]; P.bb = brick[3];
  const bf16* x = (const bf16*)a->x;
  const long long ld = a->ldx;
  bool ok = true;
  // parity sub-lattices of the input
  for (int rd = 0; rd < a->sd && ok; ++rd)
    for (int rh = 0; rh < a->sh && ok; ++rh)
      for (int rw = 0; rw < a->sw && ok; ++rw) {
        const int mi = (rd * a->sh + rh) * a->sw + rw;
        int dims[4] = {cdiv(a->Wi - rw, a->sw), cdiv(a->Hi - rh, a->sh), cdiv(a->Di - rd, a->sd), a->B};
        const bf16* base = x + ((long long)rd * a->Hi * a->Wi + (long long)rh * a->Wi + rw) * ld;
        // an empty lattice can never be hit by a valid tap: alias it to a 1-voxel lattice (all loads out of bounds)
        if (dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) { dims[0] = dims[1] = dims[2] = 1; base = x; }
        const long long strides[4] = {ld * a->sw, ld * a->Wi * a->sh, ld * a->Wi * a->Hi * a->sd,
                                      ld * a->Wi * a->Hi * a->Di};
        ok = tc_encode_act_map(&maps.a[mi], base, a->Cin, a->ldx, dims, strides, kc, brick);
      }
  const int taps = a->kd * a->kh * a->kw;
  P.n_tile = pick_n_tile(a->Cout);
  ok = ok && tc_encode_w_map(&maps.b, (const bf16*)a->w, (long long)taps * a->Cout, a->Cin, P.n_tile, kc);
  if (!ok) { set_error("conv3d_fprop(tcgen05): cuTensorMapEncodeTiled failed"); return MVD_ERR_CUDA; }
  int nt = 0;
  for (int td = 0; td < a->kd; ++td)
    for (int th = 0; th < a->kh; ++th)
      for (int tw = 0; tw < a->kw; ++tw) {
        int qd, rd, qh, rh, qw, rw;
        floordivmod(td - a->pd, a->sd, qd, rd);
        floordivmod(th - a->ph, a->sh, qh, rh);
        floordivmod(tw - a->pw, a->sw, qw, rw);
        TcTap& t = P.taps[nt++];
        t.map = (rd * a->sh + rh) * a->sw + rw;
        t.dz = qd; t.dy = qh; t.dx = qw;
        t.wrow = ((td * a->kh + th) * a->kw + tw) * a->Cout;
      }
  P.kchunks = a->Cin / kc;
  P.B = a->B;
  add_class(P, a->Do, a->Ho, a->Wo, 0, nt, 0);
  P.num_n_tiles = a->Cout / P.n_tile;
  P.out = (bf16*)a->y;
  P.sw = a->ldy; P.sh = (long long)a->ldy * a->Wo; P.sd = P.sh * a->Ho; P.sb = P.sd * a->Do;
  P.bias = a->bias;
  P.accumulate = a->accumulate;     // the adjoint of an up-convolution adds into the gradient a segmentation head left
  P.stats = a->stats; P.Ntot = a->Cout;
  P.nbias = a->Cout;
  if (P.stats && (P.n_tile != a->Cout || (a->Cout != 32 && a->Cout != 64))) {
    set_error("conv3d_fprop(tcgen05): fused InstanceNorm sums need Cout = 32 or 64");
    return MVD_ERR_UNSUPPORTED;
  }
  SplitK sk;
  sk.scratch = tc_splitk_wanted(a, 0) ? (float*)a->workspace : nullptr;
  sk.bytes = a->workspace_bytes;
  sk.nvox = (long long)a->B * a->Do * a->Ho * a->Wo; sk.ld = a->ldy; sk.N = a->Cout;
  return launch_tc(maps, P, kc, st, "conv3d_fprop(tcgen05)", &sk);
}

// ---------------------------------------------------------------------------------------------------------------
// dgrad: produced = x (conv input lattice), gathered = y.  x[i] += y[(i + p - t)/s] W[t] when divisible: the produced
// lattice splits into s^3 parity classes r = i mod s; within a class tap t contributes iff (r + p - t) mod s == 0, with
// shift (r + p - t)/s on the (dense) y lattice.  All classes go into ONE launch.  kernel == stride (the transposed
// convolutions of the decoder): every class has exactly one tap with shift 0, so the classes collapse into a single
// GEMM with N = s^3 * Cin whose epilogue scatters the (parity, channel) columns.
// ---------------------------------------------------------------------------------------------------------------
bool tc_dgrad_supported(const mvd_conv3d_args* a) {
  const int taps = a->kd * a->kh * a->kw;
  if (!tc_shape_ok(a->Cout, a->Cin, a->ldy, a->ldx, a->y, a->x, a->w, taps)) return false;
  // every parity class must own at least one tap per axis, otherwise the class is a pure fill (not built here)
  if (a->kd < a->sd || a->kh < a->sh || a->kw < a->sw) return false;
  if (a->sd * a->sh * a->sw > kMaxClasses) return false;
  return get_encode_tiled() != nullptr;
}

bool tc_dgrad_fuses_norm_bwd(const mvd_conv3d_args* a) {
  const mvd_norm_bwd_stats_args* nb = a->norm_bwd;
  if (!nb || !nb->y || !nb->stats || !nb->bstats || a->stats) return false;
  if (!tc_dgrad_supported(a) || tc_subpixel_dgrad_supported(a)) return false;
  if (!(is_k3s1p1(a) && tc_halo_enabled() && !tc_splitk_wanted(a, 1))) return false;
  // built into the depth-folded halo kernel's epilogue for 32 produced channels (the two full-resolution blocks)
  return a->Cin == 32 && tc_halo_fold_eligible(a->Cin, a->Cout, a->Di) && nb->ldy % 8 == 0 && (((uintptr_t)nb->y) & 15) == 0;
}

int tc_dgrad(const mvd_conv3d_args* a, cudaStream_t st) {
  if (tc_subpixel_dgrad_supported(a)) return tc_subpixel_dgrad(a, st);
  if (is_k3s1p1(a) && tc_halo_enabled() && !tc_splitk_wanted(a, 1)) {
    int wrow[27];   // produced voxel i gathers y[i + 1 - t]: halo offset o = 2 - t per axis, i.e. tap 26 - idx
    for (int i = 0; i < 27; ++i) wrow[i] = (26 - i) * a->Cin;
    // a->stats (conv_api passes it only for Cin = 32 / 64 without accumulate): per-(sample, channel) sums of the produced
    // gradient from the epilogue -- the bias gradient of the up-convolution that produced half of this tensor
    return tc_halo_conv((const bf16*)a->y, a->ldy, a->Cout, (bf16*)a->x, a->ldx, a->Cin, (const bf16*)a->w, wrow,
                        a->bias, a->accumulate, a->stats, a->B, a->Di, a->Hi, a->Wi, st, "conv3d_dgrad(tcgen05 halo)",
                        a->norm_bwd);
  }
  const int kc = (a->Cout % 64 == 0) ? 64 : 32;
  TcMaps maps;
  TcParams P;
  memset(&P, 0, sizeof(P));
  int brick[4];
  {   // bricks live on the parity sub-lattices of the produced tensor; the transposed-conv scatter keeps the default
    const bool scatter = a->kd == a->sd && a->kh == a->sh && a->kw == a->sw && a->pd == 0 && a->ph == 0 && a->pw == 0;
    pick_brick(a->B, cdiv(a->Di, a->sd), cdiv(a->Hi, a->sh), cdiv(a->Wi, a->sw), scatter ? 0 : kSplitKMaxVoxels, brick);
  }
  P.bw = brick[0]; P.bh = brick[1]; P.bd = brick[2]; P.bb = brick[3];
  const bf16* y = (const bf16*)a->y;
  const long long ldy = a->ldy;
  const int dims[4] = {a->Wo, a->Ho, a->Do, a->B};
  const long long strides[4] = {ldy, ldy * a->Wo, ldy * a->Wo * a->Ho, ldy * a->Wo * a->Ho * a->Do};
  if (!tc_encode_act_map(&maps.a[0], y, a->Cout, a->ldy, dims, strides, kc, brick)) {
    set_error("conv3d_dgrad(tcgen05): cuTensorMapEncodeTiled(A) failed");
    return MVD_ERR_CUDA;
  }
  for (int i = 1; i < kMaxMaps; ++i) maps.a[i] = maps.a[0];
  const int taps = a->kd * a->kh * a->kw;
  const long long ldx = a->ldx;
  P.B = a->B;
  P.kchunks = a->Cout / kc;
  P.out = (bf16*)a->x;
  P.sw = ldx * a->sw; P.sh = ldx * a->Wi * a->sh; P.sd = ldx * a->Wi * a->Hi * a->sd;
  P.sb = ldx * a->Wi * a->Hi * a->Di;
  P.bias = a->bias;
  P.accumulate = a->accumulate;
  P.nbias = a->Cin;

  const bool k_eq_s = (a->kd == a->sd && a->kh == a->sh && a->kw == a->sw && a->pd == 0 && a->ph == 0 && a->pw == 0 &&
                       a->Di == a->Do * a->sd && a->Hi == a->Ho * a->sh && a->Wi == a->Wo * a->sw);
  const int scatter_n = taps * a->Cin;
  if (k_eq_s && taps <= 8 && a->Cin % 32 == 0 && pick_n_tile(scatter_n) != 0) {
    // transposed-conv forward as one GEMM: tap == parity, weights [tap][Cin][Cout] are already the [taps*Cin][Cout] matrix
    P.n_tile = pick_n_tile(scatter_n);
    if (!tc_encode_w_map(&maps.b, (const bf16*)a->w, (long long)scatter_n, a->Cout, P.n_tile, kc)) {
      set_error("conv3d_dgrad(tcgen05): cuTensorMapEncodeTiled(W) failed");
      return MVD_ERR_CUDA;
    }
    P.taps[0] = TcTap{0, 0, 0, 0, 0};
    add_class(P, a->Do, a->Ho, a->Wo, 0, 1, 0);
    P.num_n_tiles = scatter_n / P.n_tile;
    P.scatter_c = a->Cin;
    {
      static int lo = -1;
      if (lo < 0) { const char* e = getenv("MVD_TC_LANE_OWN"); lo = (e && e[0] == '0') ? 0 : 1; }
      P.lane_own = (lo && !a->accumulate && a->Cin % 32 == 0) ? 1 : 0;
    }
    for (int td = 0; td < a->kd; ++td)
      for (int th = 0; th < a->kh; ++th)
        for (int tw = 0; tw < a->kw; ++tw)
          P.par_off[(td * a->kh + th) * a->kw + tw] = ((long long)td * a->Hi * a->Wi + (long long)th * a->Wi + tw) * ldx;
    return launch_tc(maps, P, kc, st, "conv_transpose3d_fprop(tcgen05)");
  }

  P.n_tile = pick_n_tile(a->Cin);
  if (!tc_encode_w_map(&maps.b, (const bf16*)a->w, (long long)taps * a->Cin, a->Cout, P.n_tile, kc)) {
    set_error("conv3d_dgrad(tcgen05): cuTensorMapEncodeTiled(W) failed");
    return MVD_ERR_CUDA;
  }
  P.num_n_tiles = a->Cin / P.n_tile;
  int nt = 0;
  for (int rd = 0; rd < a->sd; ++rd)
    for (int rh = 0; rh < a->sh; ++rh)
      for (int rw = 0; rw < a->sw; ++rw) {
        const int tap_begin = nt;
        for (int td = 0; td < a->kd; ++td) {
          if ((rd + a->pd - td) % a->sd) continue;
          for (int th = 0; th < a->kh; ++th) {
            if ((rh + a->ph - th) % a->sh) continue;
            for (int tw = 0; tw < a->kw; ++tw) {
              if ((rw + a->pw - tw) % a->sw) continue;
              TcTap& t = P.taps[nt++];
              t.map = 0;
              t.dz = (rd + a->pd - td) / a->sd;   // exact (C++ % and / agree in sign on exact division)
              t.dy = (rh + a->ph - th) / a->sh;
              t.dx = (rw + a->pw - tw) / a->sw;
              t.wrow = ((td * a->kh + th) * a->kw + tw) * a->Cin;
            }
          }
        }
        const int Dt = cdiv(a->Di - rd, a->sd), Ht = cdiv(a->Hi - rh, a->sh), Wt = cdiv(a->Wi - rw, a->sw);
        if (nt == tap_begin) {
          if (Dt > 0 && Ht > 0 && Wt > 0) { set_error("conv3d_dgrad(tcgen05): parity class without taps"); return MVD_ERR_UNSUPPORTED; }
          continue;
        }
        add_class(P, Dt, Ht, Wt, tap_begin, nt - tap_begin,
                  ((long long)rd * a->Hi * a->Wi + (long long)rh * a->Wi + rw) * ldx);
      }
  if (P.nclasses == 0) return MVD_OK;
  SplitK sk;
  sk.scratch = tc_splitk_wanted(a, 1) ? (float*)a->workspace : nullptr;
  sk.bytes = a->workspace_bytes;
  sk.nvox = (long long)a->B * a->Di * a->Hi * a->Wi; sk.ld = a->ldx; sk.N = a->Cin;
  return launch_tc(maps, P, kc, st, "conv3d_dgrad(tcgen05)", &sk);
}

}  // namespace mvd
