// tc_common.cuh -- raw sm_100a building blocks: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (TMEM alloc, MMA,
// commit, ld) and the UMMA shared-memory / instruction descriptors.  Inline PTX only; no CUTLASS.
#pragma once
#include <cuda.h>
#include <stdint.h>
#include "common.cuh"

namespace mvd {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (CUDA error), never as a hung GPU.
static __device__ int g_tc_timeout_code;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s
      g_tc_timeout_code = code;
      __trap();
    }
  }
}

// one lane of a fully converged warp (the pattern the compiler turns into straight-line uniform-datapath code:
// UTCHMMA / UTMALDG issued from a divergent `if (lane == 0)` region are wrapped in a per-active-lane loop instead)
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "elect.sync _|P1, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- TMA -----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, void* dst, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, void* dst, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ---- tcgen05 -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole warp; writes the TMEM base address into *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (taddr.lane + i), columns c..c+31
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 16-byte vector reduction into global memory (REDG.E.ADD.F32x4): one L2 transaction for four fp32 adds
__device__ __forceinline__ void red_add_v4(float* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(__uint_as_float(a)), "f"(__uint_as_float(b)),
               "f"(__uint_as_float(c)), "f"(__uint_as_float(d))
               : "memory");
}

// ---- descriptors ---------------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (64 bit): start address >> 4 in [0,14), leading byte offset >> 4 in [16,30),
// stride byte offset >> 4 in [32,46), version = 1 in [46,48), base offset in [49,52), layout type in [61,64)
// (0 none, 2 = 128B swizzle, 4 = 64B swizzle, 6 = 32B swizzle).
enum : uint64_t { kLayoutNone = 0, kLayoutSw128 = 2, kLayoutSw64 = 4, kLayoutSw32 = 6 };
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint64_t layout, uint32_t base_offset = 0) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)(base_offset & 7) << 49) |
         (layout << 61);
}
// Instruction descriptor for kind::f16: D fp32 (c_format 1 at [4,6)), A/B bf16 (format 1 at [7,10) / [10,13)),
// a_major at 15, b_major at 16 (0 = K-major, 1 = MN-major), N >> 3 at [17,23), M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

}  // namespace tc

// host: driver entry point for tensor-map encoding, fetched at run time (no link-time libcuda dependency, so the
// library loads on machines without a driver)
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();
// L2 fetch granularity of an activation map.  Dense rows: 256 B.  When the voxel pitch is wider than the box row (a
// channel slice of the decoder's concat buffer, or a stride-2 parity lattice) a promoted fetch drags the unused
// neighbour through DRAM as well -- ncu on the stride-2 forward kernel reading the 32-channel skip out of the 64-channel
// buffer: 555 MB read for 268 MB of input -- so promotion stops at the row size.
inline CUtensorMapL2promotion tc_l2_promotion(long long row_bytes, long long pitch_bytes) {
  if (pitch_bytes <= row_bytes || row_bytes >= 256) return CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  if (row_bytes >= 128) return CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  if (row_bytes >= 64) return CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
  return CU_TENSOR_MAP_L2_PROMOTION_NONE;
}
// 5-D tensor map (C, W, H, D, B) over a pitched NDHWC bf16 lattice with box (box_c, 8, 16, 1, 1); swizzle 128B when
// box_c == 64, 64B when box_c == 32.  dims / strides are the W,H,D,B extents and element strides of the lattice.
bool tc_encode_act_map(CUtensorMap* m, const bf16* base, int C, int ld, const int dims[4], const long long strides_el[4],
                       int box_c, const int* brick = nullptr /* {w, h, d, b} box, default {8, 16, 1, 1} */);
// 2-D map over packed weights [rows][K] with box (kc, n_tile)
bool tc_encode_w_map(CUtensorMap* m, const bf16* w, long long rows, int K, int n_tile, int kc);
// halo-plane kernel for 3x3x3 / stride 1 / pad 1 (conv_tc_halo.cu): dst[v][n] = sum_{o in {0,1,2}^3} sum_k
// src[v + o - 1][k] * W[wrow[o] + n][k] (+ bias[n]); src/dst are pitched NDHWC lattices of the same extent B,D,H,W.
bool tc_halo_enabled();
bool tc_halo_fold_eligible(int N, int K, int D);   // would tc_halo_conv run the depth-folded kernel for this layer?
int tc_halo_conv(const bf16* src, int lds, int K, bf16* dst, int ldd, int N, const bf16* w, const int wrow[27],
                 const float* bias, int accumulate, double* stats, int B, int D, int H, int W, cudaStream_t st,
                 const char* who, const mvd_norm_bwd_stats_args* norm_bwd = nullptr);

}  // namespace mvd
