// tc_probe.cu -- hardware probe for UMMA shared-memory descriptor semantics (used by tests and by DESIGN.md's
// evidence): one CTA TMA-loads a [256 rows][C] bf16 tile whose values encode the row (or channel) index, multiplies
// it by an identity B tile with a caller-specified A descriptor (start offset, SBO, LBO, base offset, layout, major)
// and returns D, so the host can read off exactly which shared-memory rows / channels the tensor core fetched.
// Built by `make -C multimodal_mvd_seg_b200/csrc probe` into scripts/probes/libmvdseg_probe.so (not linked into libmvdseg.so).
#include "conv_common.cuh"
#include "tc_common.cuh"

namespace mvd {
namespace {
using namespace tc;

struct ProbeParams {
  int row_bytes;        // 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
  int start_off, sbo, lbo, base_off;
  int a_mn_major;       // 0: A K-major (rows = M), 1: A MN-major (rows = K)
  int kadv_bytes;       // second MMA: A start advanced by this many bytes
  float* out;           // [2][128][16]
};

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const __grid_constant__ CUtensorMap mapA,
                                                            const __grid_constant__ CUtensorMap mapB,
                                                            const __grid_constant__ ProbeParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar_full, bar_done;
  __shared__ uint32_t s_tmem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // 256 rows x row_bytes
  uint8_t* sB = smem + 256 * 128;     // 16 rows x 128 B (SWIZZLE_128B, K-major identity)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&s_tmem, 32);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s_tmem;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_full, (uint32_t)(256 * P.row_bytes + 16 * 128));
    tma_load_2d(&mapA, sA, &bar_full, 0, 0);
    tma_load_2d(&mapB, sB, &bar_full, 0, 0);
    mbar_wait(&bar_full, 0, 21);
    tcgen05_fence_after();
    const uint64_t layout = (P.row_bytes == 128) ? kLayoutSw128 : kLayoutSw64;
    const uint32_t a0 = smem_u32(sA) + (uint32_t)P.start_off;
    const uint64_t adesc0 = make_smem_desc(a0, (uint32_t)P.lbo, (uint32_t)P.sbo, layout, (uint32_t)P.base_off);
    const uint64_t adesc1 = make_smem_desc(a0 + (uint32_t)P.kadv_bytes, (uint32_t)P.lbo, (uint32_t)P.sbo, layout,
                                           (uint32_t)P.base_off);
    const uint64_t bdesc = make_smem_desc(smem_u32(sB), 16, 1024, kLayoutSw128);
    const uint32_t idesc = make_idesc_bf16(128, 16, P.a_mn_major, 0);
    umma_bf16(tmem, adesc0, bdesc, idesc, 0);
    umma_bf16(tmem + 16, adesc1, bdesc, idesc, 0);
    umma_commit(&bar_done);
  }
  __syncwarp();
  mbar_wait(&bar_done, 0, 22);
  tcgen05_fence_after();
  uint32_t v[32];
  tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  const int m = warp * 32 + lane;
  for (int j = 0; j < 16; ++j) {
    P.out[(0 * 128 + m) * 16 + j] = __uint_as_float(v[j]);
    P.out[(1 * 128 + m) * 16 + j] = __uint_as_float(v[16 + j]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

}  // namespace
}  // namespace mvd

using namespace mvd;

// src: bf16 [256][C] (C = row_bytes/2), ident: bf16 [16][64] with ident[n][k] = (n == k); out: float [2][128][16]
extern "C" int mvd_tc_probe(const void* src, const void* ident, int row_bytes, int start_off, int sbo, int lbo,
                            int base_off, int a_mn_major, int kadv_bytes, float* out, mvd_stream_t stream) {
  MVD_REQUIRE(src && ident && out && (row_bytes == 128 || row_bytes == 64), "tc_probe: bad arguments");
  PFN_encodeTiled enc = get_encode_tiled();
  MVD_REQUIRE(enc != nullptr, "tc_probe: no cuTensorMapEncodeTiled");
  CUtensorMap mA, mB;
  const int C = row_bytes / 2;
  {
    cuuint64_t gdim[2] = {(cuuint64_t)C, 256};
    cuuint64_t gstr[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)C, 256};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)src, gdim, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MVD_REQUIRE(r == CUDA_SUCCESS, "tc_probe: encode A failed (%d)", (int)r);
  }
  {
    cuuint64_t gdim[2] = {64, 16};
    cuuint64_t gstr[1] = {128};
    cuuint32_t box[2] = {64, 16};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)ident, gdim, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MVD_REQUIRE(r == CUDA_SUCCESS, "tc_probe: encode B failed (%d)", (int)r);
  }
  ProbeParams P{row_bytes, start_off, sbo, lbo, base_off, a_mn_major, kadv_bytes, out};
  const size_t smem = 256 * 128 + 16 * 128 + 1024;
  static bool attr = false;
  if (!attr) {
    MVD_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr = true;
  }
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mA, mB, P);
  MVD_LAUNCH_CHECK("tc_probe");
  return MVD_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// MMA issue-rate microbenchmark: one CTA issues `n_mma` tcgen05.mma (M = 128, N = n, K = 16, bf16) back to back from
// shared memory that is never refilled, round-robin over `n_acc` TMEM accumulators, optionally stepping the A start
// address by `a_step` bytes and using `a_sbo` as the 8-row-group pitch.  out[0] = cycles from first issue to the
// completion of the last MMA (tcgen05.commit -> mbarrier).
// ---------------------------------------------------------------------------------------------------------------
namespace mvd {
namespace {
struct MmaBenchParams {
  int n, n_mma, n_acc, a_sbo, a_step, a_steps_mod, row_bytes, mn_major, b_step, mode;
  long long* out;
};

__global__ void __launch_bounds__(128, 1) mma_bench_kernel(const __grid_constant__ MmaBenchParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t s_tmem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (160 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar_done, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&s_tmem, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = s_tmem;
  const uint64_t layout = (P.row_bytes == 128) ? kLayoutSw128 : kLayoutSw64;
  const uint32_t sa = smem_u32(smem);
  const uint32_t sb = smem_u32(smem) + 96 * 1024;
  const uint32_t idesc = make_idesc_bf16(128, P.n, P.mn_major, 0);
  const uint32_t lbo = P.mn_major ? 128 * P.row_bytes : 16;
  if (P.mode == 0) {
    // issue from a divergent single-lane region
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      for (int i = 0; i < P.n_mma; ++i) {
        const uint32_t ao = (uint32_t)((i & (P.a_steps_mod - 1)) * P.a_step);
        const uint32_t bo = (uint32_t)((i & 3) * P.b_step);
        const uint64_t adesc = make_smem_desc(sa + ao, lbo, (uint32_t)P.a_sbo, layout);
        const uint64_t bdesc = make_smem_desc(sb + bo, 16, 8 * 128, kLayoutSw128);
        umma_bf16(tmem + (uint32_t)((i & (P.n_acc - 1)) * P.n), adesc, bdesc, idesc, 1u);
      }
      umma_commit(&bar_done);
      mbar_wait(&bar_done, 0, 41);
      const long long t1 = clock64();
      if (blockIdx.x == 0) P.out[0] = t1 - t0;
    }
  } else if (P.mode == 2 && warp == 0) {
    // elect once, then the elected lane runs the whole issue loop (the CUTLASS / DeepGEMM pattern)
    const long long t0 = clock64();
    if (elect_one_sync()) {
      const uint64_t adesc0 = make_smem_desc(sa, lbo, (uint32_t)P.a_sbo, layout);
      const uint64_t bdesc0 = make_smem_desc(sb, 16, 8 * 128, kLayoutSw128);
      const uint32_t astep = (uint32_t)P.a_step >> 4, bstep = (uint32_t)P.b_step >> 4;
      const uint32_t amask = (uint32_t)P.a_steps_mod - 1, cmask = (uint32_t)P.n_acc - 1;
#pragma unroll 4
      for (int i = 0; i < P.n_mma; ++i) {
        umma_bf16(tmem + (((uint32_t)i & cmask) * (uint32_t)P.n), adesc0 + (uint64_t)(((uint32_t)i & amask) * astep),
                  bdesc0 + (uint64_t)(((uint32_t)i & 3u) * bstep), idesc, 1u);
      }
      umma_commit(&bar_done);
    }
    __syncwarp();
    mbar_wait(&bar_done, 0, 41);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) P.out[0] = t1 - t0;
  } else if (warp == 0) {
    // the whole warp runs the loop converged; one elected lane issues
    const long long t0 = clock64();
    for (int i = 0; i < P.n_mma; ++i) {
      const uint32_t ao = (uint32_t)((i & (P.a_steps_mod - 1)) * P.a_step);
      const uint32_t bo = (uint32_t)((i & 3) * P.b_step);
      const uint64_t adesc = make_smem_desc(sa + ao, lbo, (uint32_t)P.a_sbo, layout);
      const uint64_t bdesc = make_smem_desc(sb + bo, 16, 8 * 128, kLayoutSw128);
      if (elect_one_sync()) umma_bf16(tmem + (uint32_t)((i & (P.n_acc - 1)) * P.n), adesc, bdesc, idesc, 1u);
    }
    __syncwarp();
    if (elect_one_sync()) umma_commit(&bar_done);
    __syncwarp();
    mbar_wait(&bar_done, 0, 41);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) P.out[0] = t1 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
}  // namespace
}  // namespace mvd

extern "C" int mvd_tc_mma_bench(int n, int n_mma, int n_acc, int a_sbo, int a_step, int a_steps_mod, int row_bytes,
                                int mn_major, int b_step, int grid, int mode, long long* out_cycles, mvd_stream_t stream) {
  MVD_REQUIRE(out_cycles && n >= 16 && n <= 256 && n % 16 == 0 && n_mma > 0 && n_acc >= 1 && n_acc * n <= 512 &&
                  a_steps_mod >= 1 && grid >= 1, "tc_mma_bench: bad arguments");
  MmaBenchParams P{n, n_mma, n_acc, a_sbo, a_step, a_steps_mod, row_bytes, mn_major, b_step, mode, out_cycles};
  static bool attr = false;
  if (!attr) {
    MVD_CUDA(cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024));
    attr = true;
  }
  mma_bench_kernel<<<grid, 128, 161 * 1024 + 1024, (cudaStream_t)stream>>>(P);
  MVD_LAUNCH_CHECK("tc_mma_bench");
  return MVD_OK;
}
