// tc_probe.cu -- hardware probe for UMMA shared-memory descriptor semantics (used by tests and by DESIGN.md's
// evidence): one CTA TMA-loads a [256 rows][C] bf16 tile whose values encode the row (or channel) index, multiplies
// it by an identity B tile with a caller-specified A descriptor (start offset, SBO, LBO, base offset, layout, major)
// and returns D, so the host can read off exactly which shared-memory rows / channels the tensor core fetched.
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

