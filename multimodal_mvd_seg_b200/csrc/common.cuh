// common.cuh -- shared helpers for libmvdseg.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/mvdseg.h"

namespace mvd {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
void count_fallback();   // algo == auto fell through to the CUDA-core kernels (mvd_fallback_count)
int num_sms();

typedef __nv_bfloat16 bf16;

#define MVD_REQUIRE(cond, ...)                \
  do {                                        \
    if (!(cond)) {                            \
      mvd::set_error(__VA_ARGS__);            \
      return MVD_ERR_INVALID;                 \
    }                                         \
  } while (0)

#define MVD_LAUNCH_CHECK(name)                                                   \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      mvd::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
      return MVD_ERR_CUDA;                                                       \
    }                                                                            \
    mvd::count_launch();                                                         \
  } while (0)

#define MVD_CUDA(call)                                                           \
  do {                                                                           \
    cudaError_t e__ = (call);                                                    \
    if (e__ != cudaSuccess) {                                                    \
      mvd::set_error("%s failed: %s", #call, cudaGetErrorString(e__));           \
      return MVD_ERR_CUDA;                                                       \
    }                                                                            \
  } while (0)

__device__ __forceinline__ float bf2f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ bf16 f2bf(float v) { return __float2bfloat16_rn(v); }
// fp32 -> bf16 -> fp32 rounding.  The scalar conversion pair compiles to F2F.BF16.F32 on the quarter-rate XU pipe (ncu:
// 51 % XU utilisation in the InstanceNorm apply kernel); the packed conversion is F2FP.BF16.F32.PACK_AB on the ALU.
__device__ __forceinline__ float round_bf(float v) {
  __nv_bfloat162 p = __floats2bfloat162_rn(v, 0.f);
  return __uint_as_float(*reinterpret_cast<unsigned*>(&p) << 16);
}
__device__ __forceinline__ void round_bf2(float& a, float& b) {   // two values per conversion instruction
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  const unsigned u = *reinterpret_cast<unsigned*>(&p);
  a = __uint_as_float(u << 16);
  b = __uint_as_float(u & 0xffff0000u);
}

// 8 bf16 <-> 8 floats through one 16-byte vector
struct __align__(16) bf16x8 { __nv_bfloat162 v[4]; };

__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return p;
}

// 16-byte global accesses go through the built-in uint4: a load/store of the bf16x8 STRUCT is split by the compiler into
// four 32-bit LDG/STG (seen in SASS: every tcgen05 epilogue store was 4 requests of 16 partially written sectors; ncu
// l1tex throughput 78 % in the transposed-conv scatter) whenever the value is also touched element-wise.
__device__ __forceinline__ bf16x8 ldg16(const void* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  return *reinterpret_cast<const bf16x8*>(&u);
}
// Same load as volatile asm: ptxas otherwise sinks each of a batch of independent loads next to its use (one or two
// loads in flight per thread instead of the batch -- seen in SASS of the streaming kernels)
__device__ __forceinline__ bf16x8 ldg16_pinned(const void* p) {
  uint4 u;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p) : "memory");
  return *reinterpret_cast<const bf16x8*>(&u);
}
__device__ __forceinline__ void stg16(void* p, const bf16x8& v) {
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(&v);
}

// cp.async (LDGSTS) building blocks of the streaming kernels' private per-thread staging rings
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

inline int grid_for(long long work_items, int per_block, int max_blocks) {
  long long g = (work_items + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

}  // namespace mvd
