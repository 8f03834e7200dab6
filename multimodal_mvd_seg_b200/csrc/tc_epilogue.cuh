// tc_epilogue.cuh -- coalesced bf16 stores for the tcgen05 epilogues.
// tcgen05.ld (32x32b) hands every lane one accumulator ROW (32 consecutive output channels of one voxel).  Storing
// those 64 bytes straight from the owning lane makes each warp store instruction touch 32 different rows with 16 B
// each (half-written sectors: ncu showed 7x write amplification between L1 and L2).  Here the warp transposes the
// 32 x 64 B block through a private 2 KB shared-memory stage (XOR-swizzled, conflict-free in both directions) so that
// every store instruction writes 8 complete rows = the 8 w-adjacent voxels of one h line (512 contiguous bytes when the
// tensor is dense).
#pragma once
#include "common.cuh"

namespace mvd {
namespace tc {

// row_ptr(R) -> destination of channel 0 of this 32-channel group for row R (0..31) of the warp's block, or nullptr when
// that voxel is outside the tensor.  w[i] = bf16x2 of columns (2i, 2i+1) of the lane's own row (bias already added).
template <typename RowPtrFn>
__device__ __forceinline__ void store_rows_coalesced_packed(uint8_t* stage, int lane, const uint32_t* w, RowPtrFn row_ptr,
                                                            bool accumulate) {
  const int sw_own = (lane >> 1) & 3;
#pragma unroll
  for (int g = 0; g < 4; ++g)
    *reinterpret_cast<uint4*>(stage + lane * 64 + ((g ^ sw_own) << 4)) =
        make_uint4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
  __syncwarp();
  const int c = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int R = 8 * i + (lane >> 2);
    bf16x8 v = *reinterpret_cast<const bf16x8*>(stage + R * 64 + ((c ^ ((R >> 1) & 3)) << 4));
    bf16* dst = row_ptr(R);
    if (dst) {
      bf16* d8 = dst + c * 8;
      if (accumulate) {
        float a[8], o[8];
        unpack8(v, a);
        unpack8(ldg16(d8), o);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += o[j];
        v = pack8(a);
      }
      stg16(d8, v);
    }
  }
  __syncwarp();
}

// Accumulating variant with the OLD destination values prefetched by the caller (old[i] belongs to row 8*i + lane/4,
// 16-byte column group lane&3 -- the same mapping the stores use; has[i] says whether that row exists).  Lets the
// global loads fly while the warp still waits for its accumulator / converts, instead of a load-add-store chain with
// the full memory latency exposed four times per chunk.
template <typename RowPtrFn>
__device__ __forceinline__ void prefetch_rows(int lane, RowPtrFn row_ptr, bf16x8* old, bool* has) {
  const int c = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bf16* src = row_ptr(8 * i + (lane >> 2));
    has[i] = src != nullptr;
    if (has[i]) old[i] = ldg16(src + c * 8);
  }
}

template <typename RowPtrFn>
__device__ __forceinline__ void store_rows_accumulate_packed(uint8_t* stage, int lane, const uint32_t* w, RowPtrFn row_ptr,
                                                             const bf16x8* old, const bool* has) {
  const int sw_own = (lane >> 1) & 3;
#pragma unroll
  for (int g = 0; g < 4; ++g)
    *reinterpret_cast<uint4*>(stage + lane * 64 + ((g ^ sw_own) << 4)) =
        make_uint4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
  __syncwarp();
  const int c = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int R = 8 * i + (lane >> 2);
    bf16x8 v = *reinterpret_cast<const bf16x8*>(stage + R * 64 + ((c ^ ((R >> 1) & 3)) << 4));
    if (has[i]) {
      float a[8], o[8];
      unpack8(v, a);
      unpack8(old[i], o);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += o[j];
      stg16(row_ptr(R) + c * 8, pack8(a));
    }
  }
  __syncwarp();
}

// Lane-own stores: every lane writes the 64 bytes of ITS accumulator row (32 channels of one voxel) with four 16-byte
// stores -- no shared-memory transpose.  Each store instruction then touches 32 rows with half a sector each (L2 merges
// them), which costs L1->L2 requests but no shared-memory bandwidth: the tensor pipe's operand reads take almost all of it
// for narrow layers (N = 32: 7 KB per 56-cycle MMA = 125 B/clk of the 128 B/clk), and an epilogue that goes through shared
// memory is then throttled to the leftover (measured, 32->32 at 128^3: forward 0.229 -> 0.212 ms with lane-own stores;
// profiles/r2_epilogue_shared_memory.txt).  w[i] = bf16x2 of columns (2i, 2i+1); returns the values written (after the
// optional accumulate) in w.  (Not for the stride-2 sub-pixel data gradient: there the accumulate path's old values arrive
// better through the transposed, prefetched form -- 0.239 vs 0.206 ms for 32->64 at 64^3 when tried.)
__device__ __forceinline__ void store_row_lane_own(bf16* dst, uint32_t* w, bool accumulate) {
  if (accumulate) {
    uint4 old[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) old[g] = *reinterpret_cast<const uint4*>(dst + 8 * g);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint32_t o[4] = {old[g].x, old[g].y, old[g].z, old[g].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float lo = __uint_as_float(w[4 * g + k] << 16) + __uint_as_float(o[k] << 16);
        const float hi = __uint_as_float(w[4 * g + k] & 0xffff0000u) + __uint_as_float(o[k] & 0xffff0000u);
        __nv_bfloat162 pk = __floats2bfloat162_rn(lo, hi);
        w[4 * g + k] = *reinterpret_cast<uint32_t*>(&pk);
      }
    }
  }
#pragma unroll
  for (int g = 0; g < 4; ++g)
    *reinterpret_cast<uint4*>(dst + 8 * g) = make_uint4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
}

// ---- InstanceNorm statistics fused into the conv epilogue ---------------------------------------------------------
// Every lane holds one accumulator row (32 channels).  Column sums over the warp's 32 rows are formed with a
// recursive-halving exchange (16+8+4+2+1 = 31 shuffles per quantity instead of 5 x 32): afterwards lane l owns channel l.
// (used once per flush by LaneStats and per tile by the stem kernel, whose register budget is too small for LaneStats)
__device__ __forceinline__ float warp_column_sum32(float* v, int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// ---- per-lane statistics (no shuffles in the tile loop) ----------------------------------------------------------
// Every lane keeps fp32 partial sums of ITS accumulator row for all columns across the CTA's work items; the cross-lane
// reduction + fp64 atomics run once per (sample, CTA).  Valid while n_tile == C == 32 * SC and the N-tile origin is 0.
template <int SC>   // number of 32-column chunks with fused statistics (0, 1, 2): sizes the register arrays
struct LaneStats {
  float s0[SC > 0 ? 32 : 1], q0[SC > 0 ? 32 : 1], s1[SC > 1 ? 32 : 1], q1[SC > 1 ? 32 : 1];
  int b;
  __device__ __forceinline__ void reset(int b_) {
    b = b_;
    if (SC > 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) s0[i] = q0[i] = 0.f;
    }
    if (SC > 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) s1[i] = q1[i] = 0.f;
    }
  }
  __device__ __forceinline__ void flush(double* stats, int C, int lane) {
    if (SC > 0) {
      const float a = warp_column_sum32(s0, lane), c = warp_column_sum32(q0, lane);
      double* d = stats + ((long long)b * C + lane) * 2;
      atomicAdd(d, (double)a);
      atomicAdd(d + 1, (double)c);
      if (SC > 1) {
        const float a1 = warp_column_sum32(s1, lane), c1 = warp_column_sum32(q1, lane);
        atomicAdd(d + 64, (double)a1);
        atomicAdd(d + 65, (double)c1);
      }
    }
  }
};


// One 32-column accumulator chunk of a lane's row: + bias (pre-rounded, from shared memory), ONE packed fp32->bf16
// conversion per column pair (F2FP on the ALU pipe) that serves both the store (w2[16] = bf16x2 words) and the
// InstanceNorm sums of the rounded values (per-lane fp32 partials, see LaneStats).
template <int SC>
__device__ __forceinline__ void epilogue_chunk(const uint32_t* v, const float* bias32, LaneStats<SC>& st, int c,
                                               bool ok, uint32_t* w2, const float2* bias_reg = nullptr) {
  if (bias_reg) {     // bias of a single-chunk layer kept in registers by the caller (no shared-memory reads per step)
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      __nv_bfloat162 pk = __floats2bfloat162_rn(__uint_as_float(v[j]) + bias_reg[j >> 1].x,
                                                __uint_as_float(v[j + 1]) + bias_reg[j >> 1].y);
      w2[j >> 1] = *reinterpret_cast<uint32_t*>(&pk);
    }
  } else if (bias32) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float2 bb = *reinterpret_cast<const float2*>(bias32 + j);
      __nv_bfloat162 pk = __floats2bfloat162_rn(__uint_as_float(v[j]) + bb.x, __uint_as_float(v[j + 1]) + bb.y);
      w2[j >> 1] = *reinterpret_cast<uint32_t*>(&pk);
    }
  } else {     // no bias (data gradients): no shared-memory reads next to the tensor pipe's operand traffic
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      __nv_bfloat162 pk = __floats2bfloat162_rn(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
      w2[j >> 1] = *reinterpret_cast<uint32_t*>(&pk);
    }
  }
  if (SC > 0 && ok) {
    if (SC == 1 || c == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float lo = __uint_as_float(w2[j >> 1] << 16), hi = __uint_as_float(w2[j >> 1] & 0xffff0000u);
        st.s0[j] += lo; st.q0[j] = fmaf(lo, lo, st.q0[j]);
        st.s0[j + 1] += hi; st.q0[j + 1] = fmaf(hi, hi, st.q0[j + 1]);
      }
    } else if (SC > 1) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float lo = __uint_as_float(w2[j >> 1] << 16), hi = __uint_as_float(w2[j >> 1] & 0xffff0000u);
        st.s1[j] += lo; st.q1[j] = fmaf(lo, lo, st.q1[j]);
        st.s1[j + 1] += hi; st.q1[j + 1] = fmaf(hi, hi, st.q1[j + 1]);
      }
    }
  }
}

}  // namespace tc
}  // namespace mvd
