// touchnet_b200 :: shared definitions of the packed-sequence attention kernels.
#pragma once
#include "common.cuh"

namespace tn {

constexpr int ATT_BLK = 128;  // q-block and kv-block size (tokens)
constexpr int ATT_HD = 128;   // head_dim

// meta[b][blk] = {kv_lo, kv_end, q_end, canonical}
//   forward / dQ : q-block `blk` visits kv blocks [kv_lo, kv_end)   (kv_end == 0 -> every row is padding)
//   dK/dV        : kv-block `blk` is visited by q blocks [blk, q_end) (q_end == 0 -> no query attends here)
//   canonical    : 1 iff the row's ids are non-decreasing runs followed only by zeros (TouchNet's layout,
//                  touchnet/models/llama/processing_llama.py:37-40); otherwise ranges are conservative (everything
//                  causal) and every block is masked element-wise by doc-id compare, which is exact for any ids.
struct AttnMeta {
  int32_t kv_lo, kv_end, q_end, canonical;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// byte offset of 16-byte unit `u` (0..7) of row `r` inside a [rows x 128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(uint32_t r, uint32_t u) { return r * 128u + ((u ^ (r & 7u)) << 4); }

}  // namespace tn
