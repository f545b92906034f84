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

// per-position document extent: keys/queries of position t's document live in [start, end)  (start > t for padding)
struct AttnSeg {
  int32_t start, end;
};

// layout of the int32 "meta" buffer built by tn_attn_prep (all offsets in int32 units)
__host__ __device__ inline int64_t attn_meta_flags_off(int B, int nblk) { return int64_t(B) * nblk * 4; }
__host__ __device__ inline int64_t attn_meta_seg_off(int B, int nblk) {
  return (attn_meta_flags_off(B, nblk) + B + 3) / 4 * 4;
}
// cost-ordered work lists of the persistent kernels (entries b*nblk + blk, heaviest first):
//   order_q  : q blocks by kv_end - kv_lo      (forward, dQ)
//   order_kv : kv blocks by q_end - blk        (dK/dV)
__host__ __device__ inline int64_t attn_meta_order_off(int B, int nblk) {
  return attn_meta_seg_off(B, nblk) + int64_t(B) * nblk * ATT_BLK * 2;
}
__host__ __device__ inline int64_t attn_meta_order_kv_off(int B, int nblk) {
  return attn_meta_order_off(B, nblk) + int64_t(B) * nblk;
}
__host__ __device__ inline int64_t attn_meta_total(int B, int nblk) {
  return attn_meta_order_kv_off(B, nblk) + int64_t(B) * nblk;
}

// explicit shared-state-space accesses (generic LD/ST through a pointer derived from the dynamic smem base cost an
// address-space check per access and show up as LD.E / ST.E in the SASS)
__device__ __forceinline__ void sts_u4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// byte offset of 16-byte unit `u` (0..7) of row `r` inside a [rows x 128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(uint32_t r, uint32_t u) { return r * 128u + ((u ^ (r & 7u)) << 4); }

}  // namespace tn
