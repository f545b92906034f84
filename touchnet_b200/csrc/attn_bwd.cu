// touchnet_b200 :: packed-sequence ("document") causal attention, backward.  head_dim 128, bf16, GQA.
//
// Backward of the FlexAttention call the reference reaches through hf:integrations/flex_attention.py:262-364
// (autograd of torch.nn.attention.flex_attention with the block-causal document mask of :136-247).
//
// Three launches, no atomics, deterministic:
//   attn_delta_kernel            delta[b,h,t] = sum_d O.dO
//   attn_bwd_kernel<true>        one CTA per (kv block, kv head), walked heaviest first: loops over the q heads of the GQA
//                                group and the q blocks that attend to it;  S^T = K.Q^T, dP^T = V.dO^T (TMEM),
//                                dV += P^T.dO, dK += dS^T.Q accumulate in TMEM across the whole loop.
//   attn_bwd_kernel<false>       one CTA per (q block, head): S = Q.K^T, dP = dO.V^T, dQ += dS.K.
// Both use the same skeleton: two resident 128-row tiles, two streamed 64-row tiles (4-stage TMA ring), double buffered
// score tiles in TMEM, two softmax-grad warpgroups ping-ponging over the iterations (thread = resident row), one
// MMA-issuing thread.  P^T / dS^T never touch shared memory: they are written as bf16 over the score tiles they came
// from (tcgen05.st) and feed the accumulate MMAs as TMEM A operands (TS form) - the SS form of these MMAs read 8 KB of
// shared memory per 64 tensor-pipe cycles and made the kernels shared-memory-bandwidth bound.  Masking is decided per
// warp and per 32-column half from the per-row document extents (no compares on interior tiles, no exp2 / TMEM reads on
// fully masked ones).
#include "../../include/touchnet_b200.h"
#include "attn_common.cuh"
#include "host.h"

namespace tn {

constexpr int BWD_THREADS = 320;                      // TMA warp, MMA warp, 2 softmax-grad warpgroups (ping-pong)
constexpr int SUB = 64;                               // streamed rows per iteration
constexpr int RES_BYTES = ATT_BLK * ATT_HD * 2;       // 32 KB resident tile (2 chunks of 16 KB)
constexpr int RES_CHUNK = RES_BYTES / 2;
constexpr int STR_BYTES = SUB * ATT_HD * 2;           // 16 KB streamed tile (2 chunks of 8 KB)
constexpr int STR_CHUNK = STR_BYTES / 2;
constexpr int NST = 4;                                // streamed-tile ring depth

struct BwdSmem {
  static constexpr int R1 = 0;
  static constexpr int R2 = R1 + RES_BYTES;
  static constexpr int T = R2 + RES_BYTES;                  // NST stages x (T1, T2)
  static constexpr int COL = T + NST * 2 * STR_BYTES;        // 2 x {lse2[64], delta[64], doc[64]}
  static constexpr int BARS = COL + 2 * 3 * SUB * 4;
  static constexpr int TOTAL = BARS + 256;
  static constexpr int ALLOC = TOTAL + 1024;
};

struct AttnBwdParams {
  const int32_t* doc;
  const AttnMeta* meta;
  const AttnSeg* seg;
  const float* lse;
  const float* delta;
  bf16* out1;   // dKdV: dV ; dQ: dQ      (all-masked fast path)
  bf16* out2;   // dKdV: dK
  int64_t ld1, ld2;
  int B, T, H, KV, nblk;
  int Tq, q_blk_off;   // context parallelism: Q/dO/dQ/lse/delta hold rows [q_blk_off*128, +Tq) of the global sequence
  float scale, scale_log2;
  // fused inverse RoPE (backward of hf apply_rotary_pos_emb, modeling_llama.py:151-168) on dQ and dK in the epilogue:
  // cos/sin [rows, 64] bf16 indexed like the output tensor's rows (NULL = gradients w.r.t. the rotated q/k)
  const bf16* rope_cos;
  const bf16* rope_sin;
  // cost-ordered work list (tn_attn_prep: kv blocks by number of attending q blocks for dK/dV, q blocks by number of kv
  // blocks for dQ; entries b*nblk + blk, heaviest first) walked by a 1-D grid, so that the last wave is made of the
  // cheapest items; NULL = (block, head, batch) grid (context-parallel windows)
  const int32_t* order;
};

__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ O, int64_t ldo,
                                                         const bf16* __restrict__ dO, int64_t lddo,
                                                         float* __restrict__ delta, int B, int T, int H) {
  // half-warp per (b, t, h): 16 lanes x 8 elements
  const int64_t pair = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 4;
  const int sub = threadIdx.x & 15;
  const int64_t total = int64_t(B) * T * H;
  float acc = 0.f;
  if (pair < total) {
    const int h = int(pair % H);
    const int64_t bt = pair / H;
    const uint4 a = *reinterpret_cast<const uint4*>(O + bt * ldo + int64_t(h) * ATT_HD + sub * 8);
    const uint4 g = *reinterpret_cast<const uint4*>(dO + bt * lddo + int64_t(h) * ATT_HD + sub * 8);
    acc = bf16lo(a.x) * bf16lo(g.x) + bf16hi(a.x) * bf16hi(g.x) + bf16lo(a.y) * bf16lo(g.y) + bf16hi(a.y) * bf16hi(g.y) +
          bf16lo(a.z) * bf16lo(g.z) + bf16hi(a.z) * bf16hi(g.z) + bf16lo(a.w) * bf16lo(g.w) + bf16hi(a.w) * bf16hi(g.w);
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (pair < total && sub == 0) {
    const int h = int(pair % H);
    const int64_t bt = pair / H;
    const int64_t b = bt / T, t = bt - b * T;
    delta[(b * H + h) * T + t] = acc;
  }
}


// One 128 x 64 tile of the softmax gradient: thread = resident row, 64 streamed columns, handled as two 32-column halves.
//   p = exp2(x*scale_log2 - lse2)  (masked -> 0),  ds = p * (y - delta);  both packed to bf16 pairs.
// MASK 0: no mask; 1: allowed columns form the range [lo, hi]; 2: element-wise document-id compare (ids that are not
// non-decreasing runs).  Column vectors (lse2 | delta | doc, 64 entries each) sit in shared memory at `col_u32`; for dQ
// (DKDV = false) lse2 / delta are the thread's own row values.  Straight-line code per variant: a per-element branch
// serialises the MUFU chain (round 1 had one and ran the loop ~8x slower than the tensor pipe).
template <bool DKDV, int MASK>
__device__ __forceinline__ void bwd_half(int half, uint32_t x_t, uint32_t y_t, uint32_t col_u32, int lo, int hi,
                                         int32_t self_doc, int self_pos, int c0, float self_lse2, float self_delta,
                                         float scale_log2, uint32_t (&pk)[32], uint32_t (&dk)[32]) {
  uint32_t xv[32], yv[32];
  tmem_ld32(x_t + half * 32, xv);
  tmem_ld32(y_t + half * 32, yv);
  tmem_ld_wait();
#pragma unroll
  for (int i4 = 0; i4 < 8; ++i4) {
    const int cb = half * 32 + i4 * 4;
    float l2[4] = {self_lse2, self_lse2, self_lse2, self_lse2};
    float dl[4] = {self_delta, self_delta, self_delta, self_delta};
    if (DKDV) {
      const uint4 a = lds_u4(col_u32 + cb * 4), b = lds_u4(col_u32 + SUB * 4 + cb * 4);
      l2[0] = __uint_as_float(a.x); l2[1] = __uint_as_float(a.y); l2[2] = __uint_as_float(a.z); l2[3] = __uint_as_float(a.w);
      dl[0] = __uint_as_float(b.x); dl[1] = __uint_as_float(b.y); dl[2] = __uint_as_float(b.z); dl[3] = __uint_as_float(b.w);
    }
    int32_t id[4] = {0, 0, 0, 0};
    if (MASK == 2) {
      const uint4 d = lds_u4(col_u32 + 2 * SUB * 4 + cb * 4);
      id[0] = int32_t(d.x); id[1] = int32_t(d.y); id[2] = int32_t(d.z); id[3] = int32_t(d.w);
    }
    float pe[4], de[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = cb + e;
      bool ok = true;
      if (MASK == 1) ok = (c >= lo) & (c <= hi);
      if (MASK == 2) {
        const int col_pos = c0 + c;
        ok = (DKDV ? (self_pos <= col_pos) : (col_pos <= self_pos)) & (id[e] == self_doc) & (self_doc > 0);
      }
      const float xs = ok ? __uint_as_float(xv[i4 * 4 + e]) : -INFINITY;     // masked -> exp2(-inf) = 0
      pe[e] = fast_exp2(fmaf(xs, scale_log2, -l2[e]));
      de[e] = pe[e] * (__uint_as_float(yv[i4 * 4 + e]) - dl[e]);
    }
    pk[cb >> 1] = pack_bf16x2(pe[0], pe[1]);
    pk[(cb >> 1) + 1] = pack_bf16x2(pe[2], pe[3]);
    dk[cb >> 1] = pack_bf16x2(de[0], de[1]);
    dk[(cb >> 1) + 1] = pack_bf16x2(de[2], de[3]);
  }
}

// MODE 0: block pair inside one document, off the diagonal (no mask); 1: canonical ids - every 32-column half is
// classified per WARP (all rows allow every column -> no compares; no row allows any -> zeros, no TMEM read, no exp2;
// else range compares); 2: element-wise ids.
template <bool DKDV, int MODE>
__device__ __forceinline__ void bwd_softmax_grad(uint32_t x_t, uint32_t y_t, uint32_t col_u32, int lo, int hi,
                                                 int32_t self_doc, int self_pos, int c0, float self_lse2,
                                                 float self_delta, float scale_log2, uint32_t (&pk)[32],
                                                 uint32_t (&dk)[32]) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    if (MODE == 0) {
      bwd_half<DKDV, 0>(half, x_t, y_t, col_u32, lo, hi, self_doc, self_pos, c0, self_lse2, self_delta, scale_log2, pk, dk);
    } else if (MODE == 2) {
      bwd_half<DKDV, 2>(half, x_t, y_t, col_u32, lo, hi, self_doc, self_pos, c0, self_lse2, self_delta, scale_log2, pk, dk);
    } else {
      const bool fa = (lo <= half * 32) && (hi >= half * 32 + 31);
      const bool em = (hi < half * 32) || (lo > half * 32 + 31) || (lo > hi);
      if (__all_sync(0xffffffffu, em)) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { pk[half * 16 + i] = 0u; dk[half * 16 + i] = 0u; }
      } else if (__all_sync(0xffffffffu, fa)) {
        bwd_half<DKDV, 0>(half, x_t, y_t, col_u32, lo, hi, self_doc, self_pos, c0, self_lse2, self_delta, scale_log2, pk, dk);
      } else {
        bwd_half<DKDV, 1>(half, x_t, y_t, col_u32, lo, hi, self_doc, self_pos, c0, self_lse2, self_delta, scale_log2, pk, dk);
      }
    }
  }
}

// DKDV = true : resident (R1,R2) = (K,V) block `blk` of kv head `hy`; streamed (T1,T2) = (Q,dO) 64-row sub-blocks
// DKDV = false: resident (R1,R2) = (Q,dO) block `blk` of head `hy`;   streamed (T1,T2) = (K,V) 64-row sub-blocks
template <bool DKDV>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmR1, const __grid_constant__ CUtensorMap tmR2,
                const __grid_constant__ CUtensorMap tmT1, const __grid_constant__ CUtensorMap tmT2,
                const __grid_constant__ CUtensorMap tmOut1, const __grid_constant__ CUtensorMap tmOut2,
                const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sR1 = smem + BwdSmem::R1;
  uint8_t* sR2 = smem + BwdSmem::R2;
  uint8_t* sT = smem + BwdSmem::T;
  float* sCol = reinterpret_cast<float*>(smem + BwdSmem::COL);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BwdSmem::BARS);
  uint64_t* r_full = bars + 0;
  uint64_t* t_full = bars + 1;    // [NST]
  uint64_t* t_empty = bars + 5;   // [NST]
  uint64_t* xy_full = bars + 9;   // [2]
  uint64_t* pds_full = bars + 11; // [2]
  uint64_t* acc_done = bars + 13; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const uint32_t warp = warp_id(), lane = lane_id();
  const int q_off = p.q_blk_off * ATT_BLK;                          // global position of local query row 0
  const int nq_loc = (p.Tq + ATT_BLK - 1) / ATT_BLK;
  int blk_loc, hy, b;
  if (p.order) {
    const int heads = DKDV ? p.KV : p.H;
    const int s = int(blockIdx.x) / heads;
    hy = int(blockIdx.x) - s * heads;
    const int e = p.order[s];
    b = e / p.nblk;
    blk_loc = e - b * p.nblk;
  } else {
    blk_loc = DKDV ? int(blockIdx.x) : int(gridDim.x) - 1 - int(blockIdx.x);
    hy = blockIdx.y; b = blockIdx.z;
  }
  const int blk = DKDV ? blk_loc : blk_loc + p.q_blk_off;           // GLOBAL block index of the resident tile
  const int G = p.H / p.KV;
  const int r0 = blk * ATT_BLK;                                     // global position of resident row 0
  const int r0l = DKDV ? r0 : blk_loc * ATT_BLK;                    // row 0 inside the resident tensors (K/V global, Q/dO local)
  const int res_rows = DKDV ? p.T : p.Tq;                           // rows of the resident / output tensors
  const int s_off = DKDV ? q_off : 0;                               // streamed tensors: Q/dO are local, K/V global
  const AttnMeta meta = p.meta[b * p.nblk + blk];
  // streamed block range (global block indices) and iteration count
  const int sb_lo = DKDV ? max(blk, p.q_blk_off) : meta.kv_lo;
  const int sb_end = DKDV ? min(meta.q_end, p.q_blk_off + nq_loc) : meta.kv_end;
  const int nsb = sb_end > sb_lo ? sb_end - sb_lo : 0;
  const int n = (DKDV ? G : 1) * nsb * 2;

  if (n == 0) {
    // no (query, key) pair touches this block: gradients are exactly zero
    for (int i = threadIdx.x; i < ATT_BLK * (ATT_HD / 8); i += BWD_THREADS) {
      const int r = i / (ATT_HD / 8), c = i % (ATT_HD / 8);
      if (r0l + r < res_rows) {
        const int64_t tok = int64_t(b) * res_rows + r0l + r;
        *reinterpret_cast<uint4*>(p.out1 + tok * p.ld1 + int64_t(hy) * ATT_HD + c * 8) = make_uint4(0, 0, 0, 0);
        if (DKDV) *reinterpret_cast<uint4*>(p.out2 + tok * p.ld2 + int64_t(hy) * ATT_HD + c * 8) = make_uint4(0, 0, 0, 0);
      }
    }
    return;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmR1); tma_prefetch_desc(&tmR2); tma_prefetch_desc(&tmT1); tma_prefetch_desc(&tmT2);
    tma_prefetch_desc(&tmOut1);
    if (DKDV) tma_prefetch_desc(&tmOut2);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(r_full, 1);
    for (int s = 0; s < NST; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&xy_full[s], 1); mbar_init(&pds_full[s], 128); mbar_init(&acc_done[s], 1); }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc1 = tmem_base + 256, tmem_acc2 = tmem_base + 384;

  // iteration t -> (streamed head, streamed row0)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(r_full, 2 * RES_BYTES);
      tma_load_3d(sR1, &tmR1, r_full, hy * ATT_HD, r0l, b);
      tma_load_3d(sR1 + RES_CHUNK, &tmR1, r_full, hy * ATT_HD + 64, r0l, b);
      tma_load_3d(sR2, &tmR2, r_full, hy * ATT_HD, r0l, b);
      tma_load_3d(sR2 + RES_CHUNK, &tmR2, r_full, hy * ATT_HD + 64, r0l, b);
      int pu = 0, phd = 0;                                             // position inside the head's streamed range, head offset
      for (int t = 0; t < n; ++t) {
        const int s = t % NST;
        mbar_wait(&t_empty[s], ((t / NST) & 1) ^ 1);
        mbar_arrive_expect_tx(&t_full[s], 2 * STR_BYTES);
        uint8_t* d1 = sT + s * 2 * STR_BYTES;
        uint8_t* d2 = d1 + STR_BYTES;
        const int hs = DKDV ? hy * G + phd : hy / G;
        const int row0 = (sb_lo + (pu >> 1)) * ATT_BLK + (pu & 1) * SUB - s_off;   // row inside the streamed tensor
        if (++pu == nsb * 2 && DKDV) { pu = 0; ++phd; }
        tma_load_3d(d1, &tmT1, &t_full[s], hs * ATT_HD, row0, b);
        tma_load_3d(d1 + STR_CHUNK, &tmT1, &t_full[s], hs * ATT_HD + 64, row0, b);
        tma_load_3d(d2, &tmT2, &t_full[s], hs * ATT_HD, row0, b);
        tma_load_3d(d2 + STR_CHUNK, &tmT2, &t_full[s], hs * ATT_HD + 64, row0, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_xy = make_idesc_bf16(128, SUB, 0, 0);
      constexpr uint32_t idesc_acc = make_idesc_bf16(128, ATT_HD, 0, 1);
      const uint32_t r1 = smem_u32(sR1), r2 = smem_u32(sR2);
      auto issue_xy = [&](int t) {
        const int s = t % NST;
        mbar_wait(&t_full[s], (t / NST) & 1);
        tc_fence_after();
        const uint32_t t1 = smem_u32(sT + s * 2 * STR_BYTES), t2 = t1 + STR_BYTES;
        const uint32_t x_t = tmem_base + (t & 1) * 128, y_t = x_t + 64;
#pragma unroll
        for (int k = 0; k < ATT_HD / 16; ++k) {
          const uint32_t ro = (k >> 2) * RES_CHUNK + (k & 3) * 32, so = (k >> 2) * STR_CHUNK + (k & 3) * 32;
          umma_ss(x_t, make_sdesc_sw128(r1 + ro, 0, 1024), make_sdesc_sw128(t1 + so, 0, 1024), idesc_xy, k > 0);
        }
#pragma unroll
        for (int k = 0; k < ATT_HD / 16; ++k) {
          const uint32_t ro = (k >> 2) * RES_CHUNK + (k & 3) * 32, so = (k >> 2) * STR_CHUNK + (k & 3) * 32;
          umma_ss(y_t, make_sdesc_sw128(r2 + ro, 0, 1024), make_sdesc_sw128(t2 + so, 0, 1024), idesc_xy, k > 0);
        }
        umma_commit(&xy_full[t & 1]);
      };
      mbar_wait(r_full, 0);
      issue_xy(0);
      for (int t = 0; t < n; ++t) {
        if (t + 1 < n) issue_xy(t + 1);
        const int s = t % NST;
        mbar_wait(&pds_full[t & 1], (t >> 1) & 1);
        tc_fence_after();
        const uint32_t t1 = smem_u32(sT + s * 2 * STR_BYTES), t2 = t1 + STR_BYTES;
        // P^T / dS^T are bf16 tiles the softmax-grad warps wrote OVER the score tiles they came from (TMEM A operand,
        // tcgen05.mma TS form): columns [0,32) of the S^T and of the dP^T tile of this iteration's buffer
        const uint32_t pt = tmem_base + (t & 1) * 128, dst = pt + 64;
#pragma unroll
        for (int k = 0; k < SUB / 16; ++k) {
          if (DKDV) {
            // dV += P^T.dO   (B = dO tile, MN-major: hd contiguous)      dK += dS^T.Q
            umma_ts(tmem_acc1, pt + k * 8, make_sdesc_sw128(t2 + k * 2048, STR_CHUNK, 1024), idesc_acc, (t > 0 || k > 0) ? 1u : 0u);
            umma_ts(tmem_acc2, dst + k * 8, make_sdesc_sw128(t1 + k * 2048, STR_CHUNK, 1024), idesc_acc, (t > 0 || k > 0) ? 1u : 0u);
          } else {
            // dQ += dS.K
            umma_ts(tmem_acc1, dst + k * 8, make_sdesc_sw128(t1 + k * 2048, STR_CHUNK, 1024), idesc_acc, (t > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(&t_empty[s]);
        umma_commit(&acc_done[t & 1]);
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax-grad warps: thread = resident row =====================
    // two warpgroups: group g handles iterations t = g, g+2, ... (its own score buffer, P/dS tiles, column vectors and
    // named barrier), so one group's exp/convert work overlaps the other's and the tensor pipe sees back-to-back MMAs
    const int grp = (int(warp) - 2) >> 2;
    const uint32_t quad = warp & 3u;
    const uint32_t r = quad * 32 + lane;
    const int tid = int(threadIdx.x) - 64 - 128 * grp;
    const int self_pos = r0 + int(r);
    const int32_t* docb = p.doc + int64_t(b) * p.T;
    const int32_t self_doc = (self_pos < p.T) ? docb[self_pos] : 0;
    const uint32_t lane_sel = (quad * 32u) << 16;
    // canonical rows: the partner positions of this row are one contiguous range (two integer compares per element)
    //   dK/dV (thread = key k):   queries q in [k, seg_end(k))        dQ (thread = query q): keys k in [seg_start(q), q]
    const AttnSeg myseg = (self_pos < p.T) ? p.seg[int64_t(b) * p.nblk * ATT_BLK + self_pos] : AttnSeg{self_pos + 1, self_pos};
    const int range_lo_pos = DKDV ? self_pos : myseg.start;
    const int range_hi_pos = DKDV ? myseg.end - 1 : self_pos;
    float self_lse2 = 0.f, self_delta = 0.f;
    if (!DKDV) {
      const int lpos = self_pos - q_off;                                // local query row
      const int64_t idx = (int64_t(b) * p.H + hy) * p.Tq + lpos;
      self_lse2 = (lpos < p.Tq) ? p.lse[idx] * 1.4426950408889634f : __int_as_float(0x7f800000);
      self_delta = (lpos < p.Tq) ? p.delta[idx] : 0.f;
    }

    // column vectors (for dK/dV: lse, delta of the streamed q rows; doc ids for the element-wise mask mode) are fetched one
    // iteration ahead so that their global-memory latency is off the critical path.  Iterations are decoded incrementally
    // (head offset, position inside the head's streamed range): no integer division and no dependent global load sits
    // between two tiles (round 1 had both; together they cost more than the softmax gradient itself).
    const int nsb2 = nsb * 2;
    const bool need_cols = DKDV || !meta.canonical;
    auto row0_of = [&](int u) { return (sb_lo + (u >> 1)) * ATT_BLK + (u & 1) * SUB; };
    auto fetch_col = [&](int hd, int u, int32_t& d_out, float& f_out) {     // raw values; scaled when staged
      const int c = tid & 63, pos = row0_of(u) + c;
      const int hs_ = DKDV ? hy * G + hd : hy / G;
      d_out = -2; f_out = 0.f;
      if (tid < 64) {
        d_out = (pos < p.T) ? docb[pos] : -2;
        if (DKDV) f_out = (pos - q_off < p.Tq) ? p.lse[(int64_t(b) * p.H + hs_) * p.Tq + pos - q_off] : __int_as_float(0x7f800000);
      } else if (DKDV) {
        f_out = (pos - q_off < p.Tq) ? p.delta[(int64_t(b) * p.H + hs_) * p.Tq + pos - q_off] : 0.f;
      }
    };
    auto adv2 = [&](int& hd, int& u) { u += 2; if (DKDV) { while (u >= nsb2) { u -= nsb2; ++hd; } } };
    int hd_cur = 0, u_cur = grp;                       // decode of iteration t = grp (grp < 2 <= nsb2)
    if (DKDV) { while (u_cur >= nsb2) { u_cur -= nsb2; ++hd_cur; } }
    int hd_nxt = hd_cur, u_nxt = u_cur;
    int32_t nxt_doc = -2; float nxt_f = 0.f;
    if (grp < n && need_cols) fetch_col(hd_cur, u_cur, nxt_doc, nxt_f);

    for (int t = grp; t < n; t += 2) {
      const int c0 = row0_of(u_cur);
      mbar_wait(&xy_full[t & 1], (t >> 1) & 1);
      tc_fence_after();
      float* col = sCol + (t & 1) * 3 * SUB;
      adv2(hd_nxt, u_nxt);
      if (need_cols) {
        float* col_lse2 = col;
        float* col_delta = col + SUB;
        int32_t* col_doc = reinterpret_cast<int32_t*>(col + 2 * SUB);
        const int c = tid & 63;
        if (tid < 64) { col_doc[c] = nxt_doc; if (DKDV) col_lse2[c] = nxt_f * 1.4426950408889634f; }
        else if (DKDV) col_delta[c] = nxt_f;
        if (t + 2 < n) fetch_col(hd_nxt, u_nxt, nxt_doc, nxt_f);   // in flight during this iteration's math
        named_bar_sync(1 + grp, 128);
      }
      hd_cur = hd_nxt; u_cur = u_nxt;

      const uint32_t x_t = tmem_base + (t & 1) * 128 + lane_sel, y_t = x_t + 64;
      uint32_t pk[32], dk[32];
      const int lo = range_lo_pos - c0, hi = range_hi_pos - c0;   // allowed streamed columns: lo <= c <= hi
      const uint32_t col_u32 = smem_u32(col);
      // straight-line element code per mask mode (mode is warp-uniform; a per-element branch serialises the MUFU chain)
      // canonical ids: every 32-column half is classified by warp votes on the per-row ranges held in registers
      if (meta.canonical) bwd_softmax_grad<DKDV, 1>(x_t, y_t, col_u32, lo, hi, self_doc, self_pos, c0, self_lse2, self_delta, p.scale_log2, pk, dk);
      else bwd_softmax_grad<DKDV, 2>(x_t, y_t, col_u32, lo, hi, self_doc, self_pos, c0, self_lse2, self_delta, p.scale_log2, pk, dk);
      // P^T / dS^T -> TMEM, in place over this thread's own lane of the score tiles (every value of the lane was read above;
      // the buffer's next writer, xy(t+2), is issued behind this iteration's accumulate MMAs in the in-order tensor pipe)
      if (DKDV) tmem_st32(x_t, pk);
      tmem_st32(y_t, dk);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&pds_full[t & 1]);
    }

    // ---- epilogue: accumulators -> bf16 -> smem staging (streamed-tile ring is idle now) -> TMA store ----
    // group 0 writes accumulator 1 (dV | dQ), group 1 accumulator 2 (dK)
    // every accumulate has landed once the LAST one has (in-order pipe).  A group first waits for its own last iteration
    // (its own barrier, next phase in sequence), then for the last one overall: waiting on the other group's barrier
    // directly could alias - when this group races ahead (fully masked tiles cost it almost nothing) that barrier may
    // still be two phases behind, and a parity wait cannot tell phase k-2 from phase k.
    {
      const int last_own = (n - 1 >= grp) ? (n - 1) - ((n - 1 - grp) & 1) : -1;
      if (last_own >= 0 && last_own != n - 1) mbar_wait(&acc_done[last_own & 1], (last_own >> 1) & 1);
      mbar_wait(&acc_done[(n - 1) & 1], ((n - 1) >> 1) & 1);
    }
    tc_fence_after();
    if (DKDV || grp == 0) {
      const uint32_t acc_t = (grp == 0 ? tmem_acc1 : tmem_acc2) + lane_sel;
      uint8_t* stg = sT + grp * RES_BYTES;
      const float mul = (DKDV && grp == 0) ? 1.f : p.scale;  // dV unscaled; dK, dQ carry the softmax scale
      const bool rope = p.rope_cos != nullptr && !(DKDV && grp == 0);   // dV is never rotated
      if (!rope) {
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          uint32_t o[32];
          tmem_ld32(acc_t + c4 * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              w[e] = pack_bf16x2(__uint_as_float(o[u * 8 + 2 * e]) * mul, __uint_as_float(o[u * 8 + 2 * e + 1]) * mul);
            const int colx = c4 * 32 + u * 8;
            sts_u4(smem_u32(stg) + (colx >> 6) * RES_CHUNK + sw128_off(r, (colx & 63) >> 3), make_uint4(w[0], w[1], w[2], w[3]));
          }
        }
      } else {
        // d/dx of y = x*cos + rotate_half(x)*sin :  dx_lo = g_lo*cos + g_hi*sin,  dx_hi = g_hi*cos - g_lo*sin
        // (pairs (j, j+64) live in column chunks c4 and c4+2 of this thread's row)
        const int64_t trow = int64_t(b) * res_rows + r0l + int(r);
        const bool row_ok = r0l + int(r) < res_rows;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t lo[32], hi[32];
          tmem_ld32(acc_t + half * 32, lo);
          tmem_ld32(acc_t + 64 + half * 32, hi);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 c4v = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u), s4v = make_uint4(0, 0, 0, 0);
            if (row_ok) {
              c4v = *reinterpret_cast<const uint4*>(p.rope_cos + trow * 64 + half * 32 + u * 8);
              s4v = *reinterpret_cast<const uint4*>(p.rope_sin + trow * 64 + half * 32 + u * 8);
            }
            const uint32_t cw[4] = {c4v.x, c4v.y, c4v.z, c4v.w}, sw[4] = {s4v.x, s4v.y, s4v.z, s4v.w};
            uint32_t wl[4], wh[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float gl0 = __uint_as_float(lo[u * 8 + 2 * e]) * mul, gl1 = __uint_as_float(lo[u * 8 + 2 * e + 1]) * mul;
              const float gh0 = __uint_as_float(hi[u * 8 + 2 * e]) * mul, gh1 = __uint_as_float(hi[u * 8 + 2 * e + 1]) * mul;
              const float c0 = bf16lo(cw[e]), c1 = bf16hi(cw[e]), s0 = bf16lo(sw[e]), s1 = bf16hi(sw[e]);
              wl[e] = pack_bf16x2(fmaf(gl0, c0, gh0 * s0), fmaf(gl1, c1, gh1 * s1));
              wh[e] = pack_bf16x2(fmaf(gh0, c0, -gl0 * s0), fmaf(gh1, c1, -gl1 * s1));
            }
            const uint32_t unit = uint32_t(half * 4 + u);     // 16-byte unit inside the 64-column chunk
            sts_u4(smem_u32(stg) + sw128_off(r, unit), make_uint4(wl[0], wl[1], wl[2], wl[3]));
            sts_u4(smem_u32(stg) + RES_CHUNK + sw128_off(r, unit), make_uint4(wh[0], wh[1], wh[2], wh[3]));
          }
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + grp, 128);
      if (tid == 0) {
        const CUtensorMap* tmo = (grp == 0) ? &tmOut1 : &tmOut2;
        tma_store_3d(tmo, stg, hy * ATT_HD, r0l, b);
        tma_store_3d(tmo, stg + RES_CHUNK, hy * ATT_HD + 64, r0l, b);
        tma_store_commit();
        tma_store_wait_read<0>();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace tn

using namespace tn;

extern "C" int tn_attn_bwd_bf16(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv,
                                const void* O, int64_t ldo, const void* dO, int64_t lddo, const float* lse, float* delta,
                                void* dQ, int64_t lddq, void* dK, int64_t lddk, void* dV, int64_t lddv,
                                const int32_t* doc_ids, const int32_t* meta, int B, int T, int H, int KV, float scale,
                                int Tq, int q_blk_off, const void* rope_cos_q, const void* rope_sin_q,
                                const void* rope_cos_k, const void* rope_sin_k, tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(Q && K && V && O && dO && lse && delta && dQ && dK && dV && doc_ids && meta, "tn_attn_bwd_bf16: null pointer");
  TN_REQUIRE((rope_cos_q == nullptr) == (rope_sin_q == nullptr) && (rope_cos_k == nullptr) == (rope_sin_k == nullptr),
             "tn_attn_bwd_bf16: cos and sin tables come in pairs");
  TN_REQUIRE(B > 0 && T > 0 && H > 0 && KV > 0 && H % KV == 0, "tn_attn_bwd_bf16: bad dims");
  TN_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && lddo % 8 == 0 && lddq % 8 == 0 &&
                 lddk % 8 == 0 && lddv % 8 == 0, "tn_attn_bwd_bf16: strides must be multiples of 8");
  const int nblk = (T + ATT_BLK - 1) / ATT_BLK;
  if (Tq <= 0) { Tq = T; q_blk_off = 0; }
  TN_REQUIRE(q_blk_off >= 0 && q_blk_off * ATT_BLK + Tq <= nblk * ATT_BLK, "tn_attn_bwd_bf16: query window outside the sequence");
  const int nqb = (Tq + ATT_BLK - 1) / ATT_BLK;

  {
    const int64_t pairs = int64_t(B) * Tq * H;
    const int64_t threads = pairs * 16;
    attn_delta_kernel<<<unsigned((threads + 255) / 256), 256, 0, stream>>>(
        static_cast<const bf16*>(O), ldo, static_cast<const bf16*>(dO), lddo, delta, B, Tq, H);
    TN_CHECK_CUDA(cudaGetLastError());
  }

  CUtensorMap q128, q64, do128, do64, k128, k64, v128, v64, mdq, mdk, mdv;
  int rc;
#define TN_MAP(m, ptr, ld, heads, box, rows)                                                                        \
  if ((rc = encode_tmap_3d(&m, ptr, 2, uint64_t(heads) * ATT_HD, rows, B, uint64_t(ld) * 2, uint64_t(rows) * (ld) * 2, 64, \
                           box, 1, true)))                                                                          \
    return rc;
  TN_MAP(q128, Q, ldq, H, ATT_BLK, Tq) TN_MAP(q64, Q, ldq, H, SUB, Tq)
  TN_MAP(do128, dO, lddo, H, ATT_BLK, Tq) TN_MAP(do64, dO, lddo, H, SUB, Tq)
  TN_MAP(k128, K, ldk, KV, ATT_BLK, T) TN_MAP(k64, K, ldk, KV, SUB, T)
  TN_MAP(v128, V, ldv, KV, ATT_BLK, T) TN_MAP(v64, V, ldv, KV, SUB, T)
  TN_MAP(mdq, dQ, lddq, H, ATT_BLK, Tq) TN_MAP(mdk, dK, lddk, KV, ATT_BLK, T) TN_MAP(mdv, dV, lddv, KV, ATT_BLK, T)
#undef TN_MAP

  AttnBwdParams p{};
  p.doc = doc_ids; p.meta = reinterpret_cast<const AttnMeta*>(meta); p.lse = lse; p.delta = delta;
  p.seg = reinterpret_cast<const AttnSeg*>(meta + attn_meta_seg_off(B, nblk));
  p.B = B; p.T = T; p.H = H; p.KV = KV; p.nblk = nblk;
  p.Tq = Tq; p.q_blk_off = q_blk_off;
  p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f;

  static bool configured = false;
  if (!configured) {
    TN_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::ALLOC));
    TN_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::ALLOC));
    configured = true;
  }
  {
    AttnBwdParams pk = p;
    pk.out1 = static_cast<bf16*>(dV); pk.ld1 = lddv; pk.out2 = static_cast<bf16*>(dK); pk.ld2 = lddk;
    pk.rope_cos = static_cast<const bf16*>(rope_cos_k); pk.rope_sin = static_cast<const bf16*>(rope_sin_k);
    const bool ordered = (Tq == T && q_blk_off == 0);
    pk.order = ordered ? meta + attn_meta_order_kv_off(B, nblk) : nullptr;
    const dim3 grid = ordered ? dim3(unsigned(nblk) * KV * B) : dim3(nblk, KV, B);
    attn_bwd_kernel<true><<<grid, BWD_THREADS, BwdSmem::ALLOC, stream>>>(k128, v128, q64, do64, mdv, mdk, pk);
    TN_CHECK_CUDA(cudaGetLastError());
  }
  {
    AttnBwdParams pq = p;
    pq.out1 = static_cast<bf16*>(dQ); pq.ld1 = lddq; pq.out2 = nullptr; pq.ld2 = 0;
    pq.rope_cos = static_cast<const bf16*>(rope_cos_q); pq.rope_sin = static_cast<const bf16*>(rope_sin_q);
    const bool ordered = (Tq == T && q_blk_off == 0);
    pq.order = ordered ? meta + attn_meta_order_off(B, nblk) : nullptr;
    const dim3 grid = ordered ? dim3(unsigned(nqb) * H * B) : dim3(nqb, H, B);
    attn_bwd_kernel<false><<<grid, BWD_THREADS, BwdSmem::ALLOC, stream>>>(q128, do128, k64, v64, mdq, mdq, pq);
    TN_CHECK_CUDA(cudaGetLastError());
  }
  return TN_OK;
}
