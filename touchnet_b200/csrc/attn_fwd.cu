// touchnet_b200 :: packed-sequence ("document") causal attention, forward.  head_dim 128, bf16, GQA.
//
// Replaces torch.nn.attention.flex_attention + make_flex_block_causal_mask as reached from
//   hf:integrations/flex_attention.py:136-247, :262-364  (selected by `"attn_implementation": "flex_attention"`,
//   examples/text/pretrain/fineweb-edu/config/Llama-3_2-1B.json:7; touchnet/bin/train.py:129-131)
// mask  allow[b,q,k] = (q >= k) && doc[b,q] == doc[b,k] && doc[b,q] > 0
//
// One CTA = one 128-row q block of one head.  6 warps:
//   warp 0      TMA producer (Q once; K/V 128-row tiles through 2-stage rings) + TMEM alloc
//   warp 1      MMA issuer   (S = Q·Kᵀ into a double-buffered TMEM tile; O += P·V)
//   warps 2..5  softmax      (thread = row: tcgen05.ld S, doc-id/causal mask in registers, online softmax with lazy
//                             rescale, P -> smem (128B-swizzled, A operand of the PV MMA), final O/l -> TMA store)
// kv blocks outside the [kv_lo, kv_end) range of the q block are never touched (block-level doc skipping);
// blocks wholly inside one document below the diagonal skip the mask arithmetic.
#include "../../include/touchnet_b200.h"
#include "attn_common.cuh"
#include "host.h"

#include <stdlib.h>

namespace tn {

// ---------------------------------------------------------------------------------------------------------------
// prep: canonical flag per batch row, then per-block ranges
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) attn_canonical_kernel(const int32_t* __restrict__ doc, int32_t* __restrict__ flags,
                                                              int T) {
  const int b = blockIdx.x;
  const int32_t* d = doc + int64_t(b) * T;
  int bad = 0;
  for (int t = threadIdx.x + 1; t < T; t += blockDim.x) {
    const int32_t cur = d[t], prev = d[t - 1];
    if (cur != 0 && (prev == 0 || cur < prev)) bad = 1;
    if (cur < 0) bad = 1;
  }
  if (threadIdx.x == 0 && T > 0 && d[0] < 0) bad = 1;
  bad = __syncthreads_or(bad);
  if (threadIdx.x == 0) flags[b] = bad ? 0 : 1;
}

// one warp per (b, block)
__global__ void __launch_bounds__(256) attn_meta_kernel(const int32_t* __restrict__ doc, const int32_t* __restrict__ flags,
                                                        AttnMeta* __restrict__ meta, AttnSeg* __restrict__ seg, int B,
                                                        int T, int nblk) {
  const int gw = blockIdx.x * 8 + warp_id();
  if (gw >= B * nblk) return;
  const int b = gw / nblk, blk = gw - b * nblk;
  const uint32_t lane = lane_id();
  const int32_t* d = doc + int64_t(b) * T;
  const int t0 = blk * ATT_BLK;
  const int canonical = flags[b];
  // the 128 ids of this block: lane holds positions t0 + lane + 32*i
  int32_t ids[4];
  uint32_t valid_mask[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + int(lane) + 32 * i;
    ids[i] = (t < T) ? d[t] : 0;
    valid_mask[i] = __ballot_sync(0xffffffffu, ids[i] > 0);
  }
  int first = -1, last = -1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (valid_mask[i]) {
      if (first < 0) first = 32 * i + __ffs(valid_mask[i]) - 1;
      last = 32 * i + 31 - __clz(valid_mask[i]);
    }
  }
  AttnMeta m;
  m.canonical = canonical;
  AttnSeg* segb = seg + (int64_t(b) * nblk + blk) * ATT_BLK;
  if (first < 0 || !canonical) {
    // padding-only block, or ids that are not non-decreasing runs: segments are empty / unused (the kernels mask
    // non-canonical rows element-wise by document id over the whole causal range)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = t0 + int(lane) + 32 * i;
      segb[lane + 32 * i] = AttnSeg{t + 1, t};
    }
    if (first < 0) { m.kv_lo = 0; m.kv_end = 0; m.q_end = 0; }
    else { m.kv_lo = 0; m.kv_end = blk + 1; m.q_end = nblk; }
    if (lane == 0) meta[gw] = m;
    return;
  }
  // ---- backward scan: start of the run containing the first valid row ----
  const int32_t d_first = __shfl_sync(0xffffffffu, ids[first >> 5], first & 31);
  int run_start = t0 + first;
  {
    // inside the block
    bool done = false;
    for (int t = t0 + first - 1; t >= t0 && !done; --t) {  // <= 127 steps, warp-uniform
      const int32_t v = __shfl_sync(0xffffffffu, ids[(t - t0) >> 5], (t - t0) & 31);
      if (v == d_first) run_start = t; else done = true;
    }
    if (!done) {
      int base = t0 - 128;  // earlier blocks, 128 ids per step
      while (base >= 0 && !done) {
        int32_t v[4];
        uint32_t eq[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[i] = d[base + int(lane) + 32 * i];
          eq[i] = __ballot_sync(0xffffffffu, v[i] == d_first);
        }
        // canonical => matching ids form a suffix of this window
#pragma unroll
        for (int i = 3; i >= 0; --i) {
          if (done) break;
          if (eq[i] == 0xffffffffu) { run_start = base + 32 * i; }
          else {
            if (eq[i]) run_start = base + 32 * i + (32 - __clz(~eq[i]));
            done = true;
          }
        }
        base -= 128;
      }
    }
  }
  m.kv_lo = run_start / ATT_BLK;
  m.kv_end = blk + 1;
  // ---- forward scan: end (exclusive) of the run containing the last valid column ----
  const int32_t d_last = __shfl_sync(0xffffffffu, ids[last >> 5], last & 31);
  int run_end = t0 + last + 1;
  if (last == ATT_BLK - 1) {
    bool done = false;
    int base = t0 + 128;
    while (base < T && !done) {
      uint32_t eq[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = base + int(lane) + 32 * i;
        const int32_t v = (t < T) ? d[t] : 0;
        eq[i] = __ballot_sync(0xffffffffu, v == d_last);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (done) break;
        if (eq[i] == 0xffffffffu) { run_end = base + 32 * i + 32; }
        else {
          if (eq[i]) run_end = base + 32 * i + (__ffs(~eq[i]) - 1);
          done = true;
        }
      }
      base += 128;
    }
  }
  m.q_end = (run_end - 1) / ATT_BLK + 1;
  if (lane == 0) meta[gw] = m;

  // ---- per-position document extents [start, end) ----
  // a run starts at p if p == 0 of the block or id[p] != id[p-1]; it ends after p if id[p+1] != id[p]
  uint32_t sb[4], eb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pos = int(lane) + 32 * i;
    const int32_t prev = __shfl_up_sync(0xffffffffu, ids[i], 1);
    const int32_t prev_w = __shfl_sync(0xffffffffu, ids[i > 0 ? i - 1 : 0], 31);
    const int32_t next = __shfl_down_sync(0xffffffffu, ids[i], 1);
    const int32_t next_w = __shfl_sync(0xffffffffu, ids[i < 3 ? i + 1 : 3], 0);
    const int32_t pv = lane > 0 ? prev : (i > 0 ? prev_w : ids[i] /* block edge: resolved by run_start */);
    const int32_t nx = lane < 31 ? next : (i < 3 ? next_w : ids[i] /* block edge: resolved by run_end */);
    sb[i] = __ballot_sync(0xffffffffu, pos > 0 && pv != ids[i]);
    eb[i] = __ballot_sync(0xffffffffu, pos < ATT_BLK - 1 && nx != ids[i]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pos = int(lane) + 32 * i;
    const int t = t0 + pos;
    AttnSeg sg;
    if (ids[i] <= 0) { sg.start = t + 1; sg.end = t; }
    else {
      // last run boundary at or before pos
      int st = -1;
      {
        const uint32_t mcur = sb[i] & (0xffffffffu >> (31 - lane));
        if (mcur) st = 32 * i + 31 - __clz(mcur);
        else {
#pragma unroll
          for (int w = 3; w >= 0; --w)
            if (w < i && st < 0 && sb[w]) st = 32 * w + 31 - __clz(sb[w]);
        }
      }
      sg.start = st >= 0 ? t0 + st : run_start;       // no boundary inside the block: the run began earlier
      int en = -1;
      {
        const uint32_t mcur = eb[i] & (0xffffffffu << lane);
        if (mcur) en = 32 * i + __ffs(mcur) - 1;
        else {
#pragma unroll
          for (int w = 0; w < 4; ++w)
            if (w > i && en < 0 && eb[w]) en = 32 * w + __ffs(eb[w]) - 1;
        }
      }
      sg.end = en >= 0 ? t0 + en + 1 : run_end;        // the run continues past the block (or ends at its last valid id)
    }
    segb[pos] = sg;
  }
}

// cost-ordered work lists (one CTA; N = B*nblk entries, rank by counting: O(N^2 / 1024) compares per thread, once per step)
__global__ void __launch_bounds__(1024) attn_order_kernel(const AttnMeta* __restrict__ meta, int32_t* __restrict__ order_q,
                                                          int32_t* __restrict__ order_kv, int N, int nblk) {
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const AttnMeta m = meta[i];
    const int blk = i % nblk;
    const int cq = m.kv_end - m.kv_lo;
    const int ckv = m.q_end > blk ? m.q_end - blk : 0;
    int rq = 0, rkv = 0;
    for (int t = 0; t < N; ++t) {
      const AttnMeta o = meta[t];
      const int oq = o.kv_end - o.kv_lo;
      const int ob = t % nblk;
      const int okv = o.q_end > ob ? o.q_end - ob : 0;
      rq += (oq > cq) || (oq == cq && t < i);
      rkv += (okv > ckv) || (okv == ckv && t < i);
    }
    order_q[rq] = i;
    order_kv[rkv] = i;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// forward kernel (round 1, one CTA per (q block, head); kept for A/B runs behind TN_ATTN_FWD_V1=1)
// ---------------------------------------------------------------------------------------------------------------
constexpr int FWD_THREADS = 192;
constexpr int TILE_BYTES = ATT_BLK * ATT_HD * 2;  // 32 KB: two [128 x 128 B] swizzled chunks
constexpr int CHUNK_BYTES = TILE_BYTES / 2;

struct FwdSmem {
  static constexpr int Q = 0;
  static constexpr int K = Q + TILE_BYTES;       // 2 stages
  static constexpr int V = K + 2 * TILE_BYTES;   // 2 stages
  static constexpr int P = V + 2 * TILE_BYTES;   // P tile / O staging
  static constexpr int DOCK = P + TILE_BYTES;    // 2 x 128 int32
  static constexpr int BARS = DOCK + 2 * 128 * 4;
  static constexpr int TOTAL = BARS + 256;
  static constexpr int ALLOC = TOTAL + 1024;
};

struct AttnFwdParams {
  const int32_t* doc;
  const AttnMeta* meta;
  const AttnSeg* seg;
  float* lse;
  bf16* O;       // for the all-padding fast path
  int64_t ldo;
  int B, T, H, KV, nblk;
  int Tq, q_blk_off;  // context parallelism: the Q/O/lse tensors hold rows [q_blk_off*128, q_blk_off*128 + Tq) of the
                      // global sequence of length T; K/V/doc/meta are global.  Tq == T, off == 0 without CP.
  float scale_log2;  // softmax scale * log2(e)
};

__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + FwdSmem::Q;
  uint8_t* sK = smem + FwdSmem::K;
  uint8_t* sV = smem + FwdSmem::V;
  uint8_t* sP = smem + FwdSmem::P;
  int32_t* sDocK = reinterpret_cast<int32_t*>(smem + FwdSmem::DOCK);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::BARS);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* p_full = bars + 11;
  uint64_t* pv_done = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const uint32_t warp = warp_id(), lane = lane_id();
  const int qb_loc = int(gridDim.x) - 1 - int(blockIdx.x);  // heaviest (latest) q blocks first
  const int qb = qb_loc + p.q_blk_off;                      // global block index (meta / doc / seg / masking)
  const int h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (p.H / p.KV);
  const int q0 = qb * ATT_BLK;          // global position of row 0
  const int q0l = qb_loc * ATT_BLK;     // row 0 inside the (local) Q / O / lse tensors
  const AttnMeta meta = p.meta[b * p.nblk + qb];
  const int kv_lo = meta.kv_lo;
  const int n = meta.kv_end - meta.kv_lo;

  if (meta.kv_end == 0) {
    // every row of this q block is padding: O = 0 exactly (FlexAttention semantics), lse = +inf
    const int tid = threadIdx.x;
    for (int i = tid; i < ATT_BLK * (ATT_HD / 8); i += FWD_THREADS) {
      const int r = i / (ATT_HD / 8), c = i % (ATT_HD / 8);
      if (q0l + r < p.Tq)
        *reinterpret_cast<uint4*>(p.O + (int64_t(b) * p.Tq + q0l + r) * p.ldo + int64_t(h) * ATT_HD + c * 8) =
            make_uint4(0, 0, 0, 0);
    }
    for (int r = tid; r < ATT_BLK; r += FWD_THREADS)
      if (q0l + r < p.Tq) p.lse[(int64_t(b) * p.H + h) * p.Tq + q0l + r] = __int_as_float(0x7f800000);
    return;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1);
    }
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;          // 2 x 128 columns
  const uint32_t tmem_O = tmem_base + 256;    // 128 columns

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tma_load_3d(sQ, &tmQ, q_full, h * ATT_HD, q0l, b, kEvictFirst);
      tma_load_3d(sQ + CHUNK_BYTES, &tmQ, q_full, h * ATT_HD + 64, q0l, b, kEvictFirst);
      for (int j = 0; j < n; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const int k0 = (kv_lo + j) * ATT_BLK;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], TILE_BYTES);
        tma_load_3d(sK + s * TILE_BYTES, &tmK, &k_full[s], kvh * ATT_HD, k0, b, kEvictLast);
        tma_load_3d(sK + s * TILE_BYTES + CHUNK_BYTES, &tmK, &k_full[s], kvh * ATT_HD + 64, k0, b, kEvictLast);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[s], TILE_BYTES);
        tma_load_3d(sV + s * TILE_BYTES, &tmV, &v_full[s], kvh * ATT_HD, k0, b, kEvictLast);
        tma_load_3d(sV + s * TILE_BYTES + CHUNK_BYTES, &tmV, &v_full[s], kvh * ATT_HD + 64, k0, b, kEvictLast);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);   // S = Q Kᵀ : both K-major
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, 0, 1);  // O += P V : V is MN-major (hd contiguous)
      const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
      auto issue_s = [&](int j) {
        const int s = j & 1;
        mbar_wait(&k_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + s * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < ATT_HD / 16; ++k) {
          const uint32_t off = (k >> 2) * CHUNK_BYTES + (k & 3) * 32;
          umma_ss(tmem_S + s * 128, make_sdesc_sw128(q_addr + off, 0, 1024), make_sdesc_sw128(k_addr + off, 0, 1024),
                  idesc_s, k > 0 ? 1u : 0u);
        }
        umma_commit(&k_empty[s]);
        umma_commit(&s_full[s]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n; ++j) {
        if (j + 1 < n) issue_s(j + 1);
        const int s = j & 1;
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sV + s * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < ATT_BLK / 16; ++k) {
          const uint64_t da = make_sdesc_sw128(p_addr + (k >> 2) * CHUNK_BYTES + (k & 3) * 32, 0, 1024);
          const uint64_t db = make_sdesc_sw128(v_addr + k * 2048, CHUNK_BYTES, 1024);
          umma_ss(tmem_O, da, db, idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&v_empty[s]);
        umma_commit(pv_done);
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax / epilogue (128 threads, thread = q row) =====================
    const uint32_t quad = warp & 3u;
    const uint32_t r = quad * 32 + lane;  // row inside the q block == TMEM lane
    const int tid = int(threadIdx.x) - 64;  // 0..127 (NOT the row; used for cooperative loads only)
    const int qpos = q0 + int(r);
    const int32_t* docb = p.doc + int64_t(b) * p.T;
    const int32_t dq = (qpos < p.T) ? docb[qpos] : 0;
    const int32_t dq_last = (q0 + ATT_BLK - 1 < p.T) ? docb[q0 + ATT_BLK - 1] : 0;  // uniform
    const uint32_t lane_sel = (quad * 32u) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    const float NEG_INF = -INFINITY;

    // canonical rows: keys of row q are exactly the positions [seg_start(q), q]  -> two integer compares per element;
    // non-canonical rows: exact element-wise document-id compare (ids staged in smem)
    const AttnSeg myseg = (qpos < p.T) ? p.seg[int64_t(b) * p.nblk * ATT_BLK + qpos] : AttnSeg{qpos + 1, qpos};
    const uint32_t sP_u32 = smem_u32(sP);
    const uint32_t sDocK_u32 = smem_u32(sDocK);

    for (int j = 0; j < n; ++j) {
      const int kb = kv_lo + j;
      const int k0 = kb * ATT_BLK;
      // block needs no mask iff strictly below the diagonal and one document spans [k0, q0+127]
      const bool full = meta.canonical && (kb < qb) && (dq_last > 0) && (docb[k0] == dq_last);
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t dk_u32 = sDocK_u32 + (j & 1) * 512;
      if (!meta.canonical) {
        // (S_j complete implies every thread finished block j-2, the previous user of this buffer)
        sts_u32(dk_u32 + tid * 4, uint32_t((k0 + tid < p.T) ? docb[k0 + tid] : -1));
        named_bar_sync(1, 128);
      }
      const uint32_t s_addr = tmem_S + (j & 1) * 128 + lane_sel;

      // ---- one TMEM pass: the whole 128-wide score row lives in registers ----
      uint32_t v[4][32];
      tmem_ld32(s_addr, v[0]);
      tmem_ld32(s_addr + 32, v[1]);
      tmem_ld32(s_addr + 64, v[2]);
      tmem_ld32(s_addr + 96, v[3]);
      tmem_ld_wait();
      float mx = NEG_INF;
      if (full) {
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4)
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[c4][i]));
      } else if (meta.canonical) {
        const int lo = myseg.start - k0, hi = qpos - k0;   // allowed columns: lo <= c <= hi
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int c = c4 * 32 + i;
            const float x = (c >= lo && c <= hi) ? __uint_as_float(v[c4][i]) : NEG_INF;
            v[c4][i] = __float_as_uint(x);
            mx = fmaxf(mx, x);
          }
      } else {
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4)
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            const uint4 d4 = lds_u4(dk_u32 + (c4 * 32 + i4 * 4) * 4);
            const int32_t dd[4] = {int32_t(d4.x), int32_t(d4.y), int32_t(d4.z), int32_t(d4.w)};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = c4 * 32 + i4 * 4 + e;
              const bool ok = (k0 + c <= qpos) && (dd[e] == dq) && (dq > 0);
              const float x = ok ? __uint_as_float(v[c4][i4 * 4 + e]) : NEG_INF;
              v[c4][i4 * 4 + e] = __float_as_uint(x);
              mx = fmaxf(mx, x);
            }
          }
      }
      mx *= p.scale_log2;  // scale > 0, so max commutes with the scaling (-inf stays -inf)
      // lazy rescale: keep the old reference max unless it grew by more than 2^8
      float m_new = m_run, alpha = 1.f;
      if (mx > m_run + 8.f || (m_run == NEG_INF && mx > NEG_INF)) {
        m_new = mx;
        alpha = fast_exp2(m_run - m_new);  // m_run = -inf -> 0
      }
      const float m_use = (m_new == NEG_INF) ? 0.f : m_new;

      // ---- p = exp2(s*scale - m) (masked entries are -inf -> 0), packed to bf16 in place ----
      float psum0 = 0.f, psum1 = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(v[c4][2 * i]), p.scale_log2, -m_use));
          const float p1 = fast_exp2(fmaf(__uint_as_float(v[c4][2 * i + 1]), p.scale_log2, -m_use));
          psum0 += p0;
          psum1 += p1;
          v[c4][i] = pack_bf16x2(p0, p1);   // slot i <= 2i: already consumed
        }
      l_run = l_run * alpha + (psum0 + psum1);
      m_run = m_new;

      // ---- wait for the previous PV (frees P smem and makes O consistent), rescale O if any row needs it ----
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll 1
          for (int c4 = 0; c4 < 4; ++c4) {
            uint32_t o[32];
            tmem_ld32(tmem_O + lane_sel + c4 * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tmem_O + lane_sel + c4 * 32, o);
          }
          tmem_st_wait();
        }
      }
      // ---- P -> smem (K-major, 128B swizzle): columns 32*c4 + [0,32) are packed words v[c4][0..15] ----
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int u = (c4 & 1) * 4 + q;   // 16-byte unit inside the 64-column chunk
          sts_u4(sP_u32 + (c4 >> 1) * CHUNK_BYTES + sw128_off(r, u),
                 make_uint4(v[c4][4 * q], v[c4][4 * q + 1], v[c4][4 * q + 2], v[c4][4 * q + 3]));
        }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }

    // ---- epilogue: O / l -> bf16 -> smem -> TMA store; lse ----
    mbar_wait(pv_done, (n - 1) & 1);
    tc_fence_after();
    const float inv_l = (l_run > 0.f) ? 1.f / l_run : 0.f;
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      uint32_t o[32];
      tmem_ld32(tmem_O + lane_sel + c4 * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          w[e] = pack_bf16x2(__uint_as_float(o[u * 8 + 2 * e]) * inv_l, __uint_as_float(o[u * 8 + 2 * e + 1]) * inv_l);
        const int col = c4 * 32 + u * 8;  // 0..127
        sts_u4(sP_u32 + (col >> 6) * CHUNK_BYTES + sw128_off(r, (col & 63) >> 3), make_uint4(w[0], w[1], w[2], w[3]));
      }
    }
    if (q0l + int(r) < p.Tq)
      p.lse[(int64_t(b) * p.H + h) * p.Tq + q0l + int(r)] =
          (l_run > 0.f) ? (m_run + log2f(l_run)) * 0.6931471805599453f : __int_as_float(0x7f800000);
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (tid == 0) {
      tma_store_3d(&tmO, sP, h * ATT_HD, q0l, b);
      tma_store_3d(&tmO, sP + CHUNK_BYTES, h * ATT_HD + 64, q0l, b);
      tma_store_commit();
      tma_store_wait_read<0>();
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

extern "C" int64_t tn_attn_meta_ints(int B, int T) {
  return attn_meta_total(B, (T + ATT_BLK - 1) / ATT_BLK);
}

extern "C" int tn_attn_prep(const int32_t* doc_ids, int32_t* meta, int B, int T, tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(doc_ids && meta, "tn_attn_prep: null pointer");
  TN_REQUIRE(B > 0 && T > 0, "tn_attn_prep: empty batch");
  const int nblk = (T + ATT_BLK - 1) / ATT_BLK;
  int32_t* flags = meta + attn_meta_flags_off(B, nblk);
  AttnSeg* seg = reinterpret_cast<AttnSeg*>(meta + attn_meta_seg_off(B, nblk));
  attn_canonical_kernel<<<B, 1024, 0, stream>>>(doc_ids, flags, T);
  TN_CHECK_CUDA(cudaGetLastError());
  const int warps = B * nblk;
  attn_meta_kernel<<<(warps + 7) / 8, 256, 0, stream>>>(doc_ids, flags, reinterpret_cast<AttnMeta*>(meta), seg, B, T, nblk);
  TN_CHECK_CUDA(cudaGetLastError());
  attn_order_kernel<<<1, 1024, 0, stream>>>(reinterpret_cast<const AttnMeta*>(meta), meta + attn_meta_order_off(B, nblk),
                                            meta + attn_meta_order_kv_off(B, nblk), B * nblk, nblk);
  TN_CHECK_CUDA(cudaGetLastError());
  return TN_OK;
}

int attn_fwd2_launch(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, void* O,
                     int64_t ldo, float* lse, const int32_t* doc_ids, const int32_t* meta, int B, int T, int H, int KV,
                     float scale, int Tq, int q_blk_off, cudaStream_t stream);

static bool use_fwd_v1() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TN_ATTN_FWD_V1"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

extern "C" int tn_attn_fwd_bf16(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv,
                                void* O, int64_t ldo, float* lse, const int32_t* doc_ids, const int32_t* meta, int B,
                                int T, int H, int KV, float scale, int Tq, int q_blk_off, tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(Q && K && V && O && lse && doc_ids && meta, "tn_attn_fwd_bf16: null pointer");
  TN_REQUIRE(B > 0 && T > 0 && H > 0 && KV > 0 && H % KV == 0, "tn_attn_fwd_bf16: bad dims B=%d T=%d H=%d KV=%d", B, T, H, KV);
  TN_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "tn_attn_fwd_bf16: strides must be multiples of 8");
  TN_REQUIRE(ldq >= int64_t(H) * ATT_HD && ldo >= int64_t(H) * ATT_HD && ldk >= int64_t(KV) * ATT_HD &&
                 ldv >= int64_t(KV) * ATT_HD, "tn_attn_fwd_bf16: token stride smaller than heads*128");
  const int nblk = (T + ATT_BLK - 1) / ATT_BLK;
  if (Tq <= 0) { Tq = T; q_blk_off = 0; }
  TN_REQUIRE(q_blk_off >= 0 && q_blk_off * ATT_BLK + Tq <= nblk * ATT_BLK, "tn_attn_fwd_bf16: query window outside the sequence");
  const int nqb = (Tq + ATT_BLK - 1) / ATT_BLK;
  if (!use_fwd_v1())
    return attn_fwd2_launch(Q, ldq, K, ldk, V, ldv, O, ldo, lse, doc_ids, meta, B, T, H, KV, scale, Tq, q_blk_off, stream);
  CUtensorMap tmQ, tmK, tmV, tmO;
  int rc;
  if ((rc = encode_tmap_3d(&tmQ, Q, 2, uint64_t(H) * ATT_HD, Tq, B, ldq * 2, uint64_t(Tq) * ldq * 2, 64, ATT_BLK, 1, true))) return rc;
  if ((rc = encode_tmap_3d(&tmK, K, 2, uint64_t(KV) * ATT_HD, T, B, ldk * 2, uint64_t(T) * ldk * 2, 64, ATT_BLK, 1, true))) return rc;
  if ((rc = encode_tmap_3d(&tmV, V, 2, uint64_t(KV) * ATT_HD, T, B, ldv * 2, uint64_t(T) * ldv * 2, 64, ATT_BLK, 1, true))) return rc;
  if ((rc = encode_tmap_3d(&tmO, O, 2, uint64_t(H) * ATT_HD, Tq, B, ldo * 2, uint64_t(Tq) * ldo * 2, 64, ATT_BLK, 1, true))) return rc;
  AttnFwdParams p{};
  p.doc = doc_ids; p.meta = reinterpret_cast<const AttnMeta*>(meta); p.lse = lse;
  p.seg = reinterpret_cast<const AttnSeg*>(meta + attn_meta_seg_off(B, nblk));
  p.O = static_cast<bf16*>(O); p.ldo = ldo;
  p.B = B; p.T = T; p.H = H; p.KV = KV; p.nblk = nblk;
  p.Tq = Tq; p.q_blk_off = q_blk_off;
  p.scale_log2 = scale * 1.4426950408889634f;
  static bool configured = false;
  if (!configured) {
    TN_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::ALLOC));
    configured = true;
  }
  dim3 grid(nqb, H, B);
  attn_fwd_kernel<<<grid, FWD_THREADS, FwdSmem::ALLOC, stream>>>(tmQ, tmK, tmV, tmO, p);
  TN_CHECK_CUDA(cudaGetLastError());
  return TN_OK;
}
