// touchnet_b200 :: CTA-pair tcgen05 GEMM (cta_group::2).  256 x 256 x 64 tiles computed by two CTAs of one cluster:
// each CTA stages 128 rows of A and 128 rows of B per k-block (32 KB instead of 48 KB for the same MMA work), the
// leader CTA issues one tcgen05.mma.cta_group::2 (M=256) that reads B from both CTAs' shared memory, each CTA's TMEM
// holds its 128 accumulator rows (2 x 256 columns, double buffered), and the epilogue goes TMEM -> registers ->
// 128B-swizzled smem staging -> TMA store (fully coalesced, clipped at the tensor edges by the hardware).
//
// Same call sites as gemm.cu (F.linear forward / dgrad / wgrad of hf:models/llama/modeling_llama.py:251-289,
// 182-184); tn_gemm_bf16 / tn_gemm_swiglu_bf16 dispatch here for M,N >= 256.
#include "../../include/touchnet_b200.h"
#include "common.cuh"
#include "host.h"

namespace tn {

constexpr int P_BM = 128;        // rows of A (and of D) per CTA
constexpr int P_BN = 256;        // D columns per cluster tile; each CTA stages P_BN/2 rows of B
constexpr int P_BK = 64;
constexpr int P_THREADS = 256;
constexpr int P_A_BYTES = P_BM * P_BK * 2;        // 16 KB
constexpr int P_B_BYTES = (P_BN / 2) * P_BK * 2;  // 16 KB
constexpr int P_STAGE_BYTES = P_A_BYTES + P_B_BYTES;
constexpr int P_STG_BYTES = 128 * 128;            // one staging buffer: 128 rows x 128 B

template <int EPI>
struct PairCfg {
  static constexpr int STAGES = (EPI == 1) ? 5 : 6;
  static constexpr int NSTG = (EPI == 1) ? 3 : 2;   // swiglu: G, U, H buffers; else double buffer
  static constexpr int STG_OFF = STAGES * P_STAGE_BYTES;
  static constexpr int BAR_OFF = STG_OFF + NSTG * P_STG_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;
};

struct PairParams {
  int M, N, K;
  int num_m, num_n, num_k;   // num_m in units of 256 rows
  int group;                 // raster: tiles walk `group` M blocks x all N blocks before moving down (L2 reuse of A)
  // L2 policy of the operand loads: inside a raster group the `group` A strips are re-read by every wave, a B strip is
  // consumed within about one wave and not touched again until the next group -> A evict-last, B evict-first keeps the
  // strips that will be re-read resident in the 126 MB L2 (tn_set_gemm_l2_hints, default on)
  uint64_t hint_a, hint_b;
  int split_tail;            // 1: half tiles in the last wave (pair_work); 0 for A/B runs (tn_set_gemm_split_tail)
  const void* R;
  int64_t ldr;
  // segmented operands: several weight tensors that are separate nn.Parameters (q/k/v projections) behave as one GEMM
  //   b_seg 1: B = [B0;B1;B2] stacked along N (fused QKV forward)       -> D columns [0,off1) [off1,off2) [off2,N)
  //   b_seg 2: B = [B0;B1;B2] stacked along K (fused QKV dgrad)         -> A columns ...
  //   d_seg 1: D = [D0;D1;D2] stacked along M (fused QKV wgrad: one dW per parameter)
  int b_seg, d_seg;
  int off1, off2;
  // fused RoPE in the epilogue of the QKV forward (hf apply_rotary_pos_emb, modeling_llama.py:151-168): output columns
  // below rope_end (the q and k segments, heads of 128 columns) are rotated with the per-row cos/sin tables
  // [M, 64] bf16 (tn_rope_table); same bf16 rounding points as the unfused kernel -> bit-identical results
  const bf16* rope_cos;
  const bf16* rope_sin;
  int rope_end;
  // EPI 3 (down-proj dgrad fused with the SwiGLU backward, hf LlamaMLP.forward modeling_llama.py:182-184): the accumulator
  // tile is dH; the epilogue reads the matching G / U tiles ([M, N] bf16, row stride ld_gu) and stores dG (tmD), dU (tmD2)
  const bf16* aux_g;
  const bf16* aux_u;
  int64_t ld_gu;
};

__device__ __forceinline__ void pair_decode_tile(int tile, int num_m, int num_n, int group, int& m_blk, int& n_blk) {
  const int per_group = group * num_n;
  const int g = tile / per_group;
  const int first_m = g * group;
  const int gsize = min(num_m - first_m, group);
  const int r = tile - g * per_group;
  m_blk = first_m + (r % gsize);
  n_blk = r / gsize;
}

// Work list of one launch: the tiles of all full waves, then - when the last wave would leave more than half of the
// clusters idle - the remaining tiles as 256 x 128 HALF tiles (2x as many items, each half as long), so the tail costs
// half a wave instead of a whole one (o_proj wgrad: 256 tiles on 74 clusters = 3.46 waves -> 3 + 0.5 instead of 4).
// A half tile loads the same operand boxes (the B box of CTA r simply starts 64 rows later and only its first 64 rows
// are used), issues N = 128 MMAs into the first 128 accumulator columns and stores 128 columns.
struct PairWork { int tile; int half; };   // half: -1 = full tile, 0 / 1 = left / right 128 columns
__device__ __forceinline__ int pair_num_work(int num_tiles, int C, bool allow) {
  const int full_w = (num_tiles / C) * C, R = num_tiles - full_w;
  return (allow && R > 0 && 2 * R <= C) ? full_w + 2 * R : num_tiles;
}
__device__ __forceinline__ PairWork pair_work(int w, int num_tiles, int C, bool allow) {
  const int full_w = (num_tiles / C) * C, R = num_tiles - full_w;
  if (!(allow && R > 0 && 2 * R <= C) || w < full_w) return PairWork{w, -1};
  return PairWork{full_w + ((w - full_w) >> 1), (w - full_w) & 1};
}

__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// EPI: 0 = bf16 D = acc (+R);  1 = SwiGLU (cols [0,128) gate, [128,256) up -> G,U,H);  2 = fp32 D = acc (+R);
//      3 = SwiGLU backward on the dH accumulator: dG = (dH*U) * silu'(G), dU = dH * silu(G)  (same rounding points as
//          tn_swiglu_bwd_bf16 applied to the bf16-rounded dH, so the fused and the two-kernel paths agree bit for bit)
template <bool A_MN, bool B_MN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmB3,
                 const __grid_constant__ CUtensorMap tmD,
                 const __grid_constant__ CUtensorMap tmD2, const __grid_constant__ CUtensorMap tmD3, const PairParams p) {
  using Cfg = PairCfg<EPI>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * P_A_BYTES;
  uint8_t* sStg = smem + Cfg::STG_OFF;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const uint32_t warp = warp_id(), lane = lane_id();
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmD);
    if (p.b_seg) { tma_prefetch_desc(&tmB2); tma_prefetch_desc(&tmB3); }
    if (p.d_seg) { tma_prefetch_desc(&tmD2); tma_prefetch_desc(&tmD3); }
    if (EPI == 1) { tma_prefetch_desc(&tmB2); tma_prefetch_desc(&tmD2); tma_prefetch_desc(&tmD3); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 8); }  // 4 warps x 2 CTAs
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // peer's barriers are initialised before anything remote touches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int num_tiles = p.num_m * p.num_n;
  const bool split_ok = (EPI != 1) && p.split_tail != 0;
  const int num_work = pair_num_work(num_tiles, num_clusters, split_ok);

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; transaction bytes counted on the leader's barrier) ============
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cluster_id; w < num_work; w += num_clusters) {
        const PairWork wk = pair_work(w, num_tiles, num_clusters, split_ok);
        int m_blk, n_blk;
        pair_decode_tile(wk.tile, p.num_m, p.num_n, p.group, m_blk, n_blk);
        const int m0 = m_blk * 256 + int(rank) * P_BM;
        int n0 = (EPI == 1) ? n_blk * 128
                            : (wk.half < 0 ? n_blk * P_BN + int(rank) * (P_BN / 2)
                                           : n_blk * P_BN + wk.half * (P_BN / 2) + int(rank) * (P_BN / 4));
        const CUtensorMap* bmap = &tmB;
        if (EPI != 1 && p.b_seg == 1) {   // weight segment that owns these 128 output columns
          const int sg = (n0 >= p.off1) + (n0 >= p.off2);
          bmap = sg == 0 ? &tmB : (sg == 1 ? &tmB2 : &tmB3);
          n0 -= sg == 0 ? 0 : (sg == 1 ? p.off1 : p.off2);
        }
        for (int kb = 0; kb < p.num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * P_STAGE_BYTES);
          uint8_t* a_dst = sA + stage * P_A_BYTES;
          uint8_t* b_dst = sB + stage * P_B_BYTES;
          const int k0 = kb * P_BK;
          int kb0 = k0;   // k coordinate inside the selected B segment
          if (EPI != 1 && p.b_seg == 2) {
            const int sg = (k0 >= p.off1) + (k0 >= p.off2);
            bmap = sg == 0 ? &tmB : (sg == 1 ? &tmB2 : &tmB3);
            kb0 = k0 - (sg == 0 ? 0 : (sg == 1 ? p.off1 : p.off2));
          }
          if (!A_MN) {
            tma_load_2d_pair(a_dst, &tmA, &full_bar[stage], k0, m0, p.hint_a);
          } else {
            tma_load_2d_pair(a_dst, &tmA, &full_bar[stage], m0, k0, p.hint_a);
            tma_load_2d_pair(a_dst + 8192, &tmA, &full_bar[stage], m0 + 64, k0, p.hint_a);
          }
          if (EPI == 1) {
            tma_load_2d_pair(b_dst, rank == 0 ? &tmB : &tmB2, &full_bar[stage], k0, n0, p.hint_b);  // CTA0: gate rows, CTA1: up rows
          } else if (!B_MN) {
            tma_load_2d_pair(b_dst, bmap, &full_bar[stage], kb0, n0, p.hint_b);
          } else {
            tma_load_2d_pair(b_dst, bmap, &full_bar[stage], n0, kb0, p.hint_b);
            tma_load_2d_pair(b_dst + 8192, bmap, &full_bar[stage], n0 + 64, kb0, p.hint_b);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc_full = make_idesc_bf16(256, P_BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      constexpr uint32_t idesc_half = make_idesc_bf16(256, P_BN / 2, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = cluster_id; w < num_work; w += num_clusters) {
        const uint32_t idesc = pair_work(w, num_tiles, num_clusters, split_ok).half < 0 ? idesc_full : idesc_half;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * P_BN);
        for (int kb = 0; kb < p.num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * P_A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * P_B_BYTES);
#pragma unroll
          for (int k = 0; k < P_BK / 16; ++k) {
            const uint64_t da = A_MN ? make_sdesc_sw128(a_addr + k * 2048, 8192, 1024)
                                     : make_sdesc_sw128(a_addr + k * 32, 0, 1024);
            const uint64_t db = B_MN ? make_sdesc_sw128(b_addr + k * 2048, 8192, 1024)
                                     : make_sdesc_sw128(b_addr + k * 32, 0, 1024);
            umma_ss_pair(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[stage], 0b11);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tfull_bar[acc], 0b11);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs): TMEM -> regs -> swizzled smem -> TMA store =====================
    const uint32_t quad = warp & 3u;
    const uint32_t r = quad * 32 + lane;         // row inside this CTA's 128-row slab == TMEM lane
    const int etid = int(threadIdx.x) - 128;     // 0..127
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t n_issued = 0;                       // staging rounds issued so far (buffer ring position)
    for (int w = cluster_id; w < num_work; w += num_clusters) {
      const PairWork wk = pair_work(w, num_tiles, num_clusters, split_ok);
      int m_blk, n_blk;
      pair_decode_tile(wk.tile, p.num_m, p.num_n, p.group, m_blk, n_blk);
      const int ncols = wk.half < 0 ? P_BN : P_BN / 2;          // accumulator columns of this work item
      const int row0 = m_blk * 256 + int(rank) * P_BM;
      const int row = row0 + int(r);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + uint32_t(acc * P_BN) + ((quad * 32u) << 16);

      if (EPI == 1) {
        const int n0 = n_blk * 128;
#pragma unroll 1
        for (int c = 0; c < 128; c += 64) {
          if (n_issued > 0) {
            if (etid == 0) tma_store_wait_read<0>();
            named_bar(2, 128);
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t g[32], u[32];
            tmem_ld32(t_row + c + half * 32, g);
            tmem_ld32(t_row + 128 + c + half * 32, u);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t pg[4], pu[4], ph[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int i = q * 8 + j * 2;
                const float g0 = bf16_round(__uint_as_float(g[i])), g1 = bf16_round(__uint_as_float(g[i + 1]));
                const float u0 = bf16_round(__uint_as_float(u[i])), u1 = bf16_round(__uint_as_float(u[i + 1]));
                const float s0 = bf16_round(g0 / (1.f + __expf(-g0))), s1 = bf16_round(g1 / (1.f + __expf(-g1)));
                pg[j] = pack_bf16x2(g0, g1);
                pu[j] = pack_bf16x2(u0, u1);
                ph[j] = pack_bf16x2(s0 * u0, s1 * u1);
              }
              const uint32_t off = r * 128u + (((uint32_t(half * 4 + q)) ^ (r & 7u)) << 4);
              *reinterpret_cast<uint4*>(sStg + off) = make_uint4(pg[0], pg[1], pg[2], pg[3]);
              *reinterpret_cast<uint4*>(sStg + P_STG_BYTES + off) = make_uint4(pu[0], pu[1], pu[2], pu[3]);
              *reinterpret_cast<uint4*>(sStg + 2 * P_STG_BYTES + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
            }
          }
          fence_proxy_async_smem();
          named_bar(2, 128);
          if (etid == 0) {
            tma_store_2d(&tmD2, sStg, n0 + c, row0);                    // G
            tma_store_2d(&tmD3, sStg + P_STG_BYTES, n0 + c, row0);      // U
            tma_store_2d(&tmD, sStg + 2 * P_STG_BYTES, n0 + c, row0);   // H
            tma_store_commit();
          }
          ++n_issued;
        }
      } else {
        constexpr int CW = (EPI == 2) ? 32 : 64;   // columns per staging round (128 B per row)
        const int n0 = n_blk * P_BN + (wk.half > 0 ? P_BN / 2 : 0);
        if (EPI == 3) {
          const bool row_ok = row < p.M;
          const bf16* grow = p.aux_g + int64_t(row) * p.ld_gu;
          const bf16* urow = p.aux_u + int64_t(row) * p.ld_gu;
#pragma unroll 1
          for (int c = 0; c < ncols; c += 64) {
            if (n_issued > 0) {
              if (etid == 0) tma_store_wait_read<0>();
              named_bar(2, 128);
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t v[32];
              tmem_ld32(t_row + c + half * 32, v);
              uint4 g4[4], u4[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {                       // this row's G / U values of the 32 columns: issued with the TMEM read
                const int col = n0 + c + half * 32 + q * 8;
                const bool ok = row_ok && col < p.N;
                g4[q] = ok ? *reinterpret_cast<const uint4*>(grow + col) : make_uint4(0, 0, 0, 0);
                u4[q] = ok ? *reinterpret_cast<const uint4*>(urow + col) : make_uint4(0, 0, 0, 0);
              }
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint32_t gw[4] = {g4[q].x, g4[q].y, g4[q].z, g4[q].w}, uw[4] = {u4[q].x, u4[q].y, u4[q].z, u4[q].w};
                uint32_t og[4], ou[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float rg[2], ru[2];
#pragma unroll
                  for (int e = 0; e < 2; ++e) {
                    const float g = e ? bf16hi(gw[j]) : bf16lo(gw[j]);
                    const float u = e ? bf16hi(uw[j]) : bf16lo(uw[j]);
                    const float dh = bf16_round(__uint_as_float(v[q * 8 + j * 2 + e]));   // the bf16 dH the unfused path stores
                    const float sig = 1.f / (1.f + __expf(-g));
                    const float silu = g * sig;
                    ru[e] = dh * bf16_round(silu);
                    rg[e] = bf16_round(dh * u) * (sig * (1.f + g * (1.f - sig)));
                  }
                  og[j] = pack_bf16x2(rg[0], rg[1]);
                  ou[j] = pack_bf16x2(ru[0], ru[1]);
                }
                const uint32_t off = r * 128u + ((uint32_t(half * 4 + q) ^ (r & 7u)) << 4);
                *reinterpret_cast<uint4*>(sStg + off) = make_uint4(og[0], og[1], og[2], og[3]);
                *reinterpret_cast<uint4*>(sStg + P_STG_BYTES + off) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
              }
            }
            fence_proxy_async_smem();
            named_bar(2, 128);
            if (etid == 0) {
              tma_store_2d(&tmD, sStg, n0 + c, row0);                    // dG
              tma_store_2d(&tmD2, sStg + P_STG_BYTES, n0 + c, row0);     // dU
              tma_store_commit();
            }
            n_issued += 2;
          }
        } else if (EPI == 0 && p.rope_cos != nullptr && n0 < p.rope_end) {
          // ---- RoPE tile: 2 heads of 128 columns; pairs (j, j+64) sit in this thread's row ----
          const bool row_ok = row < p.M;
#pragma unroll 1
          for (int head = 0; head < ncols / 128; ++head) {
            if (n_issued > 0) {
              if (etid == 0) tma_store_wait_read<0>();
              named_bar(2, 128);
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t lo[32], hi[32];
              tmem_ld32(t_row + head * 128 + half * 32, lo);
              tmem_ld32(t_row + head * 128 + 64 + half * 32, hi);
              tmem_ld_wait();
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                uint4 c4v = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u), s4v = make_uint4(0, 0, 0, 0);
                if (row_ok) {
                  c4v = *reinterpret_cast<const uint4*>(p.rope_cos + int64_t(row) * 64 + half * 32 + u * 8);
                  s4v = *reinterpret_cast<const uint4*>(p.rope_sin + int64_t(row) * 64 + half * 32 + u * 8);
                }
                const uint32_t cw[4] = {c4v.x, c4v.y, c4v.z, c4v.w}, sw[4] = {s4v.x, s4v.y, s4v.z, s4v.w};
                uint32_t wl[4], wh[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float a[2] = {bf16_round(__uint_as_float(lo[u * 8 + 2 * e])), bf16_round(__uint_as_float(lo[u * 8 + 2 * e + 1]))};
                  float bb[2] = {bf16_round(__uint_as_float(hi[u * 8 + 2 * e])), bf16_round(__uint_as_float(hi[u * 8 + 2 * e + 1]))};
                  const float c[2] = {bf16lo(cw[e]), bf16hi(cw[e])}, sn[2] = {bf16lo(sw[e]), bf16hi(sw[e])};
                  float o1[2], o2[2];
#pragma unroll
                  for (int t2 = 0; t2 < 2; ++t2) {
                    o1[t2] = bf16_round(a[t2] * c[t2]) + bf16_round(-bb[t2] * sn[t2]);
                    o2[t2] = bf16_round(bb[t2] * c[t2]) + bf16_round(a[t2] * sn[t2]);
                  }
                  wl[e] = pack_bf16x2(o1[0], o1[1]);
                  wh[e] = pack_bf16x2(o2[0], o2[1]);
                }
                const uint32_t off = r * 128u + ((uint32_t(half * 4 + u) ^ (r & 7u)) << 4);
                *reinterpret_cast<uint4*>(sStg + off) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
                *reinterpret_cast<uint4*>(sStg + P_STG_BYTES + off) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
              }
            }
            fence_proxy_async_smem();
            named_bar(2, 128);
            if (etid == 0) {
              tma_store_2d(&tmD, sStg, n0 + head * 128, row0);
              tma_store_2d(&tmD, sStg + P_STG_BYTES, n0 + head * 128 + 64, row0);
              tma_store_commit();
            }
            n_issued += 2;
          }
        } else {
        const CUtensorMap* dmap = &tmD;
        int drow0 = row0;
        if (p.d_seg == 1) {                       // output rows belong to one of several gradient tensors
          const int sg = (row0 >= p.off1) + (row0 >= p.off2);
          dmap = sg == 0 ? &tmD : (sg == 1 ? &tmD2 : &tmD3);
          drow0 = row0 - (sg == 0 ? 0 : (sg == 1 ? p.off1 : p.off2));
        }
#pragma unroll 1
        for (int c = 0; c < ncols; c += CW) {
          uint8_t* stg = sStg + (n_issued & 1u) * P_STG_BYTES;
          if (n_issued >= 2) {
            if (etid == 0) {
              // the store that last read this buffer has drained it (RoPE tiles commit both buffers in one group)
              if (EPI == 0 && p.rope_cos != nullptr) tma_store_wait_read<0>(); else tma_store_wait_read<1>();
            }
            named_bar(2, 128);
          }
          const int col = n0 + c;
          if (EPI == 2) {
            uint32_t v[32];
            tmem_ld32(t_row + c, v);
            tmem_ld_wait();
            if (p.R && row < p.M && col < p.N) {
              const float* rptr = reinterpret_cast<const float*>(p.R) + int64_t(row) * p.ldr + col;
              if (col + 32 <= p.N) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const float4 r4 = reinterpret_cast<const float4*>(rptr)[q];
                  v[4 * q] = __float_as_uint(__uint_as_float(v[4 * q]) + r4.x);
                  v[4 * q + 1] = __float_as_uint(__uint_as_float(v[4 * q + 1]) + r4.y);
                  v[4 * q + 2] = __float_as_uint(__uint_as_float(v[4 * q + 2]) + r4.z);
                  v[4 * q + 3] = __float_as_uint(__uint_as_float(v[4 * q + 3]) + r4.w);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (col + i < p.N) v[i] = __float_as_uint(__uint_as_float(v[i]) + rptr[i]);
              }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<uint4*>(stg + r * 128u + ((uint32_t(q) ^ (r & 7u)) << 4)) =
                  make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          } else {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t v[32];
              tmem_ld32(t_row + c + half * 32, v);
              tmem_ld_wait();
              const int hcol = col + half * 32;
              const bool has_r = p.R && row < p.M && hcol < p.N;
              const bf16* rptr = has_r ? reinterpret_cast<const bf16*>(p.R) + int64_t(row) * p.ldr + hcol : nullptr;
              const bool vec_r = has_r && (hcol + 32 <= p.N);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 r4 = make_uint4(0, 0, 0, 0);
                if (vec_r) r4 = reinterpret_cast<const uint4*>(rptr)[q];
                const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float lo = __uint_as_float(v[8 * q + 2 * j]), hi = __uint_as_float(v[8 * q + 2 * j + 1]);
                  if (vec_r) {
                    lo = bf16_round(lo) + bf16lo(rw[j]);
                    hi = bf16_round(hi) + bf16hi(rw[j]);
                  } else if (has_r) {
                    const int i0 = 8 * q + 2 * j;
                    if (hcol + i0 < p.N) lo = bf16_round(lo) + __bfloat162float(rptr[i0]);
                    if (hcol + i0 + 1 < p.N) hi = bf16_round(hi) + __bfloat162float(rptr[i0 + 1]);
                  }
                  w[j] = pack_bf16x2(lo, hi);
                }
                *reinterpret_cast<uint4*>(stg + r * 128u + ((uint32_t(half * 4 + q) ^ (r & 7u)) << 4)) =
                    make_uint4(w[0], w[1], w[2], w[3]);
              }
            }
          }
          fence_proxy_async_smem();
          named_bar(2, 128);
          if (etid == 0) {
            tma_store_2d(dmap, stg, col, drow0);
            tma_store_commit();
          }
          ++n_issued;
        }
        }   // (non-RoPE tile)
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (etid == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // nobody leaves while the peer may still read our smem / arrive on our barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair<512>(tmem_base);
  }
}

template <bool A_MN, bool B_MN, int EPI>
static int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmB2, const CUtensorMap& tmB3,
                       const CUtensorMap& tmD, const CUtensorMap& tmD2, const CUtensorMap& tmD3, const PairParams& p,
                       cudaStream_t stream) {
  using Cfg = PairCfg<EPI>;
  auto kern = gemm_pair_kernel<A_MN, B_MN, EPI>;
  static bool configured = false;
  if (!configured) {
    TN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int tiles = p.num_m * p.num_n;
  const int max_clusters = sm_count() / 2;
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  kern<<<clusters * 2, P_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, tmB2, tmB3, tmD, tmD2, tmD3, p);
  TN_CHECK_CUDA(cudaGetLastError());
  return TN_OK;
}

// ---- entry points used by gemm.cu's dispatcher ----
int gemm_pair_dispatch(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* D, int64_t ldd,
                       int d_f32, const void* R, int64_t ldr, int M, int N, int K, cudaStream_t stream) {
  PairParams p{};
  p.M = M; p.N = N; p.K = K;
  p.num_m = (M + 255) / 256; p.num_n = (N + P_BN - 1) / P_BN; p.num_k = (K + P_BK - 1) / P_BK;
  p.group = gemm_group();
  p.split_tail = gemm_split_tail();
  p.hint_a = gemm_l2_hints() ? kEvictLast : kEvictNormal;
  p.hint_b = gemm_l2_hints() ? kEvictFirst : kEvictNormal;
  p.R = R; p.ldr = ldr;
  CUtensorMap tmA, tmB, tmD;
  int rc;
  if (!a_mn) rc = encode_tmap_2d(&tmA, A, 2, uint64_t(K), uint64_t(M), uint64_t(lda) * 2, 64, P_BM, true);
  else       rc = encode_tmap_2d(&tmA, A, 2, uint64_t(M), uint64_t(K), uint64_t(lda) * 2, 64, 64, true);
  if (rc) return rc;
  if (!b_mn) rc = encode_tmap_2d(&tmB, B, 2, uint64_t(K), uint64_t(N), uint64_t(ldb) * 2, 64, P_BN / 2, true);
  else       rc = encode_tmap_2d(&tmB, B, 2, uint64_t(N), uint64_t(K), uint64_t(ldb) * 2, 64, 64, true);
  if (rc) return rc;
  if (d_f32) rc = encode_tmap_2d(&tmD, D, 4, uint64_t(N), uint64_t(M), uint64_t(ldd) * 4, 32, 128, true);
  else       rc = encode_tmap_2d(&tmD, D, 2, uint64_t(N), uint64_t(M), uint64_t(ldd) * 2, 64, 128, true);
  if (rc) return rc;
#define TN_PAIR(AMN_, BMN_)                                                                   \
  (d_f32 ? launch_pair<AMN_, BMN_, 2>(tmA, tmB, tmB, tmB, tmD, tmD, tmD, p, stream)           \
         : launch_pair<AMN_, BMN_, 0>(tmA, tmB, tmB, tmB, tmD, tmD, tmD, p, stream))
  if (!a_mn && !b_mn) return TN_PAIR(false, false);
  if (!a_mn && b_mn) return TN_PAIR(false, true);
  return TN_PAIR(true, true);
#undef TN_PAIR
}

// Down-proj dgrad with the SwiGLU backward in its epilogue: dH = dY[M,K].W[K,N] (W = down_proj.weight [d, ffn], MN-major B),
// dG / dU [M,N] from the accumulator and the saved G / U.  Caller guarantees M, N >= 256 (gemm.cu falls back otherwise).
int gemm_pair_dswiglu_dispatch(const void* dY, int64_t lddy, const void* W, int64_t ldw, const void* G, const void* U,
                               int64_t ldgu, void* dG, void* dU, int64_t lddg, int M, int N, int K, cudaStream_t stream) {
  PairParams p{};
  p.M = M; p.N = N; p.K = K;
  p.num_m = (M + 255) / 256; p.num_n = (N + P_BN - 1) / P_BN; p.num_k = (K + P_BK - 1) / P_BK;
  p.group = gemm_group();
  p.split_tail = gemm_split_tail();
  p.hint_a = gemm_l2_hints() ? kEvictLast : kEvictNormal;
  p.hint_b = gemm_l2_hints() ? kEvictFirst : kEvictNormal;
  p.aux_g = static_cast<const bf16*>(G); p.aux_u = static_cast<const bf16*>(U); p.ld_gu = ldgu;
  CUtensorMap tmA, tmB, tmD, tmD2;
  int rc;
  if ((rc = encode_tmap_2d(&tmA, dY, 2, uint64_t(K), uint64_t(M), uint64_t(lddy) * 2, 64, P_BM, true))) return rc;
  if ((rc = encode_tmap_2d(&tmB, W, 2, uint64_t(N), uint64_t(K), uint64_t(ldw) * 2, 64, 64, true))) return rc;
  if ((rc = encode_tmap_2d(&tmD, dG, 2, uint64_t(N), uint64_t(M), uint64_t(lddg) * 2, 64, 128, true))) return rc;
  if ((rc = encode_tmap_2d(&tmD2, dU, 2, uint64_t(N), uint64_t(M), uint64_t(lddg) * 2, 64, 128, true))) return rc;
  return launch_pair<false, true, 3>(tmA, tmB, tmB, tmB, tmD, tmD2, tmD, p, stream);
}

// Three weights / gradients treated as one operand (see PairParams).  mode: 0 = forward  D[M, n0+n1+n2] = A·[B0;B1;B2]ᵀ
// (B K-major), 1 = dgrad  D[M,N] = [A0|A1|A2]·[B0;B1;B2] (A = one [M, k0+k1+k2] buffer, B MN-major, segmented along K),
// 2 = wgrad  [D0;D1;D2] = Aᵀ·X (A = one [Mred, m0+m1+m2] buffer MN-major, B MN-major, D segmented along M).
int gemm_pair_seg_dispatch(int mode, const void* A, int64_t lda, const void* const* Bs, int64_t ldb, void* const* Ds,
                           int64_t ldd, int d_f32, const int* seg, int M, int N, int K, const void* rope_cos,
                           const void* rope_sin, cudaStream_t stream) {
  PairParams p{};
  if (mode == 0 && rope_cos && rope_sin) {
    p.rope_cos = static_cast<const bf16*>(rope_cos); p.rope_sin = static_cast<const bf16*>(rope_sin);
    p.rope_end = seg[0] + seg[1];
  }
  p.M = M; p.N = N; p.K = K;
  p.num_m = (M + 255) / 256; p.num_n = (N + P_BN - 1) / P_BN; p.num_k = (K + P_BK - 1) / P_BK;
  p.group = gemm_group();
  p.split_tail = gemm_split_tail();
  p.hint_a = gemm_l2_hints() ? kEvictLast : kEvictNormal;
  p.hint_b = gemm_l2_hints() ? kEvictFirst : kEvictNormal;
  p.off1 = seg[0]; p.off2 = seg[0] + seg[1];
  CUtensorMap tmA, tmB[3], tmD[3];
  int rc;
  if (mode == 0) {
    p.b_seg = 1;
    if ((rc = encode_tmap_2d(&tmA, A, 2, uint64_t(K), uint64_t(M), uint64_t(lda) * 2, 64, P_BM, true))) return rc;
    for (int i = 0; i < 3; ++i)
      if ((rc = encode_tmap_2d(&tmB[i], Bs[i], 2, uint64_t(K), uint64_t(seg[i]), uint64_t(ldb) * 2, 64, P_BN / 2, true))) return rc;
    if ((rc = encode_tmap_2d(&tmD[0], Ds[0], 2, uint64_t(N), uint64_t(M), uint64_t(ldd) * 2, 64, 128, true))) return rc;
    return launch_pair<false, false, 0>(tmA, tmB[0], tmB[1], tmB[2], tmD[0], tmD[0], tmD[0], p, stream);
  }
  if (mode == 1) {
    p.b_seg = 2;
    if ((rc = encode_tmap_2d(&tmA, A, 2, uint64_t(K), uint64_t(M), uint64_t(lda) * 2, 64, P_BM, true))) return rc;
    for (int i = 0; i < 3; ++i)   // B_i stored [seg_i (K), N] row-major, N contiguous
      if ((rc = encode_tmap_2d(&tmB[i], Bs[i], 2, uint64_t(N), uint64_t(seg[i]), uint64_t(ldb) * 2, 64, 64, true))) return rc;
    if ((rc = encode_tmap_2d(&tmD[0], Ds[0], 2, uint64_t(N), uint64_t(M), uint64_t(ldd) * 2, 64, 128, true))) return rc;
    return launch_pair<false, true, 0>(tmA, tmB[0], tmB[1], tmB[2], tmD[0], tmD[0], tmD[0], p, stream);
  }
  p.d_seg = 1;
  if ((rc = encode_tmap_2d(&tmA, A, 2, uint64_t(M), uint64_t(K), uint64_t(lda) * 2, 64, 64, true))) return rc;
  if ((rc = encode_tmap_2d(&tmB[0], Bs[0], 2, uint64_t(N), uint64_t(K), uint64_t(ldb) * 2, 64, 64, true))) return rc;
  for (int i = 0; i < 3; ++i) {
    if (d_f32) rc = encode_tmap_2d(&tmD[i], Ds[i], 4, uint64_t(N), uint64_t(seg[i]), uint64_t(ldd) * 4, 32, 128, true);
    else       rc = encode_tmap_2d(&tmD[i], Ds[i], 2, uint64_t(N), uint64_t(seg[i]), uint64_t(ldd) * 2, 64, 128, true);
    if (rc) return rc;
  }
  return d_f32 ? launch_pair<true, true, 2>(tmA, tmB[0], tmB[0], tmB[0], tmD[0], tmD[1], tmD[2], p, stream)
               : launch_pair<true, true, 0>(tmA, tmB[0], tmB[0], tmB[0], tmD[0], tmD[1], tmD[2], p, stream);
}

int gemm_pair_swiglu_dispatch(const void* X, int64_t ldx, const void* Wg, const void* Wu, int64_t ldw, void* G, void* U,
                              void* H, int64_t ldh, int M, int N, int K, cudaStream_t stream) {
  PairParams p{};
  p.M = M; p.N = N; p.K = K;
  p.num_m = (M + 255) / 256; p.num_n = (N + 127) / 128; p.num_k = (K + P_BK - 1) / P_BK;
  p.group = gemm_group();
  p.split_tail = gemm_split_tail();
  p.hint_a = gemm_l2_hints() ? kEvictLast : kEvictNormal;
  p.hint_b = gemm_l2_hints() ? kEvictFirst : kEvictNormal;
  CUtensorMap tmA, tmG, tmU, tmDG, tmDU, tmDH;
  int rc;
  if ((rc = encode_tmap_2d(&tmA, X, 2, uint64_t(K), uint64_t(M), uint64_t(ldx) * 2, 64, P_BM, true))) return rc;
  if ((rc = encode_tmap_2d(&tmG, Wg, 2, uint64_t(K), uint64_t(N), uint64_t(ldw) * 2, 64, 128, true))) return rc;
  if ((rc = encode_tmap_2d(&tmU, Wu, 2, uint64_t(K), uint64_t(N), uint64_t(ldw) * 2, 64, 128, true))) return rc;
  if ((rc = encode_tmap_2d(&tmDG, G, 2, uint64_t(N), uint64_t(M), uint64_t(ldh) * 2, 64, 128, true))) return rc;
  if ((rc = encode_tmap_2d(&tmDU, U, 2, uint64_t(N), uint64_t(M), uint64_t(ldh) * 2, 64, 128, true))) return rc;
  if ((rc = encode_tmap_2d(&tmDH, H, 2, uint64_t(N), uint64_t(M), uint64_t(ldh) * 2, 64, 128, true))) return rc;
  return launch_pair<false, false, 1>(tmA, tmG, tmU, tmU, tmDH, tmDG, tmDU, p, stream);
}

}  // namespace tn
