// touchnet_b200 :: tcgen05 GEMM for the projection / MLP matmuls of the packed decoder layer.
//
// Replaces the cuBLASLt calls behind F.linear at
//   hf:models/llama/modeling_llama.py:251-289 (q/k/v/o_proj), :182-184 (gate/up/down_proj)
//   touchnet/models/touch_audio/modeling_touch_audio.py:127 (projector)
// and their autograd backward (dgrad, wgrad).
//
// Design (B200-first, one CTA per SM, persistent):
//   warp 0      : TMA producer  — cp.async.bulk.tensor tiles (128B swizzle) into a STAGES-deep smem ring
//   warp 1      : MMA issuer    — one thread issues tcgen05.mma (128 x BN x 16, bf16 -> fp32 in TMEM)
//   warp 2      : TMEM allocator (2 accumulator buffers: epilogue of tile i overlaps mainloop of tile i+1)
//   warps 4..7  : epilogue      — tcgen05.ld accumulator rows -> registers -> (+R | SwiGLU) -> global
// Operands may be K-major or MN-major (descriptor bits), so forward, dgrad and wgrad all run on the
// tensors where they lie: no transposes, no copies.
#include <stdio.h>
#include <stdlib.h>

#include "../../include/touchnet_b200.h"
#include "common.cuh"
#include "host.h"

namespace tn {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 256;
constexpr int GROUP_M = 8;

struct GemmParams {
  int M, N, K;
  int num_m, num_n, num_k;
  void* D;
  int64_t ldd;
  const void* R;
  int64_t ldr;
  // SwiGLU epilogue outputs (EPI == 1)
  bf16* G;
  bf16* U;
  bf16* H;
  int64_t ldh;
};

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;   // 16 KB
  static constexpr int B_BYTES = BN * BK * 2;   // 32 KB (BN=256) / 16 KB (BN=128)
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment slack
  static constexpr uint32_t TMEM_COLS = 2 * BN;                               // 512 / 256
};

__device__ __forceinline__ void decode_tile(int tile, int num_m, int num_n, int& m_blk, int& n_blk) {
  const int per_group = GROUP_M * num_n;
  const int g = tile / per_group;
  const int first_m = g * GROUP_M;
  const int gsize = min(num_m - first_m, GROUP_M);
  const int r = tile - g * per_group;
  m_blk = first_m + (r % gsize);
  n_blk = r / gsize;
}

// EPI: 0 = D = acc (+R), bf16 out; 1 = SwiGLU (acc cols [0,128)=gate, [128,256)=up); 2 = D = acc (+R), fp32 out
template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmB2, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (EPI == 1) tma_prefetch_desc(&tmB2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.num_m * p.num_n;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        decode_tile(tile, p.num_m, p.num_n, m_blk, n_blk);
        const int m0 = m_blk * BM;
        const int n0 = n_blk * (EPI == 1 ? BN / 2 : BN);
        for (int kb = 0; kb < p.num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* a_dst = sA + stage * Cfg::A_BYTES;
          uint8_t* b_dst = sB + stage * Cfg::B_BYTES;
          const int k0 = kb * BK;
          if (!A_MN) {
            tma_load_2d(a_dst, &tmA, &full_bar[stage], k0, m0);  // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)  // box {64 m, 64 k}
              tma_load_2d(a_dst + j * 8192, &tmA, &full_bar[stage], m0 + 64 * j, k0);
          }
          if (EPI == 1) {
            tma_load_2d(b_dst, &tmB, &full_bar[stage], k0, n0);                      // gate rows
            tma_load_2d(b_dst + (BN / 2) * 128, &tmB2, &full_bar[stage], k0, n0);    // up rows
          } else if (!B_MN) {
            tma_load_2d(b_dst, &tmB, &full_bar[stage], k0, n0);  // box {64 k, BN n}
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)  // box {64 n, 64 k}
              tma_load_2d(b_dst + j * 8192, &tmB, &full_bar[stage], n0 + 64 * j, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(acc * BN);
        for (int kb = 0; kb < p.num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = A_MN ? make_sdesc_sw128(a_addr + k * 2048, 8192, 1024)
                                     : make_sdesc_sw128(a_addr + k * 32, 0, 1024);
            const uint64_t db = B_MN ? make_sdesc_sw128(b_addr + k * 2048, 8192, 1024)
                                     : make_sdesc_sw128(b_addr + k * 32, 0, 1024);
            umma_ss(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const uint32_t quad = warp & 3u;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      decode_tile(tile, p.num_m, p.num_n, m_blk, n_blk);
      const int row = m_blk * BM + int(quad * 32 + lane);
      const bool row_ok = row < p.M;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + uint32_t(acc * BN) + ((quad * 32u) << 16);

      if (EPI == 1) {
        const int n0 = n_blk * (BN / 2);
#pragma unroll 1
        for (int c = 0; c < BN / 2; c += 32) {
          uint32_t g[32], u[32];
          tmem_ld32(t_row + c, g);
          tmem_ld32(t_row + BN / 2 + c, u);
          tmem_ld_wait();
          const int col = n0 + c;
          if (row_ok && col < p.N) {
            uint32_t pg[16], pu[16], ph[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              // mimic the unfused reference rounding: g,u -> bf16; silu(g) -> bf16; * u -> bf16
              const float g0 = bf16_round(__uint_as_float(g[2 * i])), g1 = bf16_round(__uint_as_float(g[2 * i + 1]));
              const float u0 = bf16_round(__uint_as_float(u[2 * i])), u1 = bf16_round(__uint_as_float(u[2 * i + 1]));
              const float s0 = bf16_round(g0 / (1.f + __expf(-g0))), s1 = bf16_round(g1 / (1.f + __expf(-g1)));
              pg[i] = pack_bf16x2(g0, g1);
              pu[i] = pack_bf16x2(u0, u1);
              ph[i] = pack_bf16x2(s0 * u0, s1 * u1);
            }
            const int64_t off = int64_t(row) * p.ldh + col;
            if (col + 32 <= p.N) {
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                reinterpret_cast<uint4*>(p.H + off)[v] = make_uint4(ph[4 * v], ph[4 * v + 1], ph[4 * v + 2], ph[4 * v + 3]);
                if (p.G) reinterpret_cast<uint4*>(p.G + off)[v] = make_uint4(pg[4 * v], pg[4 * v + 1], pg[4 * v + 2], pg[4 * v + 3]);
                if (p.U) reinterpret_cast<uint4*>(p.U + off)[v] = make_uint4(pu[4 * v], pu[4 * v + 1], pu[4 * v + 2], pu[4 * v + 3]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                if (col + i >= p.N) continue;
                const uint32_t wh = ph[i >> 1], wg = pg[i >> 1], wu = pu[i >> 1];
                const uint16_t hh = (i & 1) ? uint16_t(wh >> 16) : uint16_t(wh);
                const uint16_t gg = (i & 1) ? uint16_t(wg >> 16) : uint16_t(wg);
                const uint16_t uu = (i & 1) ? uint16_t(wu >> 16) : uint16_t(wu);
                reinterpret_cast<uint16_t*>(p.H)[off + i] = hh;
                if (p.G) reinterpret_cast<uint16_t*>(p.G)[off + i] = gg;
                if (p.U) reinterpret_cast<uint16_t*>(p.U)[off + i] = uu;
              }
            }
          }
        }
      } else {
        const int n0 = n_blk * BN;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          tmem_ld32(t_row + c, v);
          tmem_ld_wait();
          const int col = n0 + c;
          if (row_ok && col < p.N) {
            if (EPI == 2) {
              float* dptr = reinterpret_cast<float*>(p.D) + int64_t(row) * p.ldd + col;
              const float* rptr = p.R ? reinterpret_cast<const float*>(p.R) + int64_t(row) * p.ldr + col : nullptr;
              if (col + 32 <= p.N) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  float4 o = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                         __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                  if (rptr) {
                    const float4 r4 = reinterpret_cast<const float4*>(rptr)[q];
                    o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
                  }
                  reinterpret_cast<float4*>(dptr)[q] = o;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (col + i < p.N) dptr[i] = __uint_as_float(v[i]) + (rptr ? rptr[i] : 0.f);
              }
            } else {
              bf16* dptr = reinterpret_cast<bf16*>(p.D) + int64_t(row) * p.ldd + col;
              const bf16* rptr = p.R ? reinterpret_cast<const bf16*>(p.R) + int64_t(row) * p.ldr + col : nullptr;
              if (col + 32 <= p.N) {
                uint4 r4[4];
                if (rptr) {
#pragma unroll
                  for (int q = 0; q < 4; ++q) r4[q] = reinterpret_cast<const uint4*>(rptr)[q];
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  uint32_t w[4];
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    float lo = __uint_as_float(v[8 * q + 2 * j]), hi = __uint_as_float(v[8 * q + 2 * j + 1]);
                    if (rptr) {
                      // reference order: matmul result rounded to bf16, then bf16 add with the residual
                      const uint32_t rw = (j == 0) ? r4[q].x : (j == 1) ? r4[q].y : (j == 2) ? r4[q].z : r4[q].w;
                      lo = bf16_round(lo) + bf16lo(rw);
                      hi = bf16_round(hi) + bf16hi(rw);
                    }
                    w[j] = pack_bf16x2(lo, hi);
                  }
                  reinterpret_cast<uint4*>(dptr)[q] = make_uint4(w[0], w[1], w[2], w[3]);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  if (col + i < p.N) {
                    float x = __uint_as_float(v[i]);
                    if (rptr) x = bf16_round(x) + __bfloat162float(rptr[i]);
                    dptr[i] = __float2bfloat16_rn(x);
                  }
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int BN, bool A_MN, bool B_MN, int EPI>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmB2, const GemmParams& p,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_kernel<BN, A_MN, B_MN, EPI>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    TN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int tiles = p.num_m * p.num_n;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, tmB2, p);
  TN_CHECK_CUDA(cudaGetLastError());
  return TN_OK;
}

}  // namespace tn

namespace tn {
int gemm_pair_dispatch(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* D, int64_t ldd,
                       int d_f32, const void* R, int64_t ldr, int M, int N, int K, cudaStream_t stream);
int gemm_pair_swiglu_dispatch(const void* X, int64_t ldx, const void* Wg, const void* Wu, int64_t ldw, void* G, void* U,
                              void* H, int64_t ldh, int M, int N, int K, cudaStream_t stream);
int gemm_pair_seg_dispatch(int mode, const void* A, int64_t lda, const void* const* Bs, int64_t ldb, void* const* Ds,
                           int64_t ldd, int d_f32, const int* seg, int M, int N, int K, const void* rope_cos,
                           const void* rope_sin, cudaStream_t stream);
// TN_GEMM_PAIR=0 forces the single-CTA kernels (A/B measurements); default: CTA pairs whenever the tile fits
static bool use_pair() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TN_GEMM_PAIR");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
}  // namespace tn

using namespace tn;

extern "C" int tn_gemm_bf16(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* D,
                            int64_t ldd, int d_f32, const void* R, int64_t ldr, int M, int N, int K,
                            tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(A && B && D, "tn_gemm_bf16: null pointer");
  TN_REQUIRE(M > 0 && N > 0 && K > 0, "tn_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  TN_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "tn_gemm_bf16: lda/ldb must be multiples of 8 elements (16 B)");
  TN_REQUIRE(ldd % (d_f32 ? 4 : 8) == 0 && (!R || ldr % (d_f32 ? 4 : 8) == 0), "tn_gemm_bf16: ldd/ldr alignment");
  TN_REQUIRE((reinterpret_cast<uintptr_t>(D) & 15) == 0 && (!R || (reinterpret_cast<uintptr_t>(R) & 15) == 0),
             "tn_gemm_bf16: D/R must be 16 B aligned");
  TN_REQUIRE(!(a_mn && !b_mn), "tn_gemm_bf16: (a_mn=1, b_mn=0) layout is not instantiated");

  if (use_pair() && M >= 256 && N >= 256)
    return gemm_pair_dispatch(A, lda, a_mn, B, ldb, b_mn, D, ldd, d_f32, R, ldr, M, N, K, stream);

  // small-N problems use the narrower tile so that the persistent grid still fills 148 SMs
  const bool narrow = (int64_t((M + BM - 1) / BM) * ((N + 255) / 256) < 148) || (N <= 128);
  const int BN = narrow ? 128 : 256;

  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.num_m = (M + BM - 1) / BM;
  p.num_n = (N + BN - 1) / BN;
  p.num_k = (K + BK - 1) / BK;
  p.D = D; p.ldd = ldd; p.R = R; p.ldr = ldr;

  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn) rc = encode_tmap_2d(&tmA, A, 2, uint64_t(K), uint64_t(M), uint64_t(lda) * 2, 64, BM, true);
  else       rc = encode_tmap_2d(&tmA, A, 2, uint64_t(M), uint64_t(K), uint64_t(lda) * 2, 64, 64, true);
  if (rc) return rc;
  if (!b_mn) rc = encode_tmap_2d(&tmB, B, 2, uint64_t(K), uint64_t(N), uint64_t(ldb) * 2, 64, uint32_t(BN), true);
  else       rc = encode_tmap_2d(&tmB, B, 2, uint64_t(N), uint64_t(K), uint64_t(ldb) * 2, 64, 64, true);
  if (rc) return rc;

#define TN_GEMM_DISPATCH(BN_, AMN_, BMN_)                                                         \
  (d_f32 ? launch_gemm<BN_, AMN_, BMN_, 2>(tmA, tmB, tmB, p, stream)                              \
         : launch_gemm<BN_, AMN_, BMN_, 0>(tmA, tmB, tmB, p, stream))
  if (!a_mn && !b_mn) return narrow ? TN_GEMM_DISPATCH(128, false, false) : TN_GEMM_DISPATCH(256, false, false);
  if (!a_mn && b_mn)  return narrow ? TN_GEMM_DISPATCH(128, false, true) : TN_GEMM_DISPATCH(256, false, true);
  return narrow ? TN_GEMM_DISPATCH(128, true, true) : TN_GEMM_DISPATCH(256, true, true);
#undef TN_GEMM_DISPATCH
}

extern "C" int tn_gemm_swiglu_bf16(const void* X, int64_t ldx, const void* Wg, const void* Wu, int64_t ldw, void* G,
                                   void* U, void* H, int64_t ldh, int M, int N, int K, tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(X && Wg && Wu && H, "tn_gemm_swiglu_bf16: null pointer");
  TN_REQUIRE(M > 0 && N > 0 && K > 0, "tn_gemm_swiglu_bf16: empty problem");
  TN_REQUIRE(ldx % 8 == 0 && ldw % 8 == 0 && ldh % 8 == 0, "tn_gemm_swiglu_bf16: leading dims must be multiples of 8");
  TN_REQUIRE((reinterpret_cast<uintptr_t>(H) & 15) == 0 && (!G || (reinterpret_cast<uintptr_t>(G) & 15) == 0) &&
                 (!U || (reinterpret_cast<uintptr_t>(U) & 15) == 0),
             "tn_gemm_swiglu_bf16: outputs must be 16 B aligned");
  if (use_pair() && M >= 256 && N >= 128 && G && U)
    return gemm_pair_swiglu_dispatch(X, ldx, Wg, Wu, ldw, G, U, H, ldh, M, N, K, stream);
  constexpr int BN = 256;  // 128 gate + 128 up accumulator columns
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.num_m = (M + BM - 1) / BM;
  p.num_n = (N + BN / 2 - 1) / (BN / 2);
  p.num_k = (K + BK - 1) / BK;
  p.G = static_cast<bf16*>(G); p.U = static_cast<bf16*>(U); p.H = static_cast<bf16*>(H); p.ldh = ldh;
  CUtensorMap tmA, tmG, tmU;
  int rc = encode_tmap_2d(&tmA, X, 2, uint64_t(K), uint64_t(M), uint64_t(ldx) * 2, 64, BM, true);
  if (rc) return rc;
  rc = encode_tmap_2d(&tmG, Wg, 2, uint64_t(K), uint64_t(N), uint64_t(ldw) * 2, 64, BN / 2, true);
  if (rc) return rc;
  rc = encode_tmap_2d(&tmU, Wu, 2, uint64_t(K), uint64_t(N), uint64_t(ldw) * 2, 64, BN / 2, true);
  if (rc) return rc;
  return launch_gemm<BN, false, false, 1>(tmA, tmG, tmU, p, stream);
}

namespace tn {
int gemm_pair_dswiglu_dispatch(const void* dY, int64_t lddy, const void* W, int64_t ldw, const void* G, const void* U,
                               int64_t ldgu, void* dG, void* dU, int64_t lddg, int M, int N, int K, cudaStream_t stream);
}

extern "C" int tn_gemm_dswiglu_bf16(const void* dY, int64_t lddy, const void* W, int64_t ldw, const void* G, const void* U,
                                    int64_t ldgu, void* dG, void* dU, int64_t lddg, int M, int N, int K,
                                    tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(dY && W && G && U && dG && dU, "tn_gemm_dswiglu_bf16: null pointer");
  TN_REQUIRE(M >= 256 && N >= 256 && K > 0, "tn_gemm_dswiglu_bf16: needs M, N >= 256 (use tn_gemm_bf16 + tn_swiglu_bwd_bf16 below that)");
  TN_REQUIRE(lddy % 8 == 0 && ldw % 8 == 0 && ldgu % 8 == 0 && lddg % 8 == 0 && N % 8 == 0,
             "tn_gemm_dswiglu_bf16: leading dims and N must be multiples of 8");
  TN_REQUIRE(((reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(U) | reinterpret_cast<uintptr_t>(dG) |
               reinterpret_cast<uintptr_t>(dU)) & 15) == 0, "tn_gemm_dswiglu_bf16: G, U, dG, dU must be 16 B aligned");
  return gemm_pair_dswiglu_dispatch(dY, lddy, W, ldw, G, U, ldgu, dG, dU, lddg, M, N, K, stream);
}

extern "C" int tn_gemm_qkv_bf16(int mode, const void* A, int64_t lda, const void* B0, const void* B1, const void* B2,
                                int64_t ldb, void* D0, void* D1, void* D2, int64_t ldd, int d_f32, int s0, int s1, int s2,
                                int M, int N, int K, const void* rope_cos, const void* rope_sin, tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(A && B0 && D0, "tn_gemm_qkv_bf16: null pointer");
  TN_REQUIRE(mode >= 0 && mode <= 2, "tn_gemm_qkv_bf16: mode must be 0 (forward), 1 (dgrad) or 2 (wgrad)");
  TN_REQUIRE(s0 > 0 && s1 > 0 && s2 > 0 && s0 % 256 == 0 && s1 % 256 == 0 && s2 % 256 == 0,
             "tn_gemm_qkv_bf16: segment sizes (%d,%d,%d) must be positive multiples of 256", s0, s1, s2);
  const int tot = s0 + s1 + s2;
  TN_REQUIRE((mode == 0 && N == tot) || (mode == 1 && K == tot) || (mode == 2 && M == tot),
             "tn_gemm_qkv_bf16: segments do not add up to the segmented dimension");
  TN_REQUIRE(M >= 256 && N >= 256 && lda % 8 == 0 && ldb % 8 == 0 && ldd % (d_f32 ? 4 : 8) == 0,
             "tn_gemm_qkv_bf16: needs M,N >= 256 and 16-byte aligned leading dimensions");
  TN_REQUIRE(mode == 2 || (B1 && B2), "tn_gemm_qkv_bf16: B1/B2 required");
  TN_REQUIRE(mode != 2 || (D1 && D2), "tn_gemm_qkv_bf16: D1/D2 required for wgrad");
  const void* Bs[3] = {B0, B1, B2};
  void* Ds[3] = {D0, D1, D2};
  const int seg[3] = {s0, s1, s2};
  TN_REQUIRE((rope_cos == nullptr) == (rope_sin == nullptr), "tn_gemm_qkv_bf16: cos and sin tables come in pairs");
  TN_REQUIRE(rope_cos == nullptr || mode == 0, "tn_gemm_qkv_bf16: fused RoPE exists for the forward projection only");
  return gemm_pair_seg_dispatch(mode, A, lda, Bs, ldb, Ds, ldd, d_f32, seg, M, N, K, rope_cos, rope_sin, stream);
}
