// touchnet_b200 :: HBM-bound kernels of the packed decoder layer (RMSNorm(+residual), RoPE, SwiGLU backward,
// embedding add, fp32->bf16 cast).  All are single-pass over HBM with 16-byte vector accesses; rounding points
// mirror the unfused bf16 reference ops so results stay within bf16 noise of the HF modules they replace.
//
// Reference call sites:
//   RMSNorm   hf:models/llama/modeling_llama.py:62-67   (weights reset by touchnet/models/llama/__init__.py:28-31)
//   residual  hf:models/llama/modeling_llama.py:303-333
//   RoPE      hf:models/llama/modeling_llama.py:124-168  (called at touchnet/models/llama/pipeline_llama.py:83)
//   SwiGLU    hf:models/llama/modeling_llama.py:182-184
//   embed add touchnet/models/touch_audio/modeling_touch_audio.py:124-134
#include "../../include/touchnet_b200.h"
#include "common.cuh"
#include "host.h"

namespace tn {

// ---------------------------------------------------------------------------------------------------------------
// RMSNorm forward: one warp per row, whole row cached in registers (NV uint4 vectors per lane).
// ---------------------------------------------------------------------------------------------------------------
template <int NV, bool HAS_R>
__global__ void __launch_bounds__(256, (NV <= 16) ? 2 : 1) rmsnorm_fwd_kernel(const uint4* __restrict__ X, const uint4* __restrict__ R,
                                                          const void* __restrict__ w, int w_is_f32,
                                                          uint4* __restrict__ S_out, uint4* __restrict__ Y,
                                                          float* __restrict__ rstd_out, int64_t rows, int d,
                                                          float eps) {
  const int nvec = d >> 3;
  const int64_t row = int64_t(blockIdx.x) * 8 + warp_id();
  if (row >= rows) return;
  const uint32_t lane = lane_id();
  const uint4* x = X + row * nvec;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  uint4 v[NV];
  // all loads of the row are issued back to back (no control flow between them): NV x 16 B in flight per lane
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    v[i] = (c < nvec) ? x[c] : zero4;
  }
  if (HAS_R) {                               // fused residual add (own instantiation; not used by the decoder block)
    const uint4* r = R + row * nvec;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const uint4 b = r[c];
        uint4 a = v[i];
        a.x = pack_bf16x2(bf16lo(a.x) + bf16lo(b.x), bf16hi(a.x) + bf16hi(b.x));
        a.y = pack_bf16x2(bf16lo(a.y) + bf16lo(b.y), bf16hi(a.y) + bf16hi(b.y));
        a.z = pack_bf16x2(bf16lo(a.z) + bf16lo(b.z), bf16hi(a.z) + bf16hi(b.z));
        a.w = pack_bf16x2(bf16lo(a.w) + bf16lo(b.w), bf16hi(a.w) + bf16hi(b.w));
        if (S_out) S_out[row * nvec + c] = a;
        v[i] = a;
      }
    }
  }
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint32_t ws[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float lo = bf16lo(ws[j]), hi = bf16hi(ws[j]);
      ss += lo * lo + hi * hi;
    }
  }
  ss = warp_sum(ss);
  const float rstd = rsqrtf(ss / float(d) + eps);
  if (lane == 0) rstd_out[row] = rstd;
  // weight as packed bf16 pairs (fp32 master weights are rounded here, exactly what `.to(bf16)` does); the (L1/L2-resident)
  // weight vectors are fetched in groups of G (register budget: 2 CTAs per SM), loads first, then arithmetic
  constexpr int G = NV < 2 ? NV : 2;
#pragma unroll
  for (int i0 = 0; i0 < NV; i0 += G) {
    uint4 wp[G];
    if (w_is_f32) {
      float4 w0[G], w1[G];
#pragma unroll
      for (int k = 0; k < G; ++k) {
        const int c = lane + 32 * (i0 + k);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        w0[k] = (c < nvec) ? reinterpret_cast<const float4*>(w)[2 * c] : z;
        w1[k] = (c < nvec) ? reinterpret_cast<const float4*>(w)[2 * c + 1] : z;
      }
#pragma unroll
      for (int k = 0; k < G; ++k)
        wp[k] = make_uint4(pack_bf16x2(w0[k].x, w0[k].y), pack_bf16x2(w0[k].z, w0[k].w), pack_bf16x2(w1[k].x, w1[k].y),
                           pack_bf16x2(w1[k].z, w1[k].w));
    } else {
#pragma unroll
      for (int k = 0; k < G; ++k) {
        const int c = lane + 32 * (i0 + k);
        wp[k] = (c < nvec) ? reinterpret_cast<const uint4*>(w)[c] : zero4;
      }
    }
#pragma unroll
    for (int k = 0; k < G; ++k) {
      const int i = i0 + k;
      const int c = lane + 32 * i;
      const uint32_t ws[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
      const uint32_t wq[4] = {wp[k].x, wp[k].y, wp[k].z, wp[k].w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // reference: (x_fp32 * rstd).to(bf16) * weight   -> one packed convert, one packed bf16 multiply (the product of
        // two bf16 values is exact in fp32, so mul.rn.bf16x2 rounds once, like the fp32 product rounded to bf16)
        const uint32_t t = pack_bf16x2(bf16lo(ws[j]) * rstd, bf16hi(ws[j]) * rstd);
        o[j] = mul_bf16x2(t, wq[j]);
      }
      if (c < nvec) Y[row * nvec + c] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// RMSNorm backward: one CTA per row slice; thread t owns columns [8t, 8t+8) across all rows of its slice so the
// weight gradient accumulates in registers; per-row dot product via block reduction.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kNormBwdPartials = 592;  // 4 CTAs per SM x 148

template <int MAXT>
__global__ void __launch_bounds__(MAXT) rmsnorm_bwd_kernel(const uint4* __restrict__ S, const uint4* __restrict__ dY,
                                                           const uint4* __restrict__ dS_extra,
                                                           const void* __restrict__ w, int w_is_f32,
                                                           const float* __restrict__ rstd_in, uint4* __restrict__ dS,
                                                           float* __restrict__ dW_partial, int64_t rows, int d) {
  __shared__ float red[2][32];
  const int nvec = d >> 3;
  const int t = threadIdx.x;
  const bool active = t < nvec;
  const uint32_t warp = warp_id(), lane = lane_id();
  const int nwarps = blockDim.x >> 5;
  float wf[8];
  if (active) {
    if (w_is_f32) {
      const float4 w0 = reinterpret_cast<const float4*>(w)[2 * t], w1 = reinterpret_cast<const float4*>(w)[2 * t + 1];
      wf[0] = bf16_round(w0.x); wf[1] = bf16_round(w0.y); wf[2] = bf16_round(w0.z); wf[3] = bf16_round(w0.w);
      wf[4] = bf16_round(w1.x); wf[5] = bf16_round(w1.y); wf[6] = bf16_round(w1.z); wf[7] = bf16_round(w1.w);
    } else {
      const uint4 wb = reinterpret_cast<const uint4*>(w)[t];
      wf[0] = bf16lo(wb.x); wf[1] = bf16hi(wb.x); wf[2] = bf16lo(wb.y); wf[3] = bf16hi(wb.y);
      wf[4] = bf16lo(wb.z); wf[5] = bf16hi(wb.z); wf[6] = bf16lo(wb.w); wf[7] = bf16hi(wb.w);
    }
  }
  float dw[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  // the next row's operands are requested before the current row's block reduction, so HBM latency hides behind it
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  uint4 sv = zero4, gv = zero4, ev = zero4;
  float rstd = 0.f;
  int64_t row = blockIdx.x;
  if (row < rows) {
    rstd = rstd_in[row];
    if (active) {
      sv = S[row * nvec + t];
      gv = dY[row * nvec + t];
      if (dS_extra) ev = dS_extra[row * nvec + t];
    }
  }
  int it = 0;
  for (; row < rows; row += gridDim.x, ++it) {
    const int64_t nrow = row + gridDim.x;
    uint4 sv_n = zero4, gv_n = zero4, ev_n = zero4;
    float rstd_n = 0.f;
    if (nrow < rows) {
      rstd_n = rstd_in[nrow];
      if (active) {
        sv_n = S[nrow * nvec + t];
        gv_n = dY[nrow * nvec + t];
        if (dS_extra) ev_n = dS_extra[nrow * nvec + t];
      }
    }
    float xn[8], g[8];
    float dot = 0.f;
    if (active) {
      const uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w};
      const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        xn[2 * j] = bf16lo(sw[j]) * rstd;
        xn[2 * j + 1] = bf16hi(sw[j]) * rstd;
        const float g0 = bf16lo(gw[j]), g1 = bf16hi(gw[j]);
        const uint32_t xr = pack_bf16x2(xn[2 * j], xn[2 * j + 1]);     // the bf16 normalised activations of the forward
        dw[2 * j] += g0 * bf16lo(xr);
        dw[2 * j + 1] += g1 * bf16hi(xr);
        g[2 * j] = g0 * wf[2 * j];
        g[2 * j + 1] = g1 * wf[2 * j + 1];
        dot += g[2 * j] * xn[2 * j] + g[2 * j + 1] * xn[2 * j + 1];
      }
    }
    dot = warp_sum(dot);
    float* rb = red[it & 1];
    if (lane == 0) rb[warp] = dot;
    __syncthreads();
    float tot = (lane < nwarps) ? rb[lane] : 0.f;
    tot = warp_sum(tot);
    const float mean_dot = tot / float(d);
    if (active) {
      uint32_t o[4];
      const uint32_t ew[4] = {ev.x, ev.y, ev.z, ev.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float lo = rstd * (g[2 * j] - xn[2 * j] * mean_dot);
        const float hi = rstd * (g[2 * j + 1] - xn[2 * j + 1] * mean_dot);
        o[j] = pack_bf16x2(lo, hi);
        if (dS_extra) o[j] = add_bf16x2(o[j], ew[j]);      // bf16 + bf16 of the residual branch, rounded once
      }
      dS[row * nvec + t] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    sv = sv_n; gv = gv_n; ev = ev_n; rstd = rstd_n;
  }
  if (active) {
    float4* out = reinterpret_cast<float4*>(dW_partial + int64_t(blockIdx.x) * d + 8 * t);
    out[0] = make_float4(dw[0], dw[1], dw[2], dw[3]);
    out[1] = make_float4(dw[4], dw[5], dw[6], dw[7]);
  }
}

// column sums of the [n_part, d] partials in a fixed order (deterministic): 32 columns per CTA, 32 row lanes
__global__ void __launch_bounds__(1024) colsum_kernel(const float* __restrict__ part, int n_part, int d,
                                                      float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  float acc = 0.f;
  if (col < d)
    for (int r = ry; r < n_part; r += 32) acc += part[int64_t(r) * d + col];
  tile[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && col < d) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 32; ++r) s += tile[r][cx];
    out[col] = s;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// RoPE
// ---------------------------------------------------------------------------------------------------------------
__global__ void rope_table_kernel(const int64_t* __restrict__ pos, const float* __restrict__ inv_freq, float scaling,
                                  bf16* __restrict__ cos_out, bf16* __restrict__ sin_out, int64_t rows, int half) {
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= rows * half) return;
  const int64_t r = idx / half;
  const int j = int(idx - r * half);
  const float ang = float(pos[r]) * inv_freq[j];  // fp32, as the reference (autocast disabled)
  float s, c;
  sincosf(ang, &s, &c);
  cos_out[idx] = __float2bfloat16_rn(c * scaling);
  sin_out[idx] = __float2bfloat16_rn(s * scaling);
}

// x [rows, n_heads, hd] (row stride ldx); each thread rotates 8 (j, j+hd/2) pairs.
__global__ void __launch_bounds__(256) rope_apply_kernel(bf16* __restrict__ X, int64_t ldx,
                                                         const bf16* __restrict__ cos_tab,
                                                         const bf16* __restrict__ sin_tab, int64_t rows, int n_heads,
                                                         int hd, int inverse) {
  const int half = hd >> 1;
  const int vec_per_head = half >> 3;
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = rows * n_heads * vec_per_head;
  if (idx >= total) return;
  const int vj = int(idx % vec_per_head);
  const int64_t rh = idx / vec_per_head;
  const int h = int(rh % n_heads);
  const int64_t r = rh / n_heads;
  bf16* base = X + r * ldx + int64_t(h) * hd + vj * 8;
  const uint4 lo4 = *reinterpret_cast<const uint4*>(base);
  const uint4 hi4 = *reinterpret_cast<const uint4*>(base + half);
  const uint4 c4 = *reinterpret_cast<const uint4*>(cos_tab + r * half + vj * 8);
  const uint4 s4 = *reinterpret_cast<const uint4*>(sin_tab + r * half + vj * 8);
  const uint32_t lw[4] = {lo4.x, lo4.y, lo4.z, lo4.w}, hw[4] = {hi4.x, hi4.y, hi4.z, hi4.w};
  const uint32_t cw[4] = {c4.x, c4.y, c4.z, c4.w}, sw[4] = {s4.x, s4.y, s4.z, s4.w};
  uint32_t ol[4], oh[4];
  const float sgn = inverse ? -1.f : 1.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float a[2] = {bf16lo(lw[j]), bf16hi(lw[j])};
    float b[2] = {bf16lo(hw[j]), bf16hi(hw[j])};
    float c[2] = {bf16lo(cw[j]), bf16hi(cw[j])};
    float s[2] = {sgn * bf16lo(sw[j]), sgn * bf16hi(sw[j])};
    float o1[2], o2[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      // q*cos + rotate_half(q)*sin with bf16 rounding after every op (unfused reference semantics)
      o1[e] = bf16_round(a[e] * c[e]) + bf16_round(-b[e] * s[e]);
      o2[e] = bf16_round(b[e] * c[e]) + bf16_round(a[e] * s[e]);
    }
    ol[j] = pack_bf16x2(o1[0], o1[1]);
    oh[j] = pack_bf16x2(o2[0], o2[1]);
  }
  *reinterpret_cast<uint4*>(base) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
  *reinterpret_cast<uint4*>(base + half) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
}

// ---------------------------------------------------------------------------------------------------------------
// SwiGLU backward
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) swiglu_bwd_kernel(const bf16* __restrict__ G, const bf16* __restrict__ U,
                                                         const bf16* __restrict__ dH, bf16* __restrict__ dG,
                                                         bf16* __restrict__ dU, int64_t rows, int64_t cols,
                                                         int64_t ld) {
  const int64_t vec_per_row = cols >> 3;
  const int64_t total = rows * vec_per_row;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = idx / vec_per_row;
    const int64_t off = r * ld + (idx - r * vec_per_row) * 8;
    const uint4 g4 = *reinterpret_cast<const uint4*>(G + off);
    const uint4 u4 = *reinterpret_cast<const uint4*>(U + off);
    const uint4 d4 = *reinterpret_cast<const uint4*>(dH + off);
    const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w}, uw[4] = {u4.x, u4.y, u4.z, u4.w}, dw[4] = {d4.x, d4.y, d4.z, d4.w};
    uint32_t og[4], ou[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float rg[2], ru[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float g = e ? bf16hi(gw[j]) : bf16lo(gw[j]);
        const float u = e ? bf16hi(uw[j]) : bf16lo(uw[j]);
        const float dh = e ? bf16hi(dw[j]) : bf16lo(dw[j]);
        const float sig = 1.f / (1.f + __expf(-g));
        const float silu = g * sig;
        ru[e] = dh * bf16_round(silu);                              // dU = dH * silu(g) (silu output was bf16)
        rg[e] = bf16_round(dh * u) * (sig * (1.f + g * (1.f - sig)));  // dG = (dH*u) * silu'(g)
      }
      og[j] = pack_bf16x2(rg[0], rg[1]);
      ou[j] = pack_bf16x2(ru[0], ru[1]);
    }
    *reinterpret_cast<uint4*>(dG + off) = make_uint4(og[0], og[1], og[2], og[3]);
    *reinterpret_cast<uint4*>(dU + off) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// embedding gather + projector add (+ NaN flag)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_add_kernel(const int64_t* __restrict__ ids, const void* __restrict__ embed,
                                                        int embed_is_f32, const uint4* __restrict__ P,
                                                        uint4* __restrict__ E, int32_t* __restrict__ nan_flag,
                                                        int64_t rows, int d, int64_t vocab) {
  const int nvec = d >> 3;
  const int64_t total = rows * nvec;
  bool bad = false;
  for (int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = idx / nvec;
    const int c = int(idx - r * nvec);
    int64_t id = ids ? ids[r] : -1;
    float e[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (id >= 0 && id < vocab) {
      if (embed_is_f32) {
        const float4* src = reinterpret_cast<const float4*>(embed) + (id * nvec + c) * 2;
        const float4 a = src[0], b = src[1];
        e[0] = bf16_round(a.x); e[1] = bf16_round(a.y); e[2] = bf16_round(a.z); e[3] = bf16_round(a.w);
        e[4] = bf16_round(b.x); e[5] = bf16_round(b.y); e[6] = bf16_round(b.z); e[7] = bf16_round(b.w);
      } else {
        const uint4 a = reinterpret_cast<const uint4*>(embed)[id * nvec + c];
        e[0] = bf16lo(a.x); e[1] = bf16hi(a.x); e[2] = bf16lo(a.y); e[3] = bf16hi(a.y);
        e[4] = bf16lo(a.z); e[5] = bf16hi(a.z); e[6] = bf16lo(a.w); e[7] = bf16hi(a.w);
      }
    }
    if (P) {
      const uint4 p = P[idx];
      e[0] += bf16lo(p.x); e[1] += bf16hi(p.x); e[2] += bf16lo(p.y); e[3] += bf16hi(p.y);
      e[4] += bf16lo(p.z); e[5] += bf16hi(p.z); e[6] += bf16lo(p.w); e[7] += bf16hi(p.w);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) bad |= (e[j] != e[j]);
    E[idx] = make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]),
                        pack_bf16x2(e[6], e[7]));
  }
  if (nan_flag && __any_sync(0xffffffffu, bad) && lane_id() == 0) atomicOr(nan_flag, 1);
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst,
                                                            int64_t nvec, const float* __restrict__ src_tail,
                                                            bf16* __restrict__ dst_tail, int ntail) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += int64_t(gridDim.x) * blockDim.x) {
    const float4 v = src[i];
    dst[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
  if (blockIdx.x == 0 && int(threadIdx.x) < ntail) dst_tail[threadIdx.x] = __float2bfloat16_rn(src_tail[threadIdx.x]);
}

}  // namespace tn

using namespace tn;

extern "C" int tn_rmsnorm_bwd_num_partials(void) { return kNormBwdPartials; }

extern "C" int tn_rmsnorm_fwd_bf16(const void* X, const void* R, const void* w, int w_is_f32, void* S_out, void* Y,
                                   float* rstd, int64_t rows, int d, float eps, tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(X && w && Y && rstd, "tn_rmsnorm_fwd_bf16: null pointer");
  TN_REQUIRE(d > 0 && d % 8 == 0 && d <= 8192, "tn_rmsnorm_fwd_bf16: d=%d must be a multiple of 8 and <= 8192", d);
  if (rows == 0) return TN_OK;
  const int nvec = d / 8;
  const int nv = (nvec + 31) / 32;
  const unsigned grid = unsigned((rows + 7) / 8);
#define TN_LAUNCH_NORM(NV)                                                                                         \
  do {                                                                                                             \
    if (R)                                                                                                         \
      rmsnorm_fwd_kernel<NV, true><<<grid, 256, 0, stream>>>(static_cast<const uint4*>(X),                         \
          static_cast<const uint4*>(R), w, w_is_f32, static_cast<uint4*>(S_out), static_cast<uint4*>(Y), rstd,    \
          rows, d, eps);                                                                                           \
    else                                                                                                           \
      rmsnorm_fwd_kernel<NV, false><<<grid, 256, 0, stream>>>(static_cast<const uint4*>(X), nullptr, w, w_is_f32, \
          nullptr, static_cast<uint4*>(Y), rstd, rows, d, eps);                                                    \
  } while (0)
  if (nv <= 1) TN_LAUNCH_NORM(1);
  else if (nv <= 2) TN_LAUNCH_NORM(2);
  else if (nv <= 4) TN_LAUNCH_NORM(4);
  else if (nv <= 8) TN_LAUNCH_NORM(8);
  else if (nv <= 16) TN_LAUNCH_NORM(16);
  else TN_LAUNCH_NORM(32);
#undef TN_LAUNCH_NORM
  TN_CHECK_CUDA(cudaGetLastError());
  return TN_OK;
}

extern "C" int tn_rmsnorm_bwd_bf16(const void* S, const void* dY, const void* dS_extra, const void* w, int w_is_f32,
                                   const float* rstd, void* dS, float* dW_partial, int num_partials, float* dW,
                                   int64_t rows, int d, tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(S && dY && w && rstd && dS && dW_partial, "tn_rmsnorm_bwd_bf16: null pointer");
  TN_REQUIRE(d > 0 && d % 8 == 0 && d <= 8192, "tn_rmsnorm_bwd_bf16: d=%d must be a multiple of 8 and <= 8192", d);
  TN_REQUIRE(num_partials == kNormBwdPartials, "tn_rmsnorm_bwd_bf16: dW_partial must have %d rows", kNormBwdPartials);
  const int threads = ((d / 8 + 31) / 32) * 32;
  if (threads <= 512)
    rmsnorm_bwd_kernel<512><<<kNormBwdPartials, threads, 0, stream>>>(
        static_cast<const uint4*>(S), static_cast<const uint4*>(dY), static_cast<const uint4*>(dS_extra), w, w_is_f32,
        rstd, static_cast<uint4*>(dS), dW_partial, rows, d);
  else
    rmsnorm_bwd_kernel<1024><<<kNormBwdPartials, threads, 0, stream>>>(
        static_cast<const uint4*>(S), static_cast<const uint4*>(dY), static_cast<const uint4*>(dS_extra), w, w_is_f32,
        rstd, static_cast<uint4*>(dS), dW_partial, rows, d);
  TN_CHECK_CUDA(cudaGetLastError());
  if (dW) {   // final weight gradient: column sums over the partial rows that were written (CTAs beyond `rows` wrote zeros)
    colsum_kernel<<<unsigned((d + 31) / 32), 1024, 0, stream>>>(dW_partial, kNormBwdPartials, d, dW);
    TN_CHECK_CUDA(cudaGetLastError());
  }
  return TN_OK;
}

extern "C" int tn_rope_table(const int64_t* position_ids, const float* inv_freq, float attention_scaling, void* cos_out,
                             void* sin_out, int64_t rows, int half_dim, tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(position_ids && inv_freq && cos_out && sin_out, "tn_rope_table: null pointer");
  if (rows == 0) return TN_OK;
  const int64_t total = rows * half_dim;
  rope_table_kernel<<<unsigned((total + 255) / 256), 256, 0, stream>>>(
      position_ids, inv_freq, attention_scaling, static_cast<bf16*>(cos_out), static_cast<bf16*>(sin_out), rows, half_dim);
  TN_CHECK_CUDA(cudaGetLastError());
  return TN_OK;
}

extern "C" int tn_rope_apply_bf16(void* X, int64_t ldx, const void* cos_tab, const void* sin_tab, int64_t rows,
                                  int n_heads, int head_dim, int inverse, tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(X && cos_tab && sin_tab, "tn_rope_apply_bf16: null pointer");
  TN_REQUIRE(head_dim % 16 == 0 && ldx % 8 == 0, "tn_rope_apply_bf16: head_dim %% 16 and ldx %% 8 must be 0");
  if (rows == 0) return TN_OK;
  const int64_t total = rows * n_heads * (head_dim / 16);
  rope_apply_kernel<<<unsigned((total + 255) / 256), 256, 0, stream>>>(
      static_cast<bf16*>(X), ldx, static_cast<const bf16*>(cos_tab), static_cast<const bf16*>(sin_tab), rows, n_heads,
      head_dim, inverse);
  TN_CHECK_CUDA(cudaGetLastError());
  return TN_OK;
}

extern "C" int tn_swiglu_bwd_bf16(const void* G, const void* U, const void* dH, void* dG, void* dU, int64_t rows,
                                  int64_t cols, int64_t ld, tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(G && U && dH && dG && dU, "tn_swiglu_bwd_bf16: null pointer");
  TN_REQUIRE(cols % 8 == 0 && ld % 8 == 0, "tn_swiglu_bwd_bf16: cols and ld must be multiples of 8");
  if (rows == 0) return TN_OK;
  const int64_t total = rows * (cols / 8);
  const int64_t blocks = (total + 255) / 256;
  const unsigned grid = unsigned(blocks < int64_t(sm_count()) * 16 ? blocks : int64_t(sm_count()) * 16);
  swiglu_bwd_kernel<<<grid, 256, 0, stream>>>(static_cast<const bf16*>(G), static_cast<const bf16*>(U),
                                              static_cast<const bf16*>(dH), static_cast<bf16*>(dG),
                                              static_cast<bf16*>(dU), rows, cols, ld);
  TN_CHECK_CUDA(cudaGetLastError());
  return TN_OK;
}

extern "C" int tn_embed_add_bf16(const int64_t* input_ids, const void* embed, int embed_is_f32, const void* P, void* E,
                                 int32_t* nan_flag, int64_t rows, int d, int64_t vocab, tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(E && (input_ids == nullptr || embed != nullptr), "tn_embed_add_bf16: null pointer");
  TN_REQUIRE(d % 8 == 0, "tn_embed_add_bf16: d must be a multiple of 8");
  if (rows == 0) return TN_OK;
  const int64_t total = rows * (d / 8);
  const int64_t blocks = (total + 255) / 256;
  const unsigned grid = unsigned(blocks < int64_t(sm_count()) * 16 ? blocks : int64_t(sm_count()) * 16);
  embed_add_kernel<<<grid, 256, 0, stream>>>(input_ids, embed, embed_is_f32, static_cast<const uint4*>(P),
                                             static_cast<uint4*>(E), nan_flag, rows, d, vocab);
  TN_CHECK_CUDA(cudaGetLastError());
  return TN_OK;
}

extern "C" int tn_cast_f32_bf16(const float* src, void* dst, int64_t n, tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(src && dst, "tn_cast_f32_bf16: null pointer");
  if (n == 0) return TN_OK;
  TN_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0,
             "tn_cast_f32_bf16: alignment");
  const int64_t nvec = n / 4;
  const int ntail = int(n - nvec * 4);
  int64_t blocks = (nvec + 255) / 256;
  if (blocks < 1) blocks = 1;
  const unsigned grid = unsigned(blocks < int64_t(sm_count()) * 16 ? blocks : int64_t(sm_count()) * 16);
  cast_f32_bf16_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(src), static_cast<uint2*>(dst), nvec,
                                                 src + nvec * 4, static_cast<bf16*>(dst) + nvec * 4, ntail);
  TN_CHECK_CUDA(cudaGetLastError());
  return TN_OK;
}
