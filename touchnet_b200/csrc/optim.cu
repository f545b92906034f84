// touchnet_b200 :: optimizer step next to the path.
//
// Replaces   touchnet/utils/optimizer.py:127-172   torch.optim.AdamW(betas=(0.9,0.95), weight_decay=0.1, fused=True)
//            touchnet/utils/distributed.py:426-491 clip_grad_norm_ (get_total_norm + clip_grads_with_norm_)
// on the local (FSDP-sharded) fp32 parameter / gradient / moment shards.  One pass per tensor: the clip coefficient is
// read from device memory (no host sync between the norm and the step) and the bf16 working copy of the updated weight is
// written in the same pass, so the next forward needs no fp32->bf16 cast.  HBM-bound: 16 B read + 12 (+2) B written
// per parameter; the reference path (clip = read+write grads, fused AdamW, next-step cast) moves 8 + 28 + 6 = 42 B.
#include "../../include/touchnet_b200.h"
#include "common.cuh"
#include "host.h"

namespace tn {

constexpr int OPT_THREADS = 256;
constexpr int SUMSQ_MAX_BLOCKS = 1184;   // 8 x 148: partials buffer size (tn_sumsq_num_partials)

// deterministic two-stage sum of squares: fixed grid, fixed in-block tree, partials reduced by the caller in order
__global__ void __launch_bounds__(OPT_THREADS) sumsq_kernel(const float* __restrict__ x, int64_t n,
                                                            float* __restrict__ partials) {
  __shared__ float red[OPT_THREADS / 32];
  float acc = 0.f;
  const int64_t nvec = n >> 2;
  for (int64_t i = int64_t(blockIdx.x) * OPT_THREADS + threadIdx.x; i < nvec; i += int64_t(gridDim.x) * OPT_THREADS) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (int64_t i = nvec * 4 + threadIdx.x; i < n; i += OPT_THREADS) acc += x[i] * x[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < OPT_THREADS / 32 ? red[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) partials[blockIdx.x] = v;
  }
}

struct AdamWParams {
  float* p;
  const float* g;
  float* m;
  float* v;
  bf16* pb;
  int64_t n;
  float lr, beta1, beta2, eps, wd, bc1, bc2_sqrt;
  const float* grad_scale;
};

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, const AdamWParams& a, float gs) {
  // same operation order as torch's adam_math (ATen/native/cuda/fused_adam_utils.cuh) in ADAMW mode, fp32 throughout
  g *= gs;
  p -= a.lr * a.wd * p;
  m = a.beta1 * m + (1.f - a.beta1) * g;
  v = a.beta2 * v + (1.f - a.beta2) * g * g;
  const float step_size = a.lr / a.bc1;
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p -= step_size * m / denom;
}

__global__ void __launch_bounds__(OPT_THREADS) adamw_kernel(const AdamWParams a) {
  const float gs = a.grad_scale ? *a.grad_scale : 1.f;
  const int64_t nvec = a.n >> 2;
  for (int64_t i = int64_t(blockIdx.x) * OPT_THREADS + threadIdx.x; i < nvec; i += int64_t(gridDim.x) * OPT_THREADS) {
    float4 p = reinterpret_cast<float4*>(a.p)[i];
    const float4 g = reinterpret_cast<const float4*>(a.g)[i];
    float4 m = reinterpret_cast<float4*>(a.m)[i];
    float4 v = reinterpret_cast<float4*>(a.v)[i];
    adamw_one(p.x, g.x, m.x, v.x, a, gs);
    adamw_one(p.y, g.y, m.y, v.y, a, gs);
    adamw_one(p.z, g.z, m.z, v.z, a, gs);
    adamw_one(p.w, g.w, m.w, v.w, a, gs);
    reinterpret_cast<float4*>(a.p)[i] = p;
    reinterpret_cast<float4*>(a.m)[i] = m;
    reinterpret_cast<float4*>(a.v)[i] = v;
    if (a.pb) {
      uint2 o;
      o.x = pack_bf16x2(p.x, p.y);
      o.y = pack_bf16x2(p.z, p.w);
      reinterpret_cast<uint2*>(a.pb)[i] = o;
    }
  }
  if (blockIdx.x == 0) {
    for (int64_t i = nvec * 4 + threadIdx.x; i < a.n; i += OPT_THREADS) {
      float p = a.p[i], m = a.m[i], v = a.v[i];
      adamw_one(p, a.g[i], m, v, a, gs);
      a.p[i] = p; a.m[i] = m; a.v[i] = v;
      if (a.pb) a.pb[i] = __float2bfloat16(p);
    }
  }
}

__global__ void __launch_bounds__(OPT_THREADS) scale_kernel(float* __restrict__ x, int64_t n, const float* __restrict__ s) {
  const float f = *s;
  if (f == 1.f) return;
  const int64_t nvec = n >> 2;
  for (int64_t i = int64_t(blockIdx.x) * OPT_THREADS + threadIdx.x; i < nvec; i += int64_t(gridDim.x) * OPT_THREADS) {
    float4 v = reinterpret_cast<float4*>(x)[i];
    v.x *= f; v.y *= f; v.z *= f; v.w *= f;
    reinterpret_cast<float4*>(x)[i] = v;
  }
  if (blockIdx.x == 0)
    for (int64_t i = nvec * 4 + threadIdx.x; i < n; i += OPT_THREADS) x[i] *= f;
}

__global__ void __launch_bounds__(OPT_THREADS) scale_bf16_kernel(uint4* __restrict__ x, int64_t n, const float* __restrict__ s) {
  const float f = *s;
  if (f == 1.f) return;                       // the usual upstream gradient of a loss: nothing to do, nothing read
  const int64_t nvec = n >> 3;
  for (int64_t i = int64_t(blockIdx.x) * OPT_THREADS + threadIdx.x; i < nvec; i += int64_t(gridDim.x) * OPT_THREADS) {
    uint4 v = x[i];
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = pack_bf16x2(bf16lo(w[j]) * f, bf16hi(w[j]) * f);
    x[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  if (blockIdx.x == 0) {
    bf16* t = reinterpret_cast<bf16*>(x);
    for (int64_t i = nvec * 8 + threadIdx.x; i < n; i += OPT_THREADS) t[i] = __float2bfloat16_rn(__bfloat162float(t[i]) * f);
  }
}

static unsigned grid_for(int64_t n, int max_blocks) {
  int64_t b = (n / 4 + OPT_THREADS - 1) / OPT_THREADS;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return unsigned(b);
}

}  // namespace tn

using namespace tn;

extern "C" int tn_sumsq_num_partials(void) { return SUMSQ_MAX_BLOCKS; }

extern "C" int tn_sumsq_f32(const float* x, int64_t n, float* partials, int* n_partials_used, tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(x && partials && n_partials_used, "tn_sumsq_f32: null pointer");
  TN_REQUIRE(n >= 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0, "tn_sumsq_f32: x must be 16-byte aligned");
  const unsigned grid = grid_for(n, SUMSQ_MAX_BLOCKS);
  *n_partials_used = int(grid);
  sumsq_kernel<<<grid, OPT_THREADS, 0, stream>>>(x, n, partials);
  TN_CHECK_CUDA(cudaGetLastError());
  return TN_OK;
}

extern "C" int tn_adamw_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* param_bf16,
                            int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                            float bias_correction1, float bias_correction2_sqrt, const float* grad_scale,
                            tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(param && grad && exp_avg && exp_avg_sq, "tn_adamw_f32: null pointer");
  TN_REQUIRE(n >= 0, "tn_adamw_f32: n=%lld", (long long)n);
  const uintptr_t al = reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                       reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq);
  TN_REQUIRE((al & 15u) == 0 && (reinterpret_cast<uintptr_t>(param_bf16) & 7u) == 0,
             "tn_adamw_f32: tensors must be 16-byte aligned (bf16 copy 8-byte)");
  TN_REQUIRE(bias_correction1 > 0.f && bias_correction2_sqrt > 0.f, "tn_adamw_f32: bias corrections must be positive");
  if (n == 0) return TN_OK;
  AdamWParams a{param, grad, exp_avg, exp_avg_sq, static_cast<bf16*>(param_bf16), n, lr, beta1, beta2, eps,
                weight_decay, bias_correction1, bias_correction2_sqrt, grad_scale};
  adamw_kernel<<<grid_for(n, sm_count() * 8), OPT_THREADS, 0, stream>>>(a);
  TN_CHECK_CUDA(cudaGetLastError());
  return TN_OK;
}

extern "C" int tn_scale_f32(float* x, int64_t n, const float* scale, tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(x && scale, "tn_scale_f32: null pointer");
  TN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15u) == 0, "tn_scale_f32: x must be 16-byte aligned");
  if (n == 0) return TN_OK;
  scale_kernel<<<grid_for(n, sm_count() * 8), OPT_THREADS, 0, stream>>>(x, n, scale);
  TN_CHECK_CUDA(cudaGetLastError());
  return TN_OK;
}

extern "C" int tn_scale_bf16(void* x, int64_t n, const float* scale, tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(x && scale, "tn_scale_bf16: null pointer");
  TN_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15u) == 0, "tn_scale_bf16: x must be 16-byte aligned");
  if (n == 0) return TN_OK;
  scale_bf16_kernel<<<grid_for(n / 2, sm_count() * 8), OPT_THREADS, 0, stream>>>(static_cast<uint4*>(x), n, scale);
  TN_CHECK_CUDA(cudaGetLastError());
  return TN_OK;
}
