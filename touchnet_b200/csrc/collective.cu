// touchnet_b200 :: data-parallel collectives over NVLink peer memory (driven by touchnet_b200/fsdp_comm.py).
//
// FSDP2 (ref: touchnet/models/helper_func.py:134-202) reduce-scatters the fp32 gradients of every decoder block and
// all-gathers its bf16 parameters through NCCL.  With every rank's communication buffer mapped into every other rank's
// address space (torch symmetric memory = CUDA IPC over NVLink/NVSwitch) both are plain pull kernels:
//   reduce-scatter  out[i] = scale * sum_p in_p[shard_offset + i]   - rank r reads ITS shard from all N inputs
//   all-gather      out[p*n + i] = in_p[i]                          - rank r reads every rank's shard
// One launch each, fixed summation order (bit-reproducible), no staging copies, grid sized well below the SM count so the
// tensor-core kernels they overlap with keep their SMs (the loads are NVLink-latency bound, not issue bound).
// Synchronisation (inputs complete before the pull, pull complete before the buffers are reused) is the caller's
// device-side barrier (touchnet_b200/fsdp_comm.py).
#include "../../include/touchnet_b200.h"
#include "common.cuh"
#include "host.h"

namespace tn {

constexpr int COLL_MAX_PEERS = 16;
constexpr int COLL_THREADS = 512;

struct PeerPtrs {
  const void* p[COLL_MAX_PEERS];
};

template <int N>
__device__ __forceinline__ float4 sum_peers(const PeerPtrs& in, int64_t idx4) {
  float4 v[N];
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = reinterpret_cast<const float4*>(in.p[k])[idx4];   // N loads in flight per thread
  float4 s = v[0];
#pragma unroll
  for (int k = 1; k < N; ++k) { s.x += v[k].x; s.y += v[k].y; s.z += v[k].z; s.w += v[k].w; }
  return s;
}

__global__ void __launch_bounds__(COLL_THREADS) peer_reduce_scatter_kernel(const PeerPtrs in, int n_peers,
                                                                           int64_t shard_offset, float* __restrict__ out,
                                                                           int64_t numel, float scale) {
  const int64_t nvec = numel >> 2, off4 = shard_offset >> 2;
  for (int64_t i = int64_t(blockIdx.x) * COLL_THREADS + threadIdx.x; i < nvec; i += int64_t(gridDim.x) * COLL_THREADS) {
    float4 s;
    switch (n_peers) {
      case 2: s = sum_peers<2>(in, off4 + i); break;
      case 4: s = sum_peers<4>(in, off4 + i); break;
      case 8: s = sum_peers<8>(in, off4 + i); break;
      default: {
        s = reinterpret_cast<const float4*>(in.p[0])[off4 + i];
        for (int k = 1; k < n_peers; ++k) {
          const float4 v = reinterpret_cast<const float4*>(in.p[k])[off4 + i];
          s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
      }
    }
    s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
    reinterpret_cast<float4*>(out)[i] = s;
  }
  if (blockIdx.x == 0)
    for (int64_t i = nvec * 4 + threadIdx.x; i < numel; i += COLL_THREADS) {
      float s = 0.f;
      for (int k = 0; k < n_peers; ++k) s += static_cast<const float*>(in.p[k])[shard_offset + i];
      out[i] = s * scale;
    }
}

// bf16 inputs, fp32 sum: out[i] = scale * sum_p float(in_p[i]).  The reduce step of the direct-push reduce-scatter
// (fsdp_comm.PushReduceScatter): every rank's bf16 gradient chunk for this rank already sits in a local receive slot
// (copy-engine pushes), so this is one HBM-bound pass; fp32 accumulation in rank order = what casting every chunk to fp32
// and reduce-scattering in fp32 (FSDP2's reduce_dtype=float32) computes.
__global__ void __launch_bounds__(COLL_THREADS) reduce_bf16_to_f32_kernel(const PeerPtrs in, int n_peers,
                                                                          float* __restrict__ out, int64_t numel,
                                                                          float scale) {
  const int64_t nvec = numel >> 3;
  for (int64_t i = int64_t(blockIdx.x) * COLL_THREADS + threadIdx.x; i < nvec; i += int64_t(gridDim.x) * COLL_THREADS) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < n_peers; ++k) {
      const uint4 v = reinterpret_cast<const uint4*>(in.p[k])[i];
      acc[0] += bf16lo(v.x); acc[1] += bf16hi(v.x); acc[2] += bf16lo(v.y); acc[3] += bf16hi(v.y);
      acc[4] += bf16lo(v.z); acc[5] += bf16hi(v.z); acc[6] += bf16lo(v.w); acc[7] += bf16hi(v.w);
    }
    float4* o = reinterpret_cast<float4*>(out) + 2 * i;
    o[0] = make_float4(acc[0] * scale, acc[1] * scale, acc[2] * scale, acc[3] * scale);
    o[1] = make_float4(acc[4] * scale, acc[5] * scale, acc[6] * scale, acc[7] * scale);
  }
  if (blockIdx.x == 0)
    for (int64_t i = nvec * 8 + threadIdx.x; i < numel; i += COLL_THREADS) {
      float s = 0.f;
      for (int k = 0; k < n_peers; ++k) s += __bfloat162float(static_cast<const bf16*>(in.p[k])[i]);
      out[i] = s * scale;
    }
}

__global__ void __launch_bounds__(COLL_THREADS) peer_all_gather_kernel(const PeerPtrs in, int n_peers, int64_t vec_each,
                                                                       uint4* __restrict__ out) {
  // blockIdx.y = source rank; 16-byte vectors
  const int p = blockIdx.y;
  const uint4* src = static_cast<const uint4*>(in.p[p]);
  uint4* dst = out + int64_t(p) * vec_each;
  for (int64_t i = int64_t(blockIdx.x) * COLL_THREADS + threadIdx.x; i < vec_each; i += int64_t(gridDim.x) * COLL_THREADS)
    dst[i] = src[i];
}

}  // namespace tn

using namespace tn;

extern "C" int tn_peer_reduce_scatter_f32(const void* const* inputs, int n_peers, int64_t shard_offset, float* out,
                                          int64_t numel, float scale, int max_ctas, tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(inputs && out, "tn_peer_reduce_scatter_f32: null pointer");
  TN_REQUIRE(n_peers >= 1 && n_peers <= COLL_MAX_PEERS, "tn_peer_reduce_scatter_f32: n_peers=%d out of range", n_peers);
  TN_REQUIRE(numel >= 0 && shard_offset >= 0 && shard_offset % 4 == 0,
             "tn_peer_reduce_scatter_f32: shard offset must be a multiple of 4 elements");
  PeerPtrs pp{};
  for (int k = 0; k < n_peers; ++k) {
    TN_REQUIRE(inputs[k] && (reinterpret_cast<uintptr_t>(inputs[k]) & 15u) == 0,
               "tn_peer_reduce_scatter_f32: input %d null or not 16-byte aligned", k);
    pp.p[k] = inputs[k];
  }
  TN_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15u) == 0, "tn_peer_reduce_scatter_f32: out not 16-byte aligned");
  if (numel == 0) return TN_OK;
  int64_t grid = (numel / 4 + COLL_THREADS - 1) / COLL_THREADS;
  const int cap = max_ctas > 0 ? max_ctas : 32;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  peer_reduce_scatter_kernel<<<unsigned(grid), COLL_THREADS, 0, stream>>>(pp, n_peers, shard_offset, out, numel, scale);
  TN_CHECK_CUDA(cudaGetLastError());
  return TN_OK;
}

extern "C" int tn_peer_all_gather(const void* const* inputs, int n_peers, int64_t bytes_each, void* out, int max_ctas,
                                  tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(inputs && out, "tn_peer_all_gather: null pointer");
  TN_REQUIRE(n_peers >= 1 && n_peers <= COLL_MAX_PEERS, "tn_peer_all_gather: n_peers=%d out of range", n_peers);
  TN_REQUIRE(bytes_each >= 0 && bytes_each % 16 == 0, "tn_peer_all_gather: bytes per rank must be a multiple of 16");
  PeerPtrs pp{};
  for (int k = 0; k < n_peers; ++k) {
    TN_REQUIRE(inputs[k] && (reinterpret_cast<uintptr_t>(inputs[k]) & 15u) == 0,
               "tn_peer_all_gather: input %d null or not 16-byte aligned", k);
    pp.p[k] = inputs[k];
  }
  TN_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15u) == 0, "tn_peer_all_gather: out not 16-byte aligned");
  if (bytes_each == 0) return TN_OK;
  const int64_t vec_each = bytes_each / 16;
  int64_t gx = (vec_each + COLL_THREADS - 1) / COLL_THREADS;
  const int cap = (max_ctas > 0 ? max_ctas : 32) / n_peers;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  peer_all_gather_kernel<<<dim3(unsigned(gx), unsigned(n_peers)), COLL_THREADS, 0, stream>>>(pp, n_peers, vec_each,
                                                                                           static_cast<uint4*>(out));
  TN_CHECK_CUDA(cudaGetLastError());
  return TN_OK;
}

extern "C" int tn_reduce_bf16_to_f32(const void* const* inputs, int n_inputs, float* out, int64_t numel, float scale,
                                     int max_ctas, tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(inputs && out, "tn_reduce_bf16_to_f32: null pointer");
  TN_REQUIRE(n_inputs >= 1 && n_inputs <= COLL_MAX_PEERS, "tn_reduce_bf16_to_f32: n_inputs=%d out of range", n_inputs);
  PeerPtrs pp{};
  for (int k = 0; k < n_inputs; ++k) {
    TN_REQUIRE(inputs[k] && (reinterpret_cast<uintptr_t>(inputs[k]) & 15u) == 0,
               "tn_reduce_bf16_to_f32: input %d null or not 16-byte aligned", k);
    pp.p[k] = inputs[k];
  }
  TN_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15u) == 0, "tn_reduce_bf16_to_f32: out not 16-byte aligned");
  if (numel == 0) return TN_OK;
  int64_t grid = (numel / 8 + COLL_THREADS - 1) / COLL_THREADS;
  const int cap = max_ctas > 0 ? max_ctas : 32;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  reduce_bf16_to_f32_kernel<<<unsigned(grid), COLL_THREADS, 0, stream>>>(pp, n_inputs, out, numel, scale);
  TN_CHECK_CUDA(cudaGetLastError());
  return TN_OK;
}
