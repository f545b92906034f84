// touchnet_b200 :: BEST-RQ random-projection tokenizer on the GPU.
//
// Replaces BestRQTokenizer.tokenize  touchnet/tokenizer/tokenizer.py:289-299:
//     xs = inputs @ quantizer            [T, D] @ [D, E]
//     xs = F.normalize(xs, dim=-1, p=2, eps=1e-8)
//     codes = vector_norm(xs[:, None, :] - codebook[None], dim=-1).argmin(-1)      (the [T, V, E] broadcast on CPU)
// which produces the LABELS of audio pre-training (touchnet/models/touch_audio/processing_touch_audio.py:100-104).
// One warp per frame: lanes split D for the projection, then split the V code words for the nearest-neighbour search;
// argmin keeps the lowest index among equal distances like torch.argmin.  Integer output -> parity is index equality.
#include <math.h>

#include "../../include/touchnet_b200.h"
#include "common.cuh"
#include "host.h"

namespace tn {


template <int E>
__global__ void __launch_bounds__(256) bestrq_kernel(const float* __restrict__ feats, int64_t ld,
                                                     const float* __restrict__ proj,      // [D, E]
                                                     const float* __restrict__ codebook,  // [V, E], rows L2-normalised
                                                     int64_t T, int D, int V, int32_t* __restrict__ codes) {
  const int64_t t = int64_t(blockIdx.x) * 8 + warp_id();
  if (t >= T) return;
  const uint32_t lane = lane_id();
  const float* x = feats + t * ld;
  float acc[E];
#pragma unroll
  for (int e = 0; e < E; ++e) acc[e] = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float xv = x[d];
    const float* pr = proj + int64_t(d) * E;
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = fmaf(xv, __ldg(pr + e), acc[e]);
  }
  float nrm2 = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    acc[e] = warp_sum(acc[e]);
    nrm2 += acc[e] * acc[e];
  }
  const float inv = 1.f / fmaxf(sqrtf(nrm2), 1e-8f);   // F.normalize(p=2, eps=1e-8)
#pragma unroll
  for (int e = 0; e < E; ++e) acc[e] *= inv;
  float best = INFINITY;
  int best_i = 0x7fffffff;
  for (int v = lane; v < V; v += 32) {
    const float* c = codebook + int64_t(v) * E;
    float d2 = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const float df = acc[e] - __ldg(c + e);
      d2 = fmaf(df, df, d2);
    }
    const float dist = sqrtf(d2);   // the reference compares norms, not squared norms
    if (dist < best) { best = dist; best_i = v; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ob < best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
  }
  if (lane == 0) codes[t] = best_i;
}

}  // namespace tn

using namespace tn;

extern "C" int tn_bestrq_tokenize_f32(const float* feats, int64_t ld, const float* proj, const float* codebook, int64_t T,
                                      int D, int E, int V, int32_t* codes, tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(feats && proj && codebook && codes, "tn_bestrq_tokenize_f32: null pointer");
  TN_REQUIRE(D > 0 && V > 0 && ld >= D, "tn_bestrq_tokenize_f32: bad dims D=%d V=%d ld=%lld", D, V, (long long)ld);
  TN_REQUIRE(E == 16 || E == 32 || E == 8, "tn_bestrq_tokenize_f32: emb_size %d not instantiated (8, 16, 32)", E);
  if (T == 0) return TN_OK;
  const unsigned grid = unsigned((T + 7) / 8);
  if (E == 16) bestrq_kernel<16><<<grid, 256, 0, stream>>>(feats, ld, proj, codebook, T, D, V, codes);
  else if (E == 32) bestrq_kernel<32><<<grid, 256, 0, stream>>>(feats, ld, proj, codebook, T, D, V, codes);
  else bestrq_kernel<8><<<grid, 256, 0, stream>>>(feats, ld, proj, codebook, T, D, V, codes);
  TN_CHECK_CUDA(cudaGetLastError());
  return TN_OK;
}
