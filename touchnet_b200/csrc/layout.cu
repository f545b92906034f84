// touchnet_b200 :: device-side assembly of the integer side of a packed batch.
//
// Replaces the per-document slice assignments of
//   touchnet/models/llama/processing_llama.py:24-104                 (batch_text)
//   touchnet/models/touch_audio/processing_touch_audio.py:117-214    (batch_pairaudio_pairtext_packed)
// for batches whose placement (row, offset) was decided by the same greedy first-fit-in-order rule on the host: instead of
// five [B,T] int64 buffers, the host ships the concatenated text tokens and five small per-document integers; the
// buffers are filled here.  Integer work: bit-exact against the host batchers (tests/test_gpu_layout.py).
#include "../../include/touchnet_b200.h"
#include "common.cuh"
#include "host.h"

namespace tn {

struct LayoutParams {
  const int32_t* doc_row;       // [n_docs] batch row
  const int32_t* doc_off;       // [n_docs] first position inside the row
  const int32_t* doc_audio;     // [n_docs] audio positions in front of the text (0 for text-only documents)
  const int32_t* doc_sid;       // [n_docs] document id inside its row (1, 2, ...)
  const int64_t* tok_off;       // [n_docs + 1] offsets into `tokens`
  const int64_t* tokens;        // concatenated text tokens of all documents (without bos / eos)
  int n_docs, B, T;
  int64_t pad, bos, eos;
  int64_t *input_ids, *labels, *position_ids, *attention_mask, *sentence_lens;
};

__global__ void layout_fill_kernel(const LayoutParams p) {
  const int64_t n = int64_t(p.B) * p.T;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    p.input_ids[i] = p.pad;
    p.labels[i] = -100;            // ignore_idx of the reference
    p.position_ids[i] = 0;
    p.attention_mask[i] = 0;
    p.sentence_lens[i] = 1;
  }
}

__global__ void layout_docs_kernel(const LayoutParams p) {
  const int doc = blockIdx.x;
  const int a = p.doc_audio[doc];
  const int64_t t0 = p.tok_off[doc], n_tok = p.tok_off[doc + 1] - t0;
  const int64_t n_txt = n_tok + 1;                       // + bos (inputs) / + eos (labels)
  const int64_t total = a + n_txt;
  const int64_t base = int64_t(p.doc_row[doc]) * p.T + p.doc_off[doc];
  const int64_t sid = p.doc_sid[doc];
  for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
    const int64_t o = base + i;
    p.position_ids[o] = i;
    p.attention_mask[o] = sid;
    p.sentence_lens[o] = n_txt;
    if (i >= a) {
      const int64_t j = i - a;                            // 0 .. n_tok
      p.input_ids[o] = j == 0 ? p.bos : p.tokens[t0 + j - 1];
      p.labels[o] = j == n_tok ? p.eos : p.tokens[t0 + j];
    }
  }
}

}  // namespace tn

using namespace tn;

extern "C" int tn_pack_layout_i64(const int32_t* doc_row, const int32_t* doc_off, const int32_t* doc_audio,
                                  const int32_t* doc_sid, const int64_t* tok_off, const int64_t* tokens, int n_docs, int B,
                                  int T, int64_t pad, int64_t bos, int64_t eos, int64_t* input_ids, int64_t* labels,
                                  int64_t* position_ids, int64_t* attention_mask, int64_t* sentence_lens,
                                  tn_stream_t stream_) {
  clear_error();
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TN_REQUIRE(input_ids && labels && position_ids && attention_mask && sentence_lens, "tn_pack_layout_i64: null output");
  TN_REQUIRE(B > 0 && T > 0 && n_docs >= 0, "tn_pack_layout_i64: B=%d T=%d n_docs=%d", B, T, n_docs);
  TN_REQUIRE(n_docs == 0 || (doc_row && doc_off && doc_audio && doc_sid && tok_off && tokens),
             "tn_pack_layout_i64: null document table");
  LayoutParams p{doc_row, doc_off, doc_audio, doc_sid, tok_off, tokens, n_docs, B, T, pad, bos, eos,
                 input_ids, labels, position_ids, attention_mask, sentence_lens};
  const int64_t n = int64_t(B) * T;
  int64_t grid = (n + 255) / 256;
  if (grid > int64_t(sm_count()) * 8) grid = int64_t(sm_count()) * 8;
  layout_fill_kernel<<<unsigned(grid), 256, 0, stream>>>(p);
  TN_CHECK_CUDA(cudaGetLastError());
  if (n_docs > 0) {
    layout_docs_kernel<<<unsigned(n_docs), 256, 0, stream>>>(p);
    TN_CHECK_CUDA(cudaGetLastError());
  }
  return TN_OK;
}
