/* touchnet_b200 — C ABI of the B200-native TouchNet hot path.
 *
 * This header is the drop-in boundary: every entry point replaces one third-party kernel call site that the
 * reference (xingchensong/TouchNet) reaches from Python.  "Reference interface" below names the reference
 * file:line the call replaces (paths relative to the TouchNet checkout; `hf:` = transformers 4.51.3,
 * `ta:` = torchaudio).
 *
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller (PyTorch caching allocator); kernels never allocate
 *     and never synchronise; work is enqueued on `stream` (a cudaStream_t / CUstream; NULL = legacy default).
 *   - bf16 tensors are row-major with explicit leading dimensions (in elements).
 *   - return value 0 = success; non-zero = error, message via tn_last_error() (thread-local).
 *   - entry points are re-entrant and stream-ordered (FSDP2 side streams, activation-checkpoint recompute and the
 *     autograd device thread may all call concurrently on different streams).
 */
#ifndef TOUCHNET_B200_H
#define TOUCHNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TOUCHNET_B200_VERSION 100

typedef void* tn_stream_t; /* cudaStream_t */

const char* tn_last_error(void);
int tn_version(void);
/* 0 iff the current CUDA device is compute capability 10.x (B200). */
int tn_device_check(void);
/* Persistent kernels (the GEMMs) launch one CTA per SM; under FSDP2 the NCCL all-gather / reduce-scatter kernels of the
 * neighbouring layers need SMs of their own to overlap with them.  Leave `sms` SMs unused by persistent grids
 * (0 = default, single GPU).  Process-wide. */
int tn_set_sm_margin(int sms);
/* Tuning knob: raster group of the CTA-pair GEMM = number of 256-row M blocks whose tiles are walked across all N blocks
 * before moving on (default 8; decides which operand strips stay L2-resident between waves).  Results do not depend on it. */
int tn_set_gemm_group(int m_blocks);
int tn_set_gemm_split_tail(int on);  /* 1 (default): a last wave that would leave over half of the CTA pairs idle runs as 256x128 half tiles */
int tn_set_gemm_l2_hints(int on);   /* 1: A strips evict-last, B strips evict-first in the CTA-pair GEMM TMA loads (default 0: measured 6 % slower) */

/* ---------------------------------------------------------------------------------------------------------------
 * GEMM  D[M,N] = A·Bᵀ (+ R)          tcgen05 / TMEM / TMA, bf16 in, fp32 accumulate.
 *   a_mn = 0: A stored [M,K] (K contiguous, lda)      a_mn = 1: A stored [K,M] (M contiguous, lda)
 *   b_mn = 0: B stored [N,K] (K contiguous, ldb)      b_mn = 1: B stored [K,N] (N contiguous, ldb)
 *   d_f32 = 0: D (and R) bf16;  d_f32 = 1: D (and R) fp32.  R optional (NULL), may alias D (accumulate).
 *   forward  y = x·Wᵀ      : a_mn=0, b_mn=0   (F.linear; hf:models/llama/modeling_llama.py:251-289 q/k/v/o_proj,
 *                                               :182-184 down_proj; touchnet/models/touch_audio/modeling_touch_audio.py:127 projector)
 *   dgrad    dx = dy·W     : a_mn=0, b_mn=1
 *   wgrad    dW = dyᵀ·x    : a_mn=1, b_mn=1
 */
int tn_gemm_bf16(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* D, int64_t ldd,
                 int d_f32, const void* R, int64_t ldr, int M, int N, int K, tn_stream_t stream);

/* Fused gate/up projection + SwiGLU:  G = X·Wgᵀ, U = X·Wuᵀ, H = silu(G)⊙U   (all bf16, [M,N]; X [M,K]; Wg,Wu [N,K]).
 * G and U may be NULL (inference: only H is written).
 * Reference interface: hf:models/llama/modeling_llama.py:182-184 LlamaMLP.forward (gate_proj, up_proj, act_fn, mul). */
int tn_gemm_swiglu_bf16(const void* X, int64_t ldx, const void* Wg, const void* Wu, int64_t ldw, void* G, void* U,
                        void* H, int64_t ldh, int M, int N, int K, tn_stream_t stream);

/* q/k/v projections as ONE launch although the three weights stay separate parameters (HF FQNs, FSDP2, DCP):
 *   mode 0 forward  D0[M, s0+s1+s2] = A[M,K] · [B0;B1;B2]ᵀ         (B_i [s_i, K]; q|k|v land side by side in one buffer)
 *   mode 1 dgrad    D0[M, N]        = A[M, s0+s1+s2] · [B0;B1;B2]   (A = dq|dk|dv side by side; B_i [s_i, N])
 *   mode 2 wgrad    D_i[s_i, N]     = A[:, seg_i]ᵀ · B0[Mred=K, N]  (A = dq|dk|dv [K, s0+s1+s2]; B0 = layer input x)
 * Segment sizes must be multiples of 256.  hf:models/llama/modeling_llama.py:251-289 (q_proj, k_proj, v_proj).
 * mode 0 with rope_cos/rope_sin ([M, 64] bf16 tables of tn_rope_table): the epilogue also applies RoPE to the q and k
 * segments (hf apply_rotary_pos_emb :151-168) with the rounding points of the unfused bf16 ops (bit-identical). */
/* Down-proj dgrad fused with the SwiGLU backward (hf LlamaMLP.forward modeling_llama.py:182-184, backward of
 * `down(silu(gate) * up)`): dH = dY[M,K] . W[K,N] (W = down_proj.weight, [d, ffn] row-major) never leaves the chip -
 * the epilogue turns the accumulator tile into dG = (dH*U)*silu'(G) and dU = dH*silu(G) with the rounding points of
 * tn_gemm_bf16 + tn_swiglu_bwd_bf16 (bit-identical to that two-kernel path).  Needs M, N >= 256. */
int tn_gemm_dswiglu_bf16(const void* dY, int64_t lddy, const void* W, int64_t ldw, const void* G, const void* U,
                         int64_t ldgu, void* dG, void* dU, int64_t lddg, int M, int N, int K, tn_stream_t stream);
int tn_gemm_qkv_bf16(int mode, const void* A, int64_t lda, const void* B0, const void* B1, const void* B2, int64_t ldb,
                     void* D0, void* D1, void* D2, int64_t ldd, int d_f32, int s0, int s1, int s2, int M, int N, int K,
                     const void* rope_cos, const void* rope_sin, tn_stream_t stream);

/* SwiGLU backward (elementwise): dG = dH⊙U⊙silu'(G), dU = dH⊙silu(G).  bf16 [M,N] contiguous rows (ld). */
int tn_swiglu_bwd_bf16(const void* G, const void* U, const void* dH, void* dG, void* dU, int64_t rows, int64_t cols,
                       int64_t ld, tn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * RMSNorm (+ fused residual add).   hf:models/llama/modeling_llama.py:62-67 LlamaRMSNorm.forward and the residual
 * adds of LlamaDecoderLayer.forward :303-333.
 *   fwd: if R != NULL: S = X + R (bf16, written to S_out if non-NULL) else S = X;
 *        Y = w ⊙ bf16(S · rsqrt(mean(S²)+eps));  rstd[row] (fp32) saved for backward.
 *   bwd: dS = rstd·(w⊙dY) − rstd³/d · S · Σ(w⊙dY⊙S)  (+ dS_extra if non-NULL: gradient flowing through the residual
 *        branch), dW_partial[num_partials, d] fp32 partial sums (workspace; num_partials = tn_rmsnorm_bwd_num_partials())
 *        and, if dW != NULL, dW[d] = their column sums in a fixed order (a second, tiny kernel on the same stream).
 */
int tn_rmsnorm_fwd_bf16(const void* X, const void* R, const void* w, int w_is_f32, void* S_out, void* Y, float* rstd,
                        int64_t rows, int d, float eps, tn_stream_t stream);
int tn_rmsnorm_bwd_bf16(const void* S, const void* dY, const void* dS_extra, const void* w, int w_is_f32,
                        const float* rstd, void* dS, float* dW_partial, int num_partials, float* dW, int64_t rows, int d,
                        tn_stream_t stream);
int tn_rmsnorm_bwd_num_partials(void);

/* ---------------------------------------------------------------------------------------------------------------
 * RoPE.  hf:models/llama/modeling_llama.py:124-168 (LlamaRotaryEmbedding.forward + apply_rotary_pos_emb), called at
 * touchnet/models/llama/pipeline_llama.py:83.  position_ids restart per packed document
 * (touchnet/models/llama/processing_llama.py:98).
 *   tn_rope_table: cos/sin[b*T+t, j] = bf16(cos/sin(position_ids[b,t] · inv_freq[j]) · attention_scaling), j < hd/2
 *   tn_rope_apply: in-place x ← x·cos + rotate_half(x)·sin on X [rows, n_heads, hd] (row stride ldx elements);
 *                  inverse != 0 applies the transpose rotation (backward).
 */
int tn_rope_table(const int64_t* position_ids, const float* inv_freq, float attention_scaling, void* cos_out,
                  void* sin_out, int64_t rows, int half_dim, tn_stream_t stream);
int tn_rope_apply_bf16(void* X, int64_t ldx, const void* cos_tab, const void* sin_tab, int64_t rows, int n_heads,
                       int head_dim, int inverse, tn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Packed-sequence (block-causal "document") attention, head_dim = 128, bf16, GQA.
 * Reference interface: hf:integrations/flex_attention.py:136-247 make_flex_block_causal_mask + :262-364
 * flex_attention_forward (selected by "attn_implementation": "flex_attention",
 * examples/text/pretrain/fineweb-edu/config/Llama-3_2-1B.json:7; checked at touchnet/bin/train.py:129-131).
 *   mask: allow[b,q,k] = (q >= k) && doc[b,q] == doc[b,k] && doc[b,q] > 0      (doc = attention_mask ids, int32)
 *   Q [B,T,H,128], K/V [B,T,KV,128], O [B,T,H,128] with strides (elements): token stride ld*, batch stride = T*ld*.
 *   lse [B,H,T] fp32 (natural log; +inf for fully-masked rows whose O is exactly 0).
 *   Context parallelism (touchnet/utils/distributed.py:292-315 shards the sequence over the `cp` mesh): Tq > 0 makes
 *   Q / O / dO / dQ / lse / delta hold only the Tq query rows starting at global position q_blk_off*128, while K / V /
 *   dK / dV / doc_ids / meta stay global (T rows); Tq <= 0 means Tq = T, offset 0.
 *   tn_attn_prep builds, on the device with no host sync, the per-block kv/q ranges, a per-row "canonical" flag and
 *   the per-position document extents [start,end) the kernels mask with; meta: int32 buffer of tn_attn_meta_ints(B,T)
 *   elements.
 */
int64_t tn_attn_meta_ints(int B, int T);
int tn_attn_prep(const int32_t* doc_ids, int32_t* meta, int B, int T, tn_stream_t stream);
int tn_attn_fwd_bf16(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, void* O,
                     int64_t ldo, float* lse, const int32_t* doc_ids, const int32_t* meta, int B, int T, int H, int KV,
                     float scale, int Tq, int q_blk_off, tn_stream_t stream);
/* backward: delta [B,H,T] fp32 workspace; dQ [B,T,H,128], dK/dV [B,T,KV,128] bf16.  Optional fused inverse RoPE: with
 * rope_cos_q/sin_q ([B*Tq, 64] bf16 tables of tn_rope_table) dQ is returned w.r.t. the UN-rotated q, likewise
 * rope_cos_k/sin_k ([B*T, 64]) for dK (backward of hf apply_rotary_pos_emb, modeling_llama.py:151-168); NULL = off. */
int tn_attn_bwd_bf16(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, const void* O,
                     int64_t ldo, const void* dO, int64_t lddo, const float* lse, float* delta, void* dQ, int64_t lddq,
                     void* dK, int64_t lddk, void* dV, int64_t lddv, const int32_t* doc_ids, const int32_t* meta, int B,
                     int T, int H, int KV, float scale, int Tq, int q_blk_off, const void* rope_cos_q,
                     const void* rope_sin_q, const void* rope_cos_k, const void* rope_sin_k, tn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Audio frontend.
 *   tn_fbank: touchnet/data/functions.py:117-134 audio_compute_fbank → ta:compliance/kaldi.py:514-645 (snip_edges,
 *     remove DC, pre-emphasis, povey window, |rfft|², HTK mel bank, log) batched over utterances packed back to back.
 *     wav: fp32 in [-1,1] (wav_is_i16 = 0; scaled by 2^15 like functions.py:125) or int16 PCM (wav_is_i16 = 1).
 *     utt_offsets[n_utts+1]: sample offsets; frame_offsets[n_utts+1]: output row offsets (m_i = 1+(N_i-win)/shift).
 *     window [frame_len] and mel_filters [n_mels, n_fft/2+1] are supplied by the host mirror (povey window and HTK
 *     mel bank computed exactly as ta:compliance/kaldi.py:90-103 and :436-511 do).
 *   tn_logmel: touchnet/data/functions.py:159-190 audio_compute_log_mel_spectrogram (whisper-style: hann STFT,
 *     center/reflect, power, Slaney mel (caller supplies filters [n_mels, n_fft/2+1] fp32), log10, per-utterance
 *     max-8 clamp, (x+4)/4).  Two passes: tn_logmel_power writes log10 mel + per-utterance max (atomic), then
 *     tn_logmel_finish applies the clamp/affine in place.
 *   tn_feat_stack: touchnet/data/functions.py:258-286 audiofeat_stack (left/right replicate padding, stack/stride,
 *     optional per-row mean / unbiased-std normalisation).  dst_row0 == NULL: rows are written packed back to back;
 *     otherwise utterance u's rows go to out[(dst_row0[u] + i) * out_ld ...], i.e. straight into the
 *     input_features [B*T, F] buffer at the positions batch_pairaudio_pairtext_packed assigns
 *     (touchnet/models/touch_audio/processing_touch_audio.py:200).
 */
int tn_fbank_f32(const void* wav, int wav_is_i16, const int64_t* utt_offsets, const int64_t* frame_offsets, int n_utts,
                 int64_t total_frames, int frame_len, int frame_shift, int n_fft, const float* window,
                 const float* mel_filters, int n_mels, float preemph, float* out, tn_stream_t stream);
int tn_logmel_power_f32(const float* wav, const int64_t* utt_offsets, const int64_t* frame_offsets, int n_utts,
                        int64_t total_frames, int n_fft, int hop, const float* window, const float* mel_filters,
                        int n_mels, float* out, float* utt_max, tn_stream_t stream);
int tn_logmel_finish_f32(float* feats, const int64_t* frame_offsets, const float* utt_max, int n_utts,
                         int64_t total_frames, int n_mels, tn_stream_t stream);
int tn_feat_stack_f32(const float* feats, const int64_t* frame_offsets, const int64_t* out_offsets, int n_utts,
                      int64_t total_out_rows, int n_mels, int stack, int stride, int normalize, float* out,
                      const int64_t* dst_row0, int64_t out_ld, tn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Small fused elementwise helpers on the path.
 *   tn_embed_add: E = embed[input_ids] + P   (touchnet/models/touch_audio/modeling_touch_audio.py:124-131), with the
 *     NaN check of :133-134 folded in (nan_flag[0] set to 1 if any NaN; read lazily by the caller, no sync here).
 *   tn_cast_f32_bf16: fp32 master → bf16 working copy (what FSDP2 MixedPrecisionPolicy does at all-gather,
 *     touchnet/models/helper_func.py:163).
 */
int tn_embed_add_bf16(const int64_t* input_ids, const void* embed, int embed_is_f32, const void* P, void* E,
                      int32_t* nan_flag, int64_t rows, int d, int64_t vocab, tn_stream_t stream);
int tn_cast_f32_bf16(const float* src, void* dst, int64_t n, tn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Pack-loss cross-entropy next to the path.  touchnet/loss/__init__.py:7-28 +
 * touchnet/loss/cross_entropy.py:12-50 + touchnet/utils/metrics.py:26-50, without fp32 logits:
 *   fwd: lse[r] = logsumexp(logits[r,:]) (fp32), ce[r] = lse[r] - logits[r,label[r]] (0 where label is the ignore
 *        index, i.e. outside [0,V)), argmax[r] (optional, for the accuracy metric).
 *   bwd: IN PLACE  logits[r,:] <- (softmax(logits[r,:]) - onehot(label[r])) * scale * grad_scalar[0] / sentence_lens[r]
 *        (rows with an ignored label become 0).  grad_scalar: device pointer to the upstream scalar gradient (or NULL = 1).
 */
int tn_pack_ce_fwd_bf16(const void* logits, int64_t ld, const int64_t* labels, float* lse, float* ce, int32_t* argmax,
                        int64_t M, int V, tn_stream_t stream);
int tn_pack_ce_bwd_bf16(void* logits, int64_t ld, const int64_t* labels, const int64_t* sentence_lens, const float* lse,
                        const float* grad_scalar, float scale, int64_t M, int V, tn_stream_t stream);
/* vocabulary-parallel backward (tensor-parallel lm_head with loss parallel, ref: touchnet/utils/distributed.py:322-323,
 * touchnet/models/llama/parallelize_llama.py:127-131): `logits` holds columns [v0, v0+V_local) of a V_total-wide vocabulary,
 * labels are GLOBAL ids, lse the GLOBAL logsumexp (all-reduced by the caller from per-shard tn_pack_ce_fwd_bf16 results). */
int tn_pack_ce_bwd_vp_bf16(void* logits, int64_t ld, const int64_t* labels_global, const int64_t* sentence_lens,
                           const float* lse_global, const float* grad_scalar, float scale, int64_t M, int V_local, int64_t v0,
                           int64_t V_total, tn_stream_t stream);
/* fwd and bwd (upstream gradient 1) of the above in ONE launch over a chunk of rows: the building block of the fused
 * lm_head + loss (touchnet_b200/loss.py::FusedLinearCEFn - the reference's best path is Liger's fused-linear-cross-entropy,
 * touchnet/bin/train.py:443-445, which never materialises [B,T,V] logits either).  IN PLACE like the bwd entry point. */
int tn_pack_ce_fused_bf16(void* logits, int64_t ld, const int64_t* labels, const int64_t* sentence_lens, float* lse,
                          float* ce, int32_t* argmax, float scale, int64_t M, int V, tn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * BEST-RQ tokenizer: labels of audio pre-training.  touchnet/tokenizer/tokenizer.py:289-299.
 *   codes[t] = argmin_v || normalize(feats[t,:] @ proj) - codebook[v,:] ||_2   (lowest index on ties, like torch.argmin)
 *   feats [T, D] fp32 (row stride ld), proj [D, E], codebook [V, E] (rows already L2-normalised, :267), E in {8,16,32}.
 */
int tn_bestrq_tokenize_f32(const float* feats, int64_t ld, const float* proj, const float* codebook, int64_t T, int D,
                           int E, int V, int32_t* codes, tn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Optimizer step next to the path, on the rank-local fp32 shards.
 *   tn_sumsq_f32: partials[0..*n_partials_used) = block-wise sums of x[i]^2 (deterministic; the caller adds them up, sums
 *     over tensors / ranks and takes the root) - replaces torch.nn.utils.get_total_norm in
 *     touchnet/utils/distributed.py:426-491.  `partials` holds tn_sumsq_num_partials() floats.
 *   tn_scale_f32: x *= scale[0] (device scalar, no-op when it is 1) - clip_grads_with_norm_ of the same function.
 *   tn_adamw_f32: one AdamW step (decoupled weight decay, bias corrections passed in) replacing
 *     torch.optim.AdamW(..., fused=True) as built by touchnet/utils/optimizer.py:127-172; `grad_scale` (device scalar or
 *     NULL) multiplies the gradient on the fly (the clip coefficient, so clipping needs no pass of its own) and
 *     `param_bf16` (or NULL) receives the bf16 working copy of the updated parameter.
 */
int tn_sumsq_num_partials(void);
int tn_sumsq_f32(const float* x, int64_t n, float* partials, int* n_partials_used, tn_stream_t stream);
int tn_scale_f32(float* x, int64_t n, const float* scale, tn_stream_t stream);
int tn_scale_bf16(void* x, int64_t n, const float* scale, tn_stream_t stream);   /* same, bf16 data (no-op when scale is 1) */
int tn_adamw_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* param_bf16, int64_t n,
                 float lr, float beta1, float beta2, float eps, float weight_decay, float bias_correction1,
                 float bias_correction2_sqrt, const float* grad_scale, tn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Device-side assembly of the integer side of a packed batch: the per-document slice assignments of
 * touchnet/models/llama/processing_llama.py:64-102 and touchnet/models/touch_audio/processing_touch_audio.py:176-206.
 * The host keeps the greedy placement (row / offset / document id per document) and ships the concatenated text tokens;
 * the five [B,T] int64 buffers are filled here: defaults pad / -100 / 0 / 0 / 1, then per document
 *   position_ids = 0..total-1, attention_mask = sid, sentence_lens = n_txt over its `audio + n_txt` positions,
 *   input_ids = [bos, tokens...], labels = [tokens..., eos] over the text positions (n_txt = n_tokens + 1).
 */
int tn_pack_layout_i64(const int32_t* doc_row, const int32_t* doc_off, const int32_t* doc_audio, const int32_t* doc_sid,
                       const int64_t* tok_off, const int64_t* tokens, int n_docs, int B, int T, int64_t pad, int64_t bos,
                       int64_t eos, int64_t* input_ids, int64_t* labels, int64_t* position_ids, int64_t* attention_mask,
                       int64_t* sentence_lens, tn_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Data-parallel collectives over NVLink peer memory (SURVEY 8(e) "NCCL's or your own"; touchnet_b200/fsdp_comm.py).  What FSDP2 does
 * through NCCL around every decoder block (touchnet/models/helper_func.py:134-202: fp32 gradient reduce-scatter, bf16
 * parameter all-gather) as pull kernels on buffers every rank has mapped (torch symmetric memory):
 *   tn_peer_reduce_scatter_f32: out[i] = scale * sum_{p<n_peers} inputs[p][shard_offset + i], i < numel, summed in rank
 *     order (bit-reproducible); inputs = HOST array of n_peers device pointers (peer-mapped), shard_offset % 4 == 0.
 *   tn_peer_all_gather: out[p*bytes_each + j] = inputs[p][j]; bytes_each % 16 == 0.
 * max_ctas bounds the grid (default 32) so that the GEMMs these overlap with keep their SMs.  Cross-rank ordering (inputs
 * complete before the call, buffers not reused before every rank's call has finished) is the caller's barrier.
 */
int tn_peer_reduce_scatter_f32(const void* const* inputs, int n_peers, int64_t shard_offset, float* out, int64_t numel,
                               float scale, int max_ctas, tn_stream_t stream);
int tn_peer_all_gather(const void* const* inputs, int n_peers, int64_t bytes_each, void* out, int max_ctas,
                       tn_stream_t stream);
/* out[i] = scale * sum_k float(inputs[k][i]): bf16 chunks (local memory), fp32 accumulation in input order - the reduce step
 * of the direct-push reduce-scatter of touchnet_b200/fsdp_comm.py (same sum as FSDP2's fp32 reduce of bf16 gradients). */
int tn_reduce_bf16_to_f32(const void* const* inputs, int n_inputs, float* out, int64_t numel, float scale, int max_ctas,
                          tn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TOUCHNET_B200_H */
