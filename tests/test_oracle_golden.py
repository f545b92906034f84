"""CPU: the oracle (oracle/*.py) against the golden vectors generated from the reference itself
(tests/golden/make_golden.py).  This is what pins the oracle; the GPU tests then compare the CUDA path with it."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import frontend_oracle as fo
from oracle import model_oracle as mo


def power_close(ref_log, out_log, rtol=1e-4, floor=1e-6):
    """fbank tolerance, stated in the power domain: |p - p_ref| <= rtol*p_ref + floor*max_frame(p_ref).
    (log-domain differences in bins 60 dB below the frame peak are FFT rounding noise in any fp32 implementation.)"""
    pr, po = np.exp(ref_log.astype(np.float64)), np.exp(out_log.astype(np.float64))
    lim = rtol * pr + floor * pr.max(axis=1, keepdims=True)
    return float((np.abs(pr - po) / lim).max())


def test_fbank_oracle_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "frontend_fbank.npz"))
    names = [k[6:] for k in g.files if k.startswith("fbank/")]
    assert len(names) >= 5
    for name in names:
        out = fo.fbank(g["wav/" + name])
        ref = g["fbank/" + name]
        assert out.shape == ref.shape, name
        assert power_close(ref, out) < 1.0, name


def test_stack_oracle_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "frontend_fbank.npz"))
    keys = [k for k in g.files if k.startswith("stack/")]
    assert len(keys) >= 30
    for k in keys:
        _, name, cfg = k.split("/")
        st, sd, nm = map(int, cfg.split("_"))
        out = fo.stack(g["fbank/" + name], st, sd, bool(nm))
        assert out.shape == g[k].shape, k
        np.testing.assert_allclose(out, g[k], rtol=2e-4, atol=2e-4, err_msg=k)


def test_logmel_oracle_matches_torch_stft_path(golden_dir):
    g = np.load(os.path.join(golden_dir, "frontend_logmel.npz"))
    for k in [k for k in g.files if k.startswith("logmel/")]:
        _, name, nm = k.split("/")
        out = fo.log_mel(g["wav/" + name], n_mels=int(nm))
        assert out.shape == g[k].shape
        np.testing.assert_allclose(out, g[k], atol=2e-4, rtol=0, err_msg=k)


def test_slaney_filters_pinned_to_whisper_mel_filter_bank(golden_dir):
    """The restated librosa.filters.mel equals the bank stored in the fixture, which was produced by
    transformers.audio_utils.mel_filter_bank(norm="slaney", mel_scale="slaney") - the function Whisper's feature
    extractor builds its filters from (an independent implementation of librosa's published algorithm)."""
    g = np.load(os.path.join(golden_dir, "frontend_logmel.npz"))
    for n_mels in (80, 128):
        ref = g[f"slaney/{n_mels}"]
        w = fo.slaney_mel_filters(16000, 400, n_mels)
        assert w.shape == ref.shape == (n_mels, 201)
        np.testing.assert_allclose(w, ref, rtol=0, atol=1e-7)
    try:        # and live, when transformers is importable on this box (it is not needed for the assertion above)
        from transformers.audio_utils import mel_filter_bank
    except Exception:
        return
    live = mel_filter_bank(num_frequency_bins=201, num_mel_filters=80, min_frequency=0.0, max_frequency=8000.0,
                           sampling_rate=16000, norm="slaney", mel_scale="slaney").T
    np.testing.assert_allclose(fo.slaney_mel_filters(16000, 400, 80), live, rtol=0, atol=1e-7)


def test_slaney_filters_have_whisper_properties():
    """Published properties of whisper's mel_filters - shape, Slaney area normalisation, triangular support,
    monotone centres."""
    for n_mels in (80, 128):
        w = fo.slaney_mel_filters(16000, 400, n_mels)
        assert w.shape == (n_mels, 201) and (w >= 0).all()
        centres = w.argmax(axis=1)
        assert (np.diff(centres) >= 0).all()
        # Slaney norm: each filter integrates to ~1 over Hz (bin width 40 Hz), away from the edges
        area = w.sum(axis=1) * 40.0
        if n_mels == 80:
            assert np.allclose(area[5:-1], 1.0, atol=0.15)
        else:  # 128 filters on 40 Hz bins: the low filters are narrower than a bin, only the mean area holds
            assert abs(float(area.mean()) - 1.0) < 0.05
    assert abs(float(fo.slaney_mel_filters(16000, 400, 80)[0].max()) - 0.02486) < 2e-3


def _cfg_from_json(cj, **kw):
    return mo.OracleConfig(hidden_size=cj["hidden_size"], intermediate_size=cj["intermediate_size"],
                           num_hidden_layers=cj["num_hidden_layers"], num_attention_heads=cj["num_attention_heads"],
                           num_key_value_heads=cj["num_key_value_heads"], head_dim=cj["head_dim"],
                           vocab_size=cj["vocab_size"], rms_norm_eps=cj["rms_norm_eps"], rope_theta=cj["rope_theta"],
                           rope_scaling=cj["rope_scaling"], tie_word_embeddings=cj["tie_word_embeddings"], **kw)


def test_model_oracle_matches_hf_llama_and_reference_touch_audio(golden_dir):
    g = np.load(os.path.join(golden_dir, "model_tiny.npz"))
    cj = json.loads(bytes(g["config_json"]).decode())
    cfg = _cfg_from_json(cj)
    inv, _ = mo.rope_inv_freq(cfg)
    assert torch.equal(inv, torch.from_numpy(g["llama/inv_freq"]))          # llama3 rope scaling, bit exact
    doc = torch.from_numpy(g["llama/doc_ids"])
    pos = torch.from_numpy(g["llama/position_ids"])
    params = {k[9:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("llama_sd/")}
    logits = mo.llama_forward(params, cfg, input_ids=torch.from_numpy(g["llama/input_ids"]), attention_mask=doc,
                              position_ids=pos)
    ref = torch.from_numpy(g["llama/logits"])
    valid = doc > 0   # padding query rows: HF eager is undefined there, FlexAttention yields zeros
    assert float((logits - ref)[valid].abs().max()) < 1e-5
    assert torch.equal(logits[valid].argmax(-1), ref[valid].argmax(-1))      # bit-exact token indices

    cfg_ta = _cfg_from_json(cj, audio_input_size=40)
    params = {k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("ta_sd/")}
    lg = mo.touch_audio_forward(params, cfg_ta, input_ids=torch.from_numpy(g["ta/input_ids"]),
                                input_features=torch.from_numpy(g["ta/input_features"]), attention_mask=doc,
                                position_ids=pos)
    ref = torch.from_numpy(g["ta/logits"])
    assert float((lg - ref)[valid].abs().max()) < 1e-5
    assert torch.equal(lg[valid].argmax(-1), ref[valid].argmax(-1))


def test_oracle_mask_semantics():
    doc = torch.tensor([[1, 1, 2, 2, 2, 0, 0]])
    allow = mo.doc_causal_allow(doc)[0]
    assert allow[1, 0] and allow[1, 1] and not allow[0, 1]        # causal inside a document
    assert not allow[2, 1] and allow[4, 2]                        # no attention across documents
    assert not allow[5].any() and not allow[6].any()              # padding queries fully masked
    assert not allow[:, 5].any()                                  # padding keys never attended
    q = torch.randn(1, 2, 7, 8); k = torch.randn(1, 1, 7, 8); v = torch.randn(1, 1, 7, 8)
    o, lse = mo.attention(q, k, v, mo.doc_causal_allow(doc), 8 ** -0.5)
    assert torch.all(o[0, 5:] == 0) and torch.isinf(lse[0, :, 5:]).all()   # FlexAttention: exact zeros


def test_pack_loss_oracle_has_the_invariance_the_reference_tests():
    """ref: tests/touchnet/utils/test_pack_loss.py - the loss of sentences trained as a padded batch (per-sentence token
    mean, then mean over sentences: calc_batch_dp_loss) equals the loss of the same sentences PACKED into one row and
    normalised through sentence_lens / num_sentence (calc_pack_sp_loss; touchnet/loss/cross_entropy.py:12-50).  The oracle
    restatement must have exactly that property (the CUDA loss is then compared with the oracle in tests/test_gpu_loss.py)."""
    import torch
    from oracle import model_oracle as mo
    g = torch.Generator().manual_seed(0)
    V, lens = 37, [5, 11, 3, 8]
    L = max(lens)
    logits_b = torch.randn(len(lens), L, V, generator=g)
    labels_b = torch.full((len(lens), L), -100, dtype=torch.int64)
    for i, n in enumerate(lens):
        labels_b[i, :n] = torch.randint(0, V, (n,), generator=g)
    # (a) the reference test's padded-batch formula
    ce = torch.nn.functional.cross_entropy(logits_b.reshape(-1, V), labels_b.reshape(-1), reduction="none", ignore_index=-100)
    batch_loss = (ce.reshape(len(lens), -1).sum(1) / ((labels_b != -100).sum(1).float() + 1e-12)).mean()
    # (b) the same sentences packed into one row of length 32 (5 pad positions at the end)
    T = 32
    logits_p = torch.randn(1, T, V, generator=g)
    labels_p = torch.full((1, T), -100, dtype=torch.int64)
    sl = torch.ones(1, T, dtype=torch.int64)
    o = 0
    for i, n in enumerate(lens):
        logits_p[0, o:o + n] = logits_b[i, :n]
        labels_p[0, o:o + n] = labels_b[i, :n]
        sl[0, o:o + n] = n
        o += n
    per_sample = mo.pack_loss(logits_p, labels_p, sl, len(lens))
    assert abs(float(per_sample) - float(batch_loss)) < 1e-6


def test_chunked_attention_oracle_equals_dense_oracle_and_its_autograd():
    """oracle.attention_chunked (what the full-size GPU parity tests compare against) is the same arithmetic as the dense
    `attention` pinned above, and its analytic backward equals autograd of the dense form (fp32, CPU, small case incl.
    padding, GQA and a query-chunk boundary inside a document)."""
    import math
    import torch
    from tests.gpu_util import packed_doc_ids
    torch.manual_seed(0)
    B, T, H, KV, hd = 2, 192, 4, 2, 16
    doc, _ = packed_doc_ids(B, T, [[50, 100, 30], [192]])
    q = torch.randn(B * T, H * hd); k = torch.randn(B * T, KV * hd); v = torch.randn(B * T, KV * hd)
    do = torch.randn(B * T, H * hd)
    sc = 1 / math.sqrt(hd)
    o, lse, dq, dk, dv = mo.attention_chunked(q, k, v, doc, H, KV, sc, do, q_chunk=64)
    qf = q.view(B, T, H, hd).transpose(1, 2).clone().requires_grad_(True)
    kf = k.view(B, T, KV, hd).transpose(1, 2).clone().requires_grad_(True)
    vf = v.view(B, T, KV, hd).transpose(1, 2).clone().requires_grad_(True)
    o_ref, lse_ref = mo.attention(qf, kf, vf, mo.doc_causal_allow(doc), sc)
    o_ref.backward(do.view(B, T, H, hd))
    fin = torch.isfinite(lse_ref)
    assert torch.equal(torch.isinf(lse), ~fin)
    assert float((lse - lse_ref)[fin].abs().max()) < 1e-5
    assert float((o - o_ref.reshape(B * T, -1)).abs().max()) < 1e-5
    tok = lambda g: g.transpose(1, 2).reshape(B * T, -1)
    for a, r in ((dq, tok(qf.grad)), (dk, tok(kf.grad)), (dv, tok(vf.grad))):
        assert float((a - r).abs().max()) < 2e-5
