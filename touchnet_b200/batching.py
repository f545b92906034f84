"""Packed-batch layouts of the hot path's inputs + the synthetic workload generator used by tests and bench.py.

The batch producers are host-side Python in the reference and stay host-side here; they are restated (not replaced)
because they DEFINE the input layout of the path:

    batch_text                        ref: touchnet/models/llama/processing_llama.py:24-104
    batch_pairaudio_pairtext_packed   ref: touchnet/models/touch_audio/processing_touch_audio.py:117-214

Same generator signatures `f(data, config, tokenizer) -> Iterator[dict]`, same buffer keys, dtypes and the same greedy
in-order packing rule.  tests/test_batching.py pins them to the reference's own outputs (tests/golden/batching.npz).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Iterable, Iterator, Optional

import torch


def _text_buffer(config, tokenizer, with_features: Optional[int] = None) -> dict:
    B, T = config.dataset_batchsize, config.dataset_text_seqlen
    buf = {
        "input_ids": torch.zeros([B, T], dtype=torch.int64) + tokenizer.pad,
        "labels": torch.zeros([B, T], dtype=torch.int64) - 100,      # ignore_idx = -100
        "position_ids": torch.zeros([B, T], dtype=torch.int64),
        "attention_mask": torch.zeros([B, T], dtype=torch.int64),     # document ids from 1; 0 = padding
        "sentence_lens": torch.ones([B, T], dtype=torch.int64),
        "num_sentence": 0,
    }
    if with_features is None:
        buf["inputs_embeds"] = None
    else:
        buf["input_features"] = torch.zeros([B, T, with_features], dtype=torch.float32)
    return buf


def batch_text(data: Iterable[dict], config, tokenizer) -> Iterator[dict]:
    """Greedy in-order sequence packing of tokenised documents into fixed [B, T] buffers."""
    buffer = _text_buffer(config, tokenizer)
    b = t = 0
    sid = 1
    for sample in data:
        n = len(sample["input_ids"]) + 1  # +1 for sos/eos
        if b == config.dataset_batchsize - 1:
            if t + n > config.dataset_text_seqlen:
                yield buffer
                buffer = _text_buffer(config, tokenizer)
                b = t = 0
                sid = 1
        else:
            if t + n > config.dataset_text_seqlen:
                b += 1
                t = 0
                sid = 1
        buffer["input_ids"][b, t:t + n] = torch.tensor([tokenizer.bos] + list(sample["input_ids"]), dtype=torch.int64)
        buffer["labels"][b, t:t + n] = torch.tensor(list(sample["input_ids"]) + [tokenizer.eos], dtype=torch.int64)
        buffer["position_ids"][b, t:t + n] = torch.arange(0, n, dtype=torch.int64)
        buffer["attention_mask"][b, t:t + n] = sid
        buffer["sentence_lens"][b, t:t + n] = n
        buffer["num_sentence"] += 1
        t += n
        sid += 1
    if (not config.dataloader_drop_last_batch) and (t > 0 or b > 0):
        yield buffer


def batch_pairaudio_pairtext_packed(data: Iterable[dict], config, tokenizer) -> Iterator[dict]:
    """Audio frames then <bos>text per document, packed; labels only on text; features zero on text positions."""
    assert config.dataset_audio_seqlen == config.dataset_text_seqlen
    F = config.audiofeat_num_mel_bins * config.audiofeat_stack_length
    T = config.dataset_audio_seqlen
    buffer = _text_buffer(config, tokenizer, with_features=F)
    b = t = 0
    sid = 1
    for sample in data:
        a = sample["audiofeat"].size(0)
        n_txt = len(sample["input_ids"]) + 1
        total = a + n_txt
        if total > T:
            continue
        if b == config.dataset_batchsize - 1:
            if t + total > T:
                buffer["shift_labels"] = buffer["labels"]
                yield buffer
                buffer = _text_buffer(config, tokenizer, with_features=F)
                b = t = 0
                sid = 1
        else:
            if t + total > T:
                b += 1
                t = 0
                sid = 1
        buffer["input_features"][b, t:t + a] = sample["audiofeat"].to("cpu", torch.float32)
        buffer["input_ids"][b, t + a:t + total] = torch.tensor([tokenizer.bos] + list(sample["input_ids"]), dtype=torch.int64)
        buffer["labels"][b, t + a:t + total] = torch.tensor(list(sample["input_ids"]) + [tokenizer.eos], dtype=torch.int64)
        buffer["position_ids"][b, t:t + total] = torch.arange(0, total, dtype=torch.int64)
        buffer["attention_mask"][b, t:t + total] = sid
        buffer["sentence_lens"][b, t:t + total] = n_txt
        buffer["num_sentence"] += 1
        t += total
        sid += 1
    if (not config.dataloader_drop_last_batch) and (b > 0 or t > 0):
        buffer["shift_labels"] = buffer["labels"]
        yield buffer


# ---------------------------------------------------------------------------------------------------------------
# synthetic workload: seeded, reproduces the layouts above exactly
# ---------------------------------------------------------------------------------------------------------------
SYN_TOKENIZER = SimpleNamespace(pad=0, bos=1, eos=2)


def synthetic_text_samples(seed: int, vocab: int, T: int, mu_len: float = 600.0, sigma: float = 1.0) -> Iterator[dict]:
    """Documents with LogNormal(ln mu, sigma) lengths clamped to [2, T-1], tokens ~ U[3, vocab)."""
    g = torch.Generator().manual_seed(seed)
    while True:
        n = int(round(math.exp(math.log(mu_len) + sigma * float(torch.randn((), generator=g)))))
        n = max(2, min(n, T - 1))
        yield {"input_ids": torch.randint(3, vocab, (n,), generator=g).tolist()}


def synthetic_utterances(seed: int, vocab: int, sample_rate: int = 16000, min_s: float = 1.0, max_s: float = 30.0,
                         stride: int = 4, frame_shift_ms: float = 10.0, frame_len_ms: float = 25.0,
                         int16: bool = False) -> Iterator[dict]:
    """Utterances d ~ U[min_s, max_s], waveform 0.3*U(-1,1); text length ~ a * U[0.05, 0.25] for a feature frames."""
    g = torch.Generator().manual_seed(seed)
    win, shift = int(sample_rate * frame_len_ms / 1000), int(sample_rate * frame_shift_ms / 1000)
    while True:
        d = min_s + (max_s - min_s) * float(torch.rand((), generator=g))
        n = int(sample_rate * d)
        wav = 0.3 * (2 * torch.rand(n, generator=g) - 1)
        if int16:
            wav = (wav * 32768.0).round().clamp(-32768, 32767).to(torch.int16)
        m = 1 + (n - win) // shift
        a = int(math.ceil(m / stride))
        t = max(1, int(round(a * (0.05 + 0.20 * float(torch.rand((), generator=g))))))
        yield {"waveform": wav, "sample_rate": sample_rate, "input_ids": torch.randint(3, vocab, (t,), generator=g).tolist(),
               "num_feat_frames": a}


def take_batch(gen: Iterator[dict]) -> dict:
    return next(gen)


def make_text_batch(seed: int, B: int, T: int, vocab: int) -> dict:
    cfg = SimpleNamespace(dataset_batchsize=B, dataset_text_seqlen=T, dataloader_drop_last_batch=True)
    return next(batch_text(synthetic_text_samples(seed, vocab, T), cfg, SYN_TOKENIZER))


def plan_audio_text_batch(seed: int, B: int, T: int, vocab: int, stride: int = 4, max_s: float = 30.0,
                          int16: bool = False) -> tuple[dict, list[dict]]:
    """Layout of one packed audio+text batch WITHOUT the features: returns (buffer with zero input_features of width 0,
    list of placed utterances with (row, offset, n_frames)) so that the GPU frontend can fill input_features from raw
    waveforms after the H2D copy (SURVEY 8(f) row 3 direction; layout identical to batch_pairaudio_pairtext_packed)."""
    placed = []
    tok = SYN_TOKENIZER
    buf = _text_buffer(SimpleNamespace(dataset_batchsize=B, dataset_text_seqlen=T), tok, with_features=0)
    b = t = 0
    sid = 1
    for s in synthetic_utterances(seed, vocab, stride=stride, max_s=max_s, int16=int16):
        a, n_txt = s["num_feat_frames"], len(s["input_ids"]) + 1
        total = a + n_txt
        if total > T:
            continue
        if t + total > T:
            if b == B - 1:
                break
            b += 1
            t = 0
            sid = 1
        buf["input_ids"][b, t + a:t + total] = torch.tensor([tok.bos] + s["input_ids"], dtype=torch.int64)
        buf["labels"][b, t + a:t + total] = torch.tensor(s["input_ids"] + [tok.eos], dtype=torch.int64)
        buf["position_ids"][b, t:t + total] = torch.arange(0, total, dtype=torch.int64)
        buf["attention_mask"][b, t:t + total] = sid
        buf["sentence_lens"][b, t:t + total] = n_txt
        buf["num_sentence"] += 1
        placed.append({"waveform": s["waveform"], "row": b, "offset": t, "frames": a})
        t += total
        sid += 1
    buf["shift_labels"] = buf["labels"]
    return buf, placed


# ---------------------------------------------------------------------------------------------------------------
# device-side assembly of the integer buffers
# ---------------------------------------------------------------------------------------------------------------
def plan_documents(samples: Iterable[dict], B: int, T: int, with_audio: bool) -> dict:
    """The greedy first-fit-in-order placement of `batch_text` / `batch_pairaudio_pairtext_packed` for ONE batch, as compact
    per-document tables (host side, integers only): row, offset, audio positions, document id, token offsets + tokens.
    Consumes `samples` until the batch is full (the sample that does not fit is dropped, like the synthetic planner)."""
    rows, offs, auds, sids, tok_off, toks = [], [], [], [], [0], []
    b = t = 0
    sid = 1
    for s in samples:
        a = int(s["num_feat_frames"] if "num_feat_frames" in s else (s["audiofeat"].size(0) if with_audio else 0)) if with_audio else 0
        n_txt = len(s["input_ids"]) + 1
        total = a + n_txt
        if total > T:
            continue
        if t + total > T:
            if b == B - 1:
                break
            b += 1
            t = 0
            sid = 1
        rows.append(b); offs.append(t); auds.append(a); sids.append(sid)
        toks.extend(int(x) for x in s["input_ids"])
        tok_off.append(len(toks))
        t += total
        sid += 1
    i32 = lambda v: torch.tensor(v, dtype=torch.int32)
    return {"doc_row": i32(rows), "doc_off": i32(offs), "doc_audio": i32(auds), "doc_sid": i32(sids),
            "tok_off": torch.tensor(tok_off, dtype=torch.int64), "tokens": torch.tensor(toks or [0], dtype=torch.int64),
            "num_sentence": len(rows), "B": B, "T": T}


def assemble_on_device(plan: dict, device, tokenizer=SYN_TOKENIZER) -> dict:
    """Fill input_ids / labels / position_ids / attention_mask / sentence_lens [B,T] int64 ON THE GPU from a
    `plan_documents` table (H2D: the tokens + 5 small vectors instead of five [B,T] buffers).  Same keys as the host
    batchers' buffers (ref: processing_llama.py:24-104, processing_touch_audio.py:117-214)."""
    from . import _lib, ops
    B, T = plan["B"], plan["T"]
    dev = torch.device(device)
    d = {k: plan[k].to(dev, non_blocking=True) for k in ("doc_row", "doc_off", "doc_audio", "doc_sid", "tok_off", "tokens")}
    out = {k: torch.empty((B, T), dtype=torch.int64, device=dev)
           for k in ("input_ids", "labels", "position_ids", "attention_mask", "sentence_lens")}
    ops._chk(out["input_ids"], "output buffers")
    n_docs = int(plan["doc_row"].numel())
    _lib.call("tn_pack_layout_i64", d["doc_row"].data_ptr(), d["doc_off"].data_ptr(), d["doc_audio"].data_ptr(),
              d["doc_sid"].data_ptr(), d["tok_off"].data_ptr(), d["tokens"].data_ptr() if n_docs else None, n_docs, B, T,
              int(tokenizer.pad), int(tokenizer.bos), int(tokenizer.eos), out["input_ids"].data_ptr(),
              out["labels"].data_ptr(), out["position_ids"].data_ptr(), out["attention_mask"].data_ptr(),
              out["sentence_lens"].data_ptr(), ops._st())
    out["num_sentence"] = plan["num_sentence"]
    out["shift_labels"] = out["labels"]
    return out
