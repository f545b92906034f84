"""GPU audio frontend behind the reference's data-stage interface.

The reference composes per-sample generator stages `f(data: Iterator[dict], config: DataConfig) -> Iterator[dict]`
(ref: touchnet/data/datapipe.py:183-205).  The three stages on the hot path are mirrored here with the same names,
argument meaning and dict keys:

    audio_compute_fbank                 ref: touchnet/data/functions.py:117-134
    audio_compute_log_mel_spectrogram   ref: touchnet/data/functions.py:159-190
    audiofeat_stack                     ref: touchnet/data/functions.py:258-286

plus batch-level entry points (`fbank_batch`, `log_mel_batch`, `stack_batch`, `fbank_stack_batch`) that featurise a
whole packed batch of utterances in one launch each - the form used after the H2D copy in the train step
(SURVEY 8(b) "data-frontend boundary", option (i)).  All arithmetic happens in libtouchnet_b200.so
(csrc/frontend.cu); windows and mel filter banks are small constants built once on the host exactly as the reference's
third-party code builds them and cached per device.
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import Iterable, Iterator, Sequence

import numpy as np
import torch

from . import _lib

MILLISECONDS_TO_SECONDS = 0.001


# ---------------------------------------------------------------------------------------------------------------
# constants (host-side, cached)
# ---------------------------------------------------------------------------------------------------------------
@lru_cache(maxsize=None)
def _povey_window(n: int, device: str) -> torch.Tensor:
    """ta: compliance/kaldi.py:99-101."""
    return torch.hann_window(n, periodic=False, dtype=torch.float32).pow(0.85).to(device)


@lru_cache(maxsize=None)
def _hann_window(n: int, device: str) -> torch.Tensor:
    """torch.hann_window(n) (periodic), ref: functions.py:171."""
    return torch.hann_window(n, dtype=torch.float32).to(device)


@lru_cache(maxsize=None)
def _kaldi_mel_banks(num_bins: int, n_fft: int, sample_freq: float, low_freq: float, high_freq: float, device: str):
    """ta: compliance/kaldi.py:436-511 get_mel_banks (no VTLN) + zero column of :622; same fp32 op order."""
    num_fft_bins = n_fft / 2
    nyquist = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyquist
    fft_bin_width = sample_freq / n_fft
    mel_low = 1127.0 * math.log(1.0 + low_freq / 700.0)
    mel_high = 1127.0 * math.log(1.0 + high_freq / 700.0)
    mel_delta = (mel_high - mel_low) / (num_bins + 1)
    b = torch.arange(num_bins).unsqueeze(1)
    left = mel_low + b * mel_delta
    center = mel_low + (b + 1.0) * mel_delta
    right = mel_low + (b + 2.0) * mel_delta
    mel = (1127.0 * (1.0 + (fft_bin_width * torch.arange(num_fft_bins)) / 700.0).log()).unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    bins = torch.max(torch.zeros(1), torch.min(up, down))
    return torch.nn.functional.pad(bins, (0, 1)).to(torch.float32).contiguous().to(device)


@lru_cache(maxsize=None)
def _slaney_mel_filters(sr: int, n_fft: int, n_mels: int, device: str) -> torch.Tensor:
    """librosa.filters.mel(sr=sr, n_fft=n_fft, n_mels=n_mels) defaults (Slaney scale + Slaney norm), ref: functions.py:180-183."""
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0

    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, f / f_sp)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    w = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        w[i] = np.maximum(0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    w *= (2.0 / (mel_f[2: n_mels + 2] - mel_f[:n_mels]))[:, None]
    return torch.from_numpy(w.astype(np.float32)).contiguous().to(device)


def _st() -> int:
    return torch.cuda.current_stream().cuda_stream


def _pack(waveforms: Sequence[torch.Tensor], device) -> tuple[torch.Tensor, list[int]]:
    lens = [int(w.numel()) for w in waveforms]
    flat = torch.cat([w.reshape(-1) for w in waveforms]) if len(waveforms) > 1 else waveforms[0].reshape(-1)
    return flat.to(device, non_blocking=True).contiguous(), lens


def _offsets(lens: Sequence[int], device) -> torch.Tensor:
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(np.asarray(lens, dtype=np.int64), out=off[1:])
    return torch.from_numpy(off).to(device, non_blocking=True)


# ---------------------------------------------------------------------------------------------------------------
# batch-level ops
# ---------------------------------------------------------------------------------------------------------------
def fbank_frames(n_samples: int, sample_rate: int, frame_length: float, frame_shift: float) -> int:
    win = int(sample_rate * frame_length * MILLISECONDS_TO_SECONDS)
    shift = int(sample_rate * frame_shift * MILLISECONDS_TO_SECONDS)
    return 0 if n_samples < win else 1 + (n_samples - win) // shift


def fbank_batch(wav: torch.Tensor, lens: Sequence[int], sample_rate: int = 16000, num_mel_bins: int = 80,
                frame_length: float = 25.0, frame_shift: float = 10.0, low_freq: float = 20.0, high_freq: float = 0.0,
                preemphasis: float = 0.97):
    """wav: packed utterances back to back on the GPU, fp32 in [-1,1] or int16 PCM.  Returns (feats [sum m_i, mel]
    fp32, frame_counts).  Semantics of kaldi.fbank(waveform * 2**15, dither=0, energy_floor=0) per utterance."""
    if not wav.is_cuda:
        raise _lib.TouchNetB200Error("fbank_batch needs the packed waveform on the GPU (no CPU path)")
    assert wav.dtype in (torch.float32, torch.int16) and wav.dim() == 1
    dev = str(wav.device)
    win = int(sample_rate * frame_length * MILLISECONDS_TO_SECONDS)
    shift = int(sample_rate * frame_shift * MILLISECONDS_TO_SECONDS)
    n_fft = 1 << (win - 1).bit_length()
    frames = [0 if n < win else 1 + (n - win) // shift for n in lens]
    total = int(sum(frames))
    out = torch.empty((total, num_mel_bins), dtype=torch.float32, device=wav.device)
    if total == 0:
        return out, frames
    utt_off, frm_off = _offsets(lens, wav.device), _offsets(frames, wav.device)
    window = _povey_window(win, dev)
    banks = _kaldi_mel_banks(num_mel_bins, n_fft, float(sample_rate), float(low_freq), float(high_freq), dev)
    _lib.call("tn_fbank_f32", wav.data_ptr(), int(wav.dtype == torch.int16), utt_off.data_ptr(), frm_off.data_ptr(),
              len(lens), total, win, shift, n_fft, window.data_ptr(), banks.data_ptr(), num_mel_bins, float(preemphasis),
              out.data_ptr(), _st())
    return out, frames


def log_mel_batch(wav: torch.Tensor, lens: Sequence[int], sample_rate: int = 16000, n_fft: int = 400,
                  hop_length: int = 160, num_mel_bins: int = 80):
    """Whisper-style log-mel per utterance (ref: functions.py:159-190).  Returns (feats [sum N_i//hop, mel], counts)."""
    if not wav.is_cuda:
        raise _lib.TouchNetB200Error("log_mel_batch needs the packed waveform on the GPU (no CPU path)")
    assert wav.dtype == torch.float32 and wav.dim() == 1
    dev = str(wav.device)
    frames = [n // hop_length for n in lens]       # 1 + N//hop stft frames, last one dropped (:176)
    total = int(sum(frames))
    out = torch.empty((total, num_mel_bins), dtype=torch.float32, device=wav.device)
    if total == 0:
        return out, frames
    utt_off, frm_off = _offsets(lens, wav.device), _offsets(frames, wav.device)
    utt_max = torch.empty(len(lens), dtype=torch.float32, device=wav.device)
    window = _hann_window(n_fft, dev)
    filters = _slaney_mel_filters(sample_rate, n_fft, num_mel_bins, dev)
    _lib.call("tn_logmel_power_f32", wav.data_ptr(), utt_off.data_ptr(), frm_off.data_ptr(), len(lens), total, n_fft,
              hop_length, window.data_ptr(), filters.data_ptr(), num_mel_bins, out.data_ptr(), utt_max.data_ptr(), _st())
    _lib.call("tn_logmel_finish_f32", out.data_ptr(), frm_off.data_ptr(), utt_max.data_ptr(), len(lens), total,
              num_mel_bins, _st())
    return out, frames


def stack_batch(feats: torch.Tensor, frames: Sequence[int], stack_length: int, stride_length: int,
                normalize: bool = True, into: torch.Tensor | None = None, dst_rows: Sequence[int] | None = None):
    """Low-frame-rate stacking per utterance (ref: functions.py:258-286).  Returns (out [sum ceil(T_i/stride), mel*stack], counts).
    With `into` ([rows, F] fp32 batch buffer) and `dst_rows` (first destination row per utterance) the stacked rows are
    written straight into the packed-batch `input_features` buffer (layout of processing_touch_audio.py:200)."""
    assert feats.is_cuda and feats.dtype == torch.float32 and feats.dim() == 2
    n_mels = feats.shape[1]
    rows = [int(math.ceil(t / stride_length)) if t > 0 else 0 for t in frames]
    total = int(sum(rows))
    if into is None:
        out = torch.empty((total, n_mels * stack_length), dtype=torch.float32, device=feats.device)
        dst, ld = None, 0
    else:
        assert into.is_cuda and into.dtype == torch.float32 and into.dim() == 2 and into.stride(1) == 1
        assert into.shape[1] == n_mels * stack_length and dst_rows is not None and len(dst_rows) == len(frames)
        out, ld = into, into.stride(0)
        dst = torch.tensor(list(dst_rows), dtype=torch.int64).to(feats.device, non_blocking=True)
    if total == 0:
        return out, rows
    frm_off, out_off = _offsets(frames, feats.device), _offsets(rows, feats.device)
    _lib.call("tn_feat_stack_f32", feats.data_ptr(), frm_off.data_ptr(), out_off.data_ptr(), len(frames), total, n_mels,
              stack_length, stride_length, int(bool(normalize)), out.data_ptr(), None if dst is None else dst.data_ptr(),
              ld, _st())
    return out, rows


def fbank_stack_batch(wav: torch.Tensor, lens: Sequence[int], *, sample_rate=16000, num_mel_bins=80, frame_length=25.0,
                      frame_shift=10.0, stack_length=5, stride_length=4, normalize=True):
    feats, frames = fbank_batch(wav, lens, sample_rate, num_mel_bins, frame_length, frame_shift)
    return stack_batch(feats, frames, stack_length, stride_length, normalize)


# ---------------------------------------------------------------------------------------------------------------
# per-sample generator stages (the reference's operator interface)
# ---------------------------------------------------------------------------------------------------------------
def _device(config) -> torch.device:
    return torch.device(getattr(config, "frontend_device", "cuda"))


def audio_compute_fbank(data: Iterable[dict], config) -> Iterator[dict]:
    """Extract fbank.  Same contract as ref: touchnet/data/functions.py:117-134."""
    for sample in data:
        assert "sample_rate" in sample
        assert "waveform" in sample
        wav, lens = _pack([sample["waveform"]], _device(config))
        mat, _ = fbank_batch(wav.float() if wav.dtype not in (torch.float32, torch.int16) else wav, lens,
                             sample_rate=sample["sample_rate"], num_mel_bins=config.audiofeat_num_mel_bins,
                             frame_length=config.audiofeat_frame_length, frame_shift=config.audiofeat_frame_shift)
        if getattr(config, "audiofeat_dither", 0.0) != 0.0:
            raise _lib.TouchNetB200Error("dither != 0 is not supported (all reference recipes use 0.0)")
        sample["audiofeat"] = mat
        yield sample


def audio_compute_log_mel_spectrogram(data: Iterable[dict], config) -> Iterator[dict]:
    """Extract whisper-style log mel spectrogram.  Same contract as ref: touchnet/data/functions.py:159-190."""
    for sample in data:
        assert "sample_rate" in sample
        assert "waveform" in sample
        waveform = sample["waveform"].squeeze(0)
        pad = getattr(config, "audiofeat_padding", 0)
        if pad > 0:
            waveform = torch.nn.functional.pad(waveform, (0, pad))
        wav, lens = _pack([waveform], _device(config))
        mat, _ = log_mel_batch(wav.float(), lens, sample_rate=sample["sample_rate"], n_fft=config.audiofeat_n_fft,
                               hop_length=config.audiofeat_hop_length, num_mel_bins=config.audiofeat_num_mel_bins)
        sample["audiofeat"] = mat
        yield sample


def audiofeat_stack(data: Iterable[dict], config) -> Iterator[dict]:
    """Stack audio features (low frame rate).  Same contract as ref: touchnet/data/functions.py:258-286."""
    for sample in data:
        assert "audiofeat" in sample
        x = sample["audiofeat"]
        if not x.is_cuda:
            x = x.to(_device(config))
        out, _ = stack_batch(x.float().contiguous(), [x.shape[0]], config.audiofeat_stack_length,
                             config.audiofeat_stride_length, config.audiofeat_normalize)
        sample["audiofeat"] = out
        yield sample
