"""GPU parity: the CUDA audio frontend through the C ABI vs (a) the golden vectors produced by the reference functions
themselves and (b) the numpy oracle on seeded inputs at BASELINE sizes (30 s utterances).
fbank tolerance (power domain): |p - p_ref| <= 1e-4*p_ref + 1e-6*max_frame(p_ref)  (see tests/test_oracle_golden.py);
stack: atol/rtol 2e-4; log-mel: atol 2e-4."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from oracle import frontend_oracle as fo
from tests.gpu_util import require_cuda
from tests.test_oracle_golden import power_close
from touchnet_b200 import frontend as fe

pytestmark = pytest.mark.gpu


def test_fbank_matches_reference_golden(golden_dir):
    dev = require_cuda()
    g = np.load(os.path.join(golden_dir, "frontend_fbank.npz"))
    names = [k[6:] for k in g.files if k.startswith("fbank/")]
    # one launch over all utterances packed back to back (batched form) ...
    wavs = [torch.from_numpy(g["wav/" + n]) for n in names]
    flat, lens = fe._pack(wavs, dev)
    feats, frames = fe.fbank_batch(flat, lens)
    feats = feats.cpu().numpy()
    off = 0
    for n, m in zip(names, frames):
        ref = g["fbank/" + n]
        assert m == ref.shape[0]
        assert power_close(ref, feats[off:off + m]) < 1.0, n
        off += m
    # ... and the per-sample stage interface of the reference (f(data, config) -> data)
    cfg = NS(audiofeat_num_mel_bins=80, audiofeat_frame_length=25, audiofeat_frame_shift=10, audiofeat_dither=0.0)
    s = next(fe.audio_compute_fbank(iter([{"waveform": wavs[0][None], "sample_rate": 16000}]), cfg))
    assert power_close(g["fbank/" + names[0]], s["audiofeat"].cpu().numpy()) < 1.0


def test_fbank_int16_input_equals_float_input(golden_dir):
    dev = require_cuda()
    g = np.load(os.path.join(golden_dir, "frontend_fbank.npz"))
    name = [k[4:] for k in g.files if k.startswith("wav/real_")][0]
    wav = torch.from_numpy(g["wav/" + name])
    pcm = (wav * 32768.0).round().to(torch.int16)                  # what is stored on disk (ref: make_data.py:202)
    a, _ = fe.fbank_batch(pcm.to(dev), [pcm.numel()])
    b, _ = fe.fbank_batch((pcm.float() / 32768.0).to(dev), [pcm.numel()])
    assert torch.equal(a, b)                                       # int16 -> float is exact: bit-identical features


def test_stack_matches_reference_golden(golden_dir):
    dev = require_cuda()
    g = np.load(os.path.join(golden_dir, "frontend_fbank.npz"))
    keys = [k for k in g.files if k.startswith("stack/")]
    for k in keys:
        _, name, cfg = k.split("/")
        st, sd, nm = map(int, cfg.split("_"))
        x = torch.from_numpy(g["fbank/" + name]).to(dev)
        out, rows = fe.stack_batch(x, [x.shape[0]], st, sd, bool(nm))
        assert tuple(out.shape) == g[k].shape, k
        np.testing.assert_allclose(out.cpu().numpy(), g[k], rtol=2e-4, atol=2e-4, err_msg=k)
    cfg = NS(audiofeat_stack_length=5, audiofeat_stride_length=4, audiofeat_normalize=True)
    name = keys[0].split("/")[1]
    s = next(fe.audiofeat_stack(iter([{"audiofeat": torch.from_numpy(g["fbank/" + name])}]), cfg))
    np.testing.assert_allclose(s["audiofeat"].cpu().numpy(), g[f"stack/{name}/5_4_1"], rtol=2e-4, atol=2e-4)


def test_logmel_matches_torch_stft_golden(golden_dir):
    dev = require_cuda()
    g = np.load(os.path.join(golden_dir, "frontend_logmel.npz"))
    for k in [k for k in g.files if k.startswith("logmel/")]:
        _, name, nm = k.split("/")
        wav = torch.from_numpy(g["wav/" + name]).to(dev)
        out, frames = fe.log_mel_batch(wav, [wav.numel()], num_mel_bins=int(nm))
        assert tuple(out.shape) == g[k].shape
        np.testing.assert_allclose(out.cpu().numpy(), g[k], atol=2e-4, rtol=0, err_msg=k)


def test_frontend_full_size_batch_vs_oracle():
    """BASELINE-size utterances (up to 30 s, 16 kHz) in one packed batch, fbank -> stack(13,12) and (5,4)."""
    dev = require_cuda()
    g = torch.Generator().manual_seed(2025)
    lens = [int(16000 * d) for d in (30.0, 1.0, 12.3456, 0.0251, 7.77)]
    wavs = [0.3 * (2 * torch.rand(n, generator=g) - 1) for n in lens]
    flat, lens = fe._pack(wavs, dev)
    feats, frames = fe.fbank_batch(flat, lens)
    off = 0
    refs = []
    for w, m in zip(wavs, frames):
        ref = fo.fbank(w.numpy())
        refs.append(ref)
        assert ref.shape[0] == m
        assert power_close(ref, feats[off:off + m].cpu().numpy()) < 1.0
        off += m
    for st, sd in ((13, 12), (5, 4)):
        out, rows = fe.stack_batch(feats, frames, st, sd, True)
        off = 0
        for ref, r in zip(refs, rows):
            want = fo.stack(ref, st, sd, True)
            assert want.shape[0] == r
            np.testing.assert_allclose(out[off:off + r].cpu().numpy(), want, rtol=5e-3, atol=5e-3)
            off += r


def test_frontend_empty_and_too_short():
    dev = require_cuda()
    flat = torch.zeros(500, device=dev)
    feats, frames = fe.fbank_batch(flat, [100, 400])               # 100 samples < one window: zero frames
    assert frames == [0, 1] and feats.shape == (1, 80)
    out, rows = fe.stack_batch(feats, frames, 5, 4, False)
    assert rows == [0, 1] and out.shape == (1, 400)
