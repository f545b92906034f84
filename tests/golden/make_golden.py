"""Generates the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the authoring container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py

What is executed to produce each file
  frontend_fbank.npz   reference functions touchnet.data.functions.audio_compute_fbank / audiofeat_stack
                       (imported per-module from /root/reference; they call the installed torchaudio kaldi.fbank)
                       on seeded noise/chirp waveforms and on the two real wavs of the reference's own tests
                       (tests/assets/dataset/*.wav, PCM parsed with the stdlib `wave` module).
  frontend_logmel.npz  the torch.stft path of touchnet/data/functions.py:168-189 executed line by line with torch,
                       with librosa.filters.mel (absent here) supplied by an INDEPENDENT third-party implementation of
                       the same published algorithm: transformers.audio_utils.mel_filter_bank(norm="slaney",
                       mel_scale="slaney") - the function WhisperFeatureExtractor builds its `mel_filters` from. The
                       bank itself is stored too (`slaney/<n_mels>`), so the oracle's restatement is pinned against it.
  model_tiny.npz       installed HF LlamaForCausalLM (eager attention, fp32) on the reference's test config
                       tests/assets/config/tiny_llama.json with a dense 4-D document mask, and the reference's own
                       TouchAudioForCausalLM wrapper (per-module import) around the same text config.
  batching.npz         reference batchers batch_text / batch_pairaudio_pairtext_packed on synthetic samples.
  bestrq.npz           reference BestRQTokenizer (random projection + nearest random code) on seeded features.
"""
import importlib
import json
import os
import sys
import types
import wave

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def import_reference():
    """Per-module import recipe of SURVEY.md 8(c): stub package root + stub librosa + stub sox_utils."""
    # probe (and cache) transformers' optional-dependency checks BEFORE the librosa stub exists
    import transformers  # noqa: F401
    from transformers.utils import import_utils as _iu
    if hasattr(_iu, "is_librosa_available"):
        _iu.is_librosa_available()
    import transformers.audio_utils  # noqa: F401
    pkg = types.ModuleType("touchnet")
    pkg.__path__ = [os.path.join(REF, "touchnet")]
    sys.modules["touchnet"] = pkg
    if "librosa" not in sys.modules:
        lib = types.ModuleType("librosa")
        lib.__spec__ = importlib.machinery.ModuleSpec("librosa", None)
        lib.filters = types.SimpleNamespace(mel=None)
        sys.modules["librosa"] = lib
    import torchaudio
    if not hasattr(torchaudio, "utils") or not hasattr(torchaudio.utils, "sox_utils"):
        utils = getattr(torchaudio, "utils", types.ModuleType("torchaudio.utils"))
        utils.sox_utils = types.SimpleNamespace(set_buffer_size=lambda *_: None)
        torchaudio.utils = utils
    functions = importlib.import_module("touchnet.data.functions")
    # the stub must not survive: transformers probes `librosa` availability and would then import soxr
    if getattr(sys.modules.get("librosa"), "__file__", None) is None:
        del sys.modules["librosa"]
    return functions


def read_wav(path):
    with wave.open(path, "rb") as w:
        assert w.getsampwidth() == 2
        sr = w.getframerate()
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16)
        if w.getnchannels() > 1:
            pcm = pcm.reshape(-1, w.getnchannels())[:, 0]
    return pcm.copy(), sr


def make_waveforms():
    g = torch.Generator().manual_seed(2025)
    out = {}
    out["noise_1s"] = (0.3 * (2 * torch.rand(16000, generator=g) - 1)).numpy().astype(np.float32)
    out["noise_odd"] = (0.3 * (2 * torch.rand(16000 + 137, generator=g) - 1)).numpy().astype(np.float32)
    t = np.arange(24000, dtype=np.float64) / 16000.0
    out["chirp"] = (0.5 * np.sin(2 * np.pi * (100.0 + 2000.0 * t) * t)).astype(np.float32)
    out["short"] = (0.1 * torch.randn(400 + 160 * 3 + 5, generator=g)).numpy().astype(np.float32)
    ds = os.path.join(REF, "tests/assets/dataset")
    if os.path.isdir(ds):
        for f in sorted(os.listdir(ds)):
            if f.endswith(".wav"):
                pcm, sr = read_wav(os.path.join(ds, f))
                if sr == 16000:
                    out["real_" + f[:-4].replace("-", "_")] = (pcm[: 16000 * 2].astype(np.float32) / 32768.0)
    return out


def golden_frontend(functions):
    from types import SimpleNamespace as NS
    out = {}
    wavs = make_waveforms()
    for name, wav in wavs.items():
        out[f"wav/{name}"] = wav
        cfg = NS(audiofeat_num_mel_bins=80, audiofeat_frame_length=25, audiofeat_frame_shift=10, audiofeat_dither=0.0)
        sample = {"waveform": torch.from_numpy(wav)[None], "sample_rate": 16000}
        fb = next(functions.audio_compute_fbank(iter([dict(sample)]), cfg))["audiofeat"]
        out[f"fbank/{name}"] = fb.numpy()
        for (st, sd) in [(5, 4), (7, 6), (13, 12)]:
            for norm in (True, False):
                c2 = NS(audiofeat_stack_length=st, audiofeat_stride_length=sd, audiofeat_normalize=norm)
                try:
                    stacked = next(functions.audiofeat_stack(iter([{"audiofeat": fb.clone()}]), c2))["audiofeat"]
                    out[f"stack/{name}/{st}_{sd}_{int(norm)}"] = stacked.numpy()
                except Exception as e:  # reference itself raises on some (T, stack, stride): record that
                    out[f"stack_err/{name}/{st}_{sd}_{int(norm)}"] = np.array([1])
    np.savez_compressed(os.path.join(HERE, "frontend_fbank.npz"), **out)
    print("frontend_fbank.npz", {k: v.shape for k, v in out.items() if k.startswith("fbank/")})

    # log-mel: functions.py:168-189; librosa.filters.mel(sr=16000, n_fft=400, n_mels) comes from transformers'
    # mel_filter_bank (Slaney scale + Slaney norm, fmin 0, fmax sr/2 = librosa's defaults), NOT from our restatement
    from transformers.audio_utils import mel_filter_bank

    def slaney_mel_filters(sr, n_fft, n_mels):
        return mel_filter_bank(num_frequency_bins=1 + n_fft // 2, num_mel_filters=n_mels, min_frequency=0.0,
                               max_frequency=sr / 2.0, sampling_rate=sr, norm="slaney",
                               mel_scale="slaney").T.astype(np.float32)
    out = {}
    for n_mels in (80, 128):
        out[f"slaney/{n_mels}"] = slaney_mel_filters(16000, 400, n_mels)
    for name in ("noise_1s", "chirp", "noise_odd"):
        wav = torch.from_numpy(wavs[name])
        for n_mels in (80, 128):
            window = torch.hann_window(400)
            stft = torch.stft(wav, 400, 160, window=window, return_complex=True)
            magnitudes = stft[..., :-1].abs() ** 2
            filters = torch.from_numpy(slaney_mel_filters(16000, 400, n_mels))
            mel_spec = filters @ magnitudes
            log_spec = torch.clamp(mel_spec, min=1e-10).log10()
            log_spec = torch.maximum(log_spec, log_spec.max() - 8.0)
            log_spec = (log_spec + 4.0) / 4.0
            out[f"logmel/{name}/{n_mels}"] = log_spec.transpose(0, 1).contiguous().numpy()
        out[f"wav/{name}"] = wavs[name]
    np.savez_compressed(os.path.join(HERE, "frontend_logmel.npz"), **out)
    print("frontend_logmel.npz", {k: v.shape for k, v in out.items() if k.startswith("logmel/")})


def doc_layout(B, T, lens_per_row):
    doc = torch.zeros(B, T, dtype=torch.int64)
    pos = torch.zeros(B, T, dtype=torch.int64)
    for b, lens in enumerate(lens_per_row):
        o = 0
        for i, n in enumerate(lens):
            doc[b, o:o + n] = i + 1
            pos[b, o:o + n] = torch.arange(n)
            o += n
    return doc, pos


def golden_model():
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg_json = json.load(open(os.path.join(REF, "tests/assets/config/tiny_llama.json")))
    out = {"config_json": np.frombuffer(json.dumps(cfg_json).encode(), dtype=np.uint8)}
    cfg = LlamaConfig(**{k: v for k, v in cfg_json.items() if k not in ("architectures", "model_type", "torch_dtype")})
    cfg._attn_implementation = "eager"
    torch.manual_seed(2025)
    model = LlamaForCausalLM(cfg).float().eval()
    # larger-than-default init so that logits are not all ~0
    with torch.no_grad():
        for n_, p_ in model.named_parameters():
            if p_.dim() == 2:
                p_.normal_(0, 0.2)
    B, T = 2, 48
    doc, pos = doc_layout(B, T, [[7, 20, 13], [30, 11]])   # row 0: 8 pad positions; row 1: 7 pad positions
    g = torch.Generator().manual_seed(7)
    ids = torch.randint(0, cfg.vocab_size, (B, T), generator=g)
    idx = torch.arange(T)
    allow = (idx[:, None] >= idx[None, :])[None] & (doc[:, :, None] == doc[:, None, :]) & (doc > 0)[:, :, None]
    mask4d = torch.zeros(B, 1, T, T).masked_fill(~allow[:, None], torch.finfo(torch.float32).min)
    with torch.no_grad():
        logits = model(input_ids=ids, attention_mask=mask4d, position_ids=pos).logits
    sd = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    for k, v in sd.items():
        out["llama_sd/" + k] = v
    out["llama/input_ids"] = ids.numpy(); out["llama/doc_ids"] = doc.numpy(); out["llama/position_ids"] = pos.numpy()
    out["llama/logits"] = logits.numpy()
    out["llama/inv_freq"] = model.model.rotary_emb.inv_freq.numpy()

    # reference TouchAudioForCausalLM wrapper (per-module import), same text config, F = 40 audio features
    import importlib
    mta = importlib.import_module("touchnet.models.touch_audio.modeling_touch_audio")
    cta = importlib.import_module("touchnet.models.touch_audio.configuration_touch_audio")
    text_cfg = dict(cfg_json); text_cfg["model_type"] = "llama"; text_cfg.pop("architectures", None)
    text_cfg["torch_dtype"] = "float32"
    tacfg = cta.TouchAudioConfig(audio_config={"model_type": "touch_audio_projector", "input_size": 40},
                                 text_config=text_cfg, pad_token_id=0)
    tacfg.text_config._attn_implementation = "eager"
    tacfg._attn_implementation = "eager"
    torch.manual_seed(2026)
    tam = mta.TouchAudioForCausalLM(tacfg).float().eval()
    with torch.no_grad():
        for n_, p_ in tam.named_parameters():
            if p_.dim() == 2:
                p_.normal_(0, 0.2)
    feats = torch.randn(B, T, 40, generator=g)
    # audio positions carry features and pad ids; text positions carry ids and zero features (processing_touch_audio.py:200-208)
    is_audio = torch.zeros(B, T, dtype=torch.bool)
    is_audio[0, :4] = True; is_audio[0, 7:19] = True; is_audio[1, :22] = True; is_audio[1, 30:36] = True
    feats = feats * is_audio[..., None]
    ids2 = torch.where(is_audio, torch.zeros_like(ids), ids)
    with torch.no_grad():
        o = tam(input_ids=ids2, input_features=feats, attention_mask=mask4d, position_ids=pos)
    for k, v in tam.state_dict().items():
        out["ta_sd/" + k] = v.detach().numpy()
    out["ta/input_ids"] = ids2.numpy(); out["ta/input_features"] = feats.numpy(); out["ta/logits"] = o.logits.numpy()
    np.savez_compressed(os.path.join(HERE, "model_tiny.npz"), **out)
    print("model_tiny.npz logits", logits.shape, o.logits.shape)


def golden_batching():
    import importlib
    from types import SimpleNamespace as NS
    pl = importlib.import_module("touchnet.models.llama.processing_llama")
    pta = importlib.import_module("touchnet.models.touch_audio.processing_touch_audio")
    tok = NS(pad=0, bos=1, eos=2)
    g = torch.Generator().manual_seed(11)
    lens = [5, 9, 3, 14, 7, 2, 11, 6, 8, 4, 10, 13]
    samples = [{"input_ids": torch.randint(3, 50, (n,), generator=g).tolist()} for n in lens]
    cfg = NS(dataset_batchsize=2, dataset_text_seqlen=32, dataloader_drop_last_batch=False)
    out = {"text/lens": np.array(lens)}
    for i, s in enumerate(samples):
        out[f"text/sample{i}"] = np.array(s["input_ids"])
    for bi, batch in enumerate(pl.batch_text(iter(samples), cfg, tok)):
        for k, v in batch.items():
            if isinstance(v, torch.Tensor):
                out[f"text/batch{bi}/{k}"] = v.numpy()
            elif isinstance(v, int):
                out[f"text/batch{bi}/{k}"] = np.array(v)
    # audio + text packed
    alens = [6, 11, 4, 9, 13, 5]
    tlens = [3, 5, 2, 6, 4, 3]
    F_ = 8
    asamples = [{"audiofeat": torch.randn(a, F_, generator=g), "input_ids": torch.randint(3, 50, (t,), generator=g).tolist()}
                for a, t in zip(alens, tlens)]
    cfg2 = NS(dataset_batchsize=2, dataset_audio_seqlen=32, dataset_text_seqlen=32, audiofeat_num_mel_bins=F_,
              audiofeat_stack_length=1, dataloader_drop_last_batch=False)
    out["at/alens"] = np.array(alens); out["at/tlens"] = np.array(tlens)
    for i, s in enumerate(asamples):
        out[f"at/feat{i}"] = s["audiofeat"].numpy(); out[f"at/ids{i}"] = np.array(s["input_ids"])
    for bi, batch in enumerate(pta.batch_pairaudio_pairtext_packed(iter(asamples), cfg2, tok)):
        for k, v in batch.items():
            if isinstance(v, torch.Tensor):
                out[f"at/batch{bi}/{k}"] = v.numpy()
            elif isinstance(v, int):
                out[f"at/batch{bi}/{k}"] = np.array(v)
    np.savez_compressed(os.path.join(HERE, "batching.npz"), **out)
    print("batching.npz", [k for k in out if k.endswith("attention_mask")])


def golden_bestrq():
    """reference BestRQTokenizer (touchnet/tokenizer/tokenizer.py:236-318) on seeded stacked-fbank-like features."""
    import importlib
    from types import SimpleNamespace as NS
    tk = importlib.import_module("touchnet.tokenizer.tokenizer")
    out = {}
    g = torch.Generator().manual_seed(77)
    for tag, (D, E, V, T) in {"default": (560, 16, 8192, 128), "small": (400, 16, 1024, 129)}.items():
        cfg = NS(tokenizer_bestrq_vocab_size=V, tokenizer_bestrq_input_size=D, tokenizer_bestrq_emb_size=E,
                 tokenizer_bestrq_init_seed=2026, tokenizer_bestrq_init_method="default")
        tok = tk.BestRQTokenizer(cfg)
        feats = torch.randn(T, D, generator=g)
        codes = tok.tokenize(feats)
        out[f"{tag}/feats"] = feats.numpy()
        out[f"{tag}/codes"] = np.array(codes, dtype=np.int64)
        # the random tensors are reproducible from the seed; keep a fingerprint instead of 0.5 MB of floats
        out[f"{tag}/quantizer_head"] = tok._quantizer.detach().numpy()[:4]
        out[f"{tag}/codebook_head"] = tok._codebook.detach().numpy()[:4]
        out[f"{tag}/sums"] = np.array([float(tok._quantizer.double().sum()), float(tok._codebook.double().sum())])
        out[f"{tag}/cfg"] = np.array([D, E, V, 2026])
    np.savez_compressed(os.path.join(HERE, "bestrq.npz"), **out)
    print("bestrq.npz", {k: v.shape for k, v in out.items() if k.endswith("codes")})


if __name__ == "__main__":
    functions = import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "frontend":
        golden_frontend(functions)
        sys.exit(0)
    golden_frontend(functions)
    golden_batching()
    golden_bestrq()
    golden_model()
