"""CPU: the restated batch producers against the reference's own outputs (tests/golden/batching.npz)."""
import os
from types import SimpleNamespace as NS

import numpy as np
import torch

from touchnet_b200 import batching


def _collect(g, prefix):
    out = {}
    for k in g.files:
        if k.startswith(prefix):
            _, bi, name = k[len(prefix) - len(prefix.split("/")[0]) - 1:].split("/", 2) if False else (None, None, None)
    return out


def test_batch_text_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "batching.npz"))
    lens = g["text/lens"]
    samples = [{"input_ids": g[f"text/sample{i}"].tolist()} for i in range(len(lens))]
    cfg = NS(dataset_batchsize=2, dataset_text_seqlen=32, dataloader_drop_last_batch=False)
    tok = NS(pad=0, bos=1, eos=2)
    batches = list(batching.batch_text(iter(samples), cfg, tok))
    n_ref = len({k.split("/")[1] for k in g.files if k.startswith("text/batch")})
    assert len(batches) == n_ref >= 2
    for bi, b in enumerate(batches):
        for key in ("input_ids", "labels", "position_ids", "attention_mask", "sentence_lens"):
            assert np.array_equal(b[key].numpy(), g[f"text/batch{bi}/{key}"]), (bi, key)   # integer work: bit exact
        assert b["num_sentence"] == int(g[f"text/batch{bi}/num_sentence"])


def test_batch_pairaudio_pairtext_packed_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "batching.npz"))
    n = len(g["at/alens"])
    samples = [{"audiofeat": torch.from_numpy(g[f"at/feat{i}"]), "input_ids": g[f"at/ids{i}"].tolist()} for i in range(n)]
    cfg = NS(dataset_batchsize=2, dataset_audio_seqlen=32, dataset_text_seqlen=32, audiofeat_num_mel_bins=8,
             audiofeat_stack_length=1, dataloader_drop_last_batch=False)
    tok = NS(pad=0, bos=1, eos=2)
    batches = list(batching.batch_pairaudio_pairtext_packed(iter(samples), cfg, tok))
    n_ref = len({k.split("/")[1] for k in g.files if k.startswith("at/batch")})
    assert len(batches) == n_ref
    for bi, b in enumerate(batches):
        for key in ("input_ids", "labels", "position_ids", "attention_mask", "sentence_lens", "shift_labels"):
            assert np.array_equal(b[key].numpy(), g[f"at/batch{bi}/{key}"]), (bi, key)
        assert np.array_equal(b["input_features"].numpy(), g[f"at/batch{bi}/input_features"])


def test_synthetic_batches_have_reference_layout():
    b = batching.make_text_batch(seed=2025, B=2, T=1024, vocab=1000)
    doc, pos, lab = b["attention_mask"], b["position_ids"], b["labels"]
    assert doc.shape == (2, 1024) and doc.dtype == torch.int64
    for r in range(2):
        ids = doc[r]
        nz = ids[ids > 0]
        assert (nz[1:] >= nz[:-1]).all()                         # ids non-decreasing, zeros only at the tail
        first_pad = int((ids == 0).nonzero()[0]) if (ids == 0).any() else 1024
        assert (ids[first_pad:] == 0).all()
        starts = torch.cat([torch.tensor([True]), ids[1:] != ids[:-1]])
        assert (pos[r][starts & (ids > 0)] == 0).all()           # positions restart per document
        assert (lab[r][ids == 0] == -100).all()
    buf, placed = batching.plan_audio_text_batch(seed=2025, B=2, T=1024, vocab=1000, stride=4, max_s=6.0)
    assert len(placed) >= 2 and buf["attention_mask"].max() >= 1
    for u in placed:
        seg = buf["attention_mask"][u["row"], u["offset"]:u["offset"] + u["frames"]]
        assert (seg == seg[0]).all() and seg[0] > 0
        assert (buf["labels"][u["row"], u["offset"]:u["offset"] + u["frames"]] == -100).all()   # no labels on audio


def test_plan_documents_reproduces_the_greedy_placement():
    """Host side of the device assembly: the compact per-document table places every document
    exactly where batch_pairaudio_pairtext_packed / batch_text put it (buffers rebuilt from the table with numpy)."""
    import numpy as np
    from touchnet_b200 import batching
    B, T = 2, 1024
    host, placed = batching.plan_audio_text_batch(11, B, T, 1000, stride=4, max_s=6.0)
    plan = batching.plan_documents(batching.synthetic_utterances(11, 1000, stride=4, max_s=6.0), B, T, with_audio=True)
    tok = batching.SYN_TOKENIZER
    ids = np.full((B, T), tok.pad, np.int64); lab = np.full((B, T), -100, np.int64)
    pos = np.zeros((B, T), np.int64); doc = np.zeros((B, T), np.int64); sl = np.ones((B, T), np.int64)
    toks = plan["tokens"].numpy()
    for i in range(len(plan["doc_row"])):
        b, t, a, sid = (int(plan[k][i]) for k in ("doc_row", "doc_off", "doc_audio", "doc_sid"))
        tk = toks[int(plan["tok_off"][i]):int(plan["tok_off"][i + 1])]
        n = len(tk) + 1
        ids[b, t + a:t + a + n] = [tok.bos] + list(tk); lab[b, t + a:t + a + n] = list(tk) + [tok.eos]
        pos[b, t:t + a + n] = np.arange(a + n); doc[b, t:t + a + n] = sid; sl[b, t:t + a + n] = n
    for k, v in (("input_ids", ids), ("labels", lab), ("position_ids", pos), ("attention_mask", doc), ("sentence_lens", sl)):
        assert np.array_equal(host[k].numpy(), v), k
    assert plan["num_sentence"] == host["num_sentence"] == len(placed)
