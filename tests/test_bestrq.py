"""BEST-RQ tokenizer: integer labels -> index parity.
CPU: the oracle and the host-side construction against the reference tokenizer's own outputs (tests/golden/bestrq.npz);
GPU: the CUDA kernel through the C ABI.  Indices must be identical except where the reference's own best-vs-second
distance margin is at fp32 rounding level (< 1e-6), which the test counts and bounds."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from oracle import frontend_oracle as fo


def _cfg(g, tag):
    D, E, V, seed = (int(x) for x in g[f"{tag}/cfg"])
    return NS(tokenizer_bestrq_vocab_size=V, tokenizer_bestrq_input_size=D, tokenizer_bestrq_emb_size=E,
              tokenizer_bestrq_init_seed=seed, tokenizer_bestrq_init_method="default")


def _tables(cfg):
    """The reference's construction (touchnet/tokenizer/tokenizer.py:246-267), restated for the oracle."""
    q = torch.empty(cfg.tokenizer_bestrq_input_size, cfg.tokenizer_bestrq_emb_size)
    cb = torch.empty(cfg.tokenizer_bestrq_vocab_size, cfg.tokenizer_bestrq_emb_size)
    gen = torch.Generator().manual_seed(cfg.tokenizer_bestrq_init_seed)
    torch.nn.init.xavier_uniform_(q, generator=gen)
    torch.nn.init.normal_(cb, generator=gen)
    return q.numpy(), torch.nn.functional.normalize(cb, dim=1, p=2, eps=1e-8).numpy()


@pytest.mark.parametrize("tag", ["default", "small"])
def test_oracle_and_construction_match_reference(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "bestrq.npz"))
    cfg = _cfg(g, tag)
    q, cb = _tables(cfg)
    assert np.array_equal(q[:4], g[f"{tag}/quantizer_head"]) and np.array_equal(cb[:4], g[f"{tag}/codebook_head"])
    assert np.allclose([q.astype(np.float64).sum(), cb.astype(np.float64).sum()], g[f"{tag}/sums"], rtol=0, atol=1e-6)
    codes, margin = fo.bestrq_tokenize(g[f"{tag}/feats"], q, cb)
    ref = g[f"{tag}/codes"]
    diff = codes != ref
    assert not (diff & (margin > 1e-6)).any()          # decisive positions: bit-exact token indices
    assert diff.sum() <= 1


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["default", "small"])
def test_cuda_tokenizer_matches_reference_codes(golden_dir, tag):
    assert torch.cuda.is_available()
    from touchnet_b200.tokenizer import BestRQTokenizer
    g = np.load(os.path.join(golden_dir, "bestrq.npz"))
    cfg = _cfg(g, tag)
    tok = BestRQTokenizer(cfg)
    feats = torch.from_numpy(g[f"{tag}/feats"]).cuda()
    codes = np.array(tok.tokenize(feats))
    q, cb = _tables(cfg)
    assert np.array_equal(tok._quantizer_cpu.numpy(), q) and np.array_equal(tok._codebook_cpu.numpy(), cb)
    _, margin = fo.bestrq_tokenize(g[f"{tag}/feats"], q, cb)
    ref = g[f"{tag}/codes"]
    diff = codes != ref
    assert not (diff & (margin > 1e-6)).any()
    assert diff.sum() <= 1
    assert tok.vocab_size == cfg.tokenizer_bestrq_vocab_size
    # a strided [T, D] view (features sliced out of a wider buffer) gives the same codes
    wide = torch.zeros(feats.shape[0], feats.shape[1] + 8, device="cuda")
    wide[:, :feats.shape[1]] = feats
    assert np.array_equal(np.array(tok.tokenize(wide[:, :feats.shape[1]])), codes)
