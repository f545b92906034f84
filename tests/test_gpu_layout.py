"""GPU parity (bit-exact, integer work): device-side assembly of the packed-batch integer buffers (tn_pack_layout_i64,
SURVEY 8(f) rank 3) vs the host batchers restated from the reference (touchnet_b200/batching.py, themselves pinned to
the reference's outputs by tests/golden/batching.npz)."""
import itertools
from types import SimpleNamespace

import pytest
import torch

from tests.gpu_util import require_cuda
from touchnet_b200 import batching

pytestmark = pytest.mark.gpu
KEYS = ("input_ids", "labels", "position_ids", "attention_mask", "sentence_lens")


@pytest.mark.parametrize("B,T", [(1, 8192), (3, 1024), (2, 300)])
def test_audio_text_layout_matches_host_batcher(B, T):
    dev = require_cuda()
    host, placed = batching.plan_audio_text_batch(2025, B, T, 128256, stride=4, max_s=10.0)
    plan = batching.plan_documents(batching.synthetic_utterances(2025, 128256, stride=4, max_s=10.0), B, T, with_audio=True)
    out = batching.assemble_on_device(plan, dev)
    assert out["num_sentence"] == host["num_sentence"] == len(placed)
    for k in KEYS:
        assert torch.equal(out[k].cpu(), host[k]), k
    assert [int(r) for r in plan["doc_row"]] == [u["row"] for u in placed]
    assert [int(o) for o in plan["doc_off"]] == [u["offset"] for u in placed]


@pytest.mark.parametrize("B,T", [(2, 1024), (4, 513)])
def test_text_layout_matches_host_batcher(B, T):
    dev = require_cuda()
    cfg = SimpleNamespace(dataset_batchsize=B, dataset_text_seqlen=T, dataloader_drop_last_batch=True)
    host = next(batching.batch_text(batching.synthetic_text_samples(7, 50257, T, mu_len=120.0), cfg, batching.SYN_TOKENIZER))
    plan = batching.plan_documents(batching.synthetic_text_samples(7, 50257, T, mu_len=120.0), B, T, with_audio=False)
    out = batching.assemble_on_device(plan, dev)
    for k in KEYS:
        assert torch.equal(out[k].cpu(), host[k]), k
    assert out["num_sentence"] == host["num_sentence"]


def test_edge_cases_empty_batch_and_empty_documents():
    dev = require_cuda()
    empty = batching.plan_documents(iter([]), 2, 256, with_audio=False)
    out = batching.assemble_on_device(empty, dev)
    tok = batching.SYN_TOKENIZER
    assert torch.all(out["input_ids"] == tok.pad) and torch.all(out["labels"] == -100)
    assert torch.all(out["attention_mask"] == 0) and torch.all(out["sentence_lens"] == 1) and torch.all(out["position_ids"] == 0)
    # documents without any text token: <bos> -> <eos> only (n_txt = 1), and one that fills the row to the last position
    docs = [{"input_ids": []}, {"input_ids": [5, 6, 7]}, {"input_ids": []}, {"input_ids": list(range(3, 3 + 249))}]
    plan = batching.plan_documents(iter(docs), 1, 256, with_audio=False)
    out = batching.assemble_on_device(plan, dev)
    cfg = SimpleNamespace(dataset_batchsize=1, dataset_text_seqlen=256, dataloader_drop_last_batch=False)
    host = next(batching.batch_text(iter(docs), cfg, tok))
    for k in KEYS:
        assert torch.equal(out[k].cpu(), host[k]), k
    assert int(out["attention_mask"][0, -1]) == 4 and int(out["labels"][0, -1]) == tok.eos
