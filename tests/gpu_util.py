"""Helpers shared by the -m gpu parity tests."""
import torch


def require_cuda():
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device; they do not skip and there is no CPU path"
    return torch.device("cuda:0")


def packed_doc_ids(B, T, lens_per_row, device="cpu"):
    """Document-id / position tensors in the layout of ref touchnet/models/llama/processing_llama.py:24-104."""
    doc = torch.zeros(B, T, dtype=torch.int64)
    pos = torch.zeros(B, T, dtype=torch.int64)
    for b, lens in enumerate(lens_per_row):
        o = 0
        for i, n in enumerate(lens):
            doc[b, o:o + n] = i + 1
            pos[b, o:o + n] = torch.arange(n)
            o += n
        assert o <= T
    return doc.to(device), pos.to(device)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_err(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).abs().max())
