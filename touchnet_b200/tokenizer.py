"""BEST-RQ tokenizer on the GPU, mirroring ref: touchnet/tokenizer/tokenizer.py:236-318 (BestRQTokenizer).

Same construction (seeded xavier-uniform projection [input_size, emb_size], seeded normal codebook [vocab, emb_size]
L2-normalised per row, built on the CPU with the same generator so the random tensors are identical), same
`tokenize(inputs) -> list[int]` contract; the projection / normalise / nearest-code search run in csrc/tokenizer.cu
instead of a [T, V, E] CPU broadcast inside the DataLoader worker."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib


class BestRQTokenizer:
    def __init__(self, config, device="cuda", **kwargs):
        self.config = config
        self.device = torch.device(device)
        self._quantizer = None
        self._codebook = None

    def _build_quantizer_and_codebook(self):
        if self._quantizer is not None:
            return
        c = self.config
        if c.tokenizer_bestrq_init_method != "default":
            raise NotImplementedError(f"Initialization method {c.tokenizer_bestrq_init_method} is not implemented.")
        q = torch.empty(c.tokenizer_bestrq_input_size, c.tokenizer_bestrq_emb_size)
        cb = torch.empty(c.tokenizer_bestrq_vocab_size, c.tokenizer_bestrq_emb_size)
        g = torch.Generator().manual_seed(c.tokenizer_bestrq_init_seed)      # same draws as the reference (:262-264)
        torch.nn.init.xavier_uniform_(q, generator=g)
        torch.nn.init.normal_(cb, generator=g)
        cb = F.normalize(cb, dim=1, p=2, eps=1e-8)
        self._quantizer_cpu, self._codebook_cpu = q, cb
        self._quantizer = q.to(self.device).contiguous()
        self._codebook = cb.to(self.device).contiguous()

    @property
    def vocab_size(self):
        self._build_quantizer_and_codebook()
        return self._codebook.size(0)

    @property
    def inv_vocab(self):
        self._build_quantizer_and_codebook()
        return self._codebook

    decoder = inv_vocab

    def tokenize_tensor(self, inputs: torch.Tensor) -> torch.Tensor:
        """[T, D] fp32 features (CUDA) -> int32 codes [T] (stays on the device)."""
        self._build_quantizer_and_codebook()
        if not inputs.is_cuda:
            inputs = inputs.to(self.device)
        x = inputs.float()
        if x.stride(-1) != 1:
            x = x.contiguous()
        T, D = x.shape
        assert D == self._quantizer.shape[0], (D, self._quantizer.shape)
        codes = torch.empty(T, dtype=torch.int32, device=x.device)
        _lib.call("tn_bestrq_tokenize_f32", x.data_ptr(), x.stride(0), self._quantizer.data_ptr(),
                  self._codebook.data_ptr(), T, D, self._quantizer.shape[1], self._codebook.shape[0], codes.data_ptr(),
                  torch.cuda.current_stream().cuda_stream)
        return codes

    def tokenize(self, inputs, **kwargs):
        return self.tokenize_tensor(inputs).tolist()

    def detokenize(self, token_ids, **kwargs):
        self._build_quantizer_and_codebook()
        return torch.index_select(self._codebook, dim=0, index=token_ids.to(self._codebook.device))
