"""`TrainSpec.parallelize_fn` for the B200 modules: `(model, world_mesh, parallel_dims, job_config) -> model`
(ref contract: touchnet/utils/train_spec.py:25-44; ref implementation: touchnet/models/llama/parallelize_llama.py:29-102).

Order as in the reference: tensor parallel -> (activation checkpointing) -> FSDP2 over `dp_shard_cp`.

* tp  -> touchnet_b200.tensor_parallel.apply_tp: the reference's `parallelize_module` plan hooks nn.Linear.forward, which
         the fused decoder block never calls, so the same sharding is applied to the parameters and the collectives run
         inside the block.
* cp  -> touchnet_b200.context_parallel.enable_context_parallel on the `cp` sub-mesh, in the shard layout torch's
         `context_parallel` context will produce (head-tail load balancing on by default, contiguous when switched off).
* dp  -> FSDP2 per decoder block + root, same mixed-precision / reshard policies as ref helper_func.py:134-202.  When the
         reference package is importable its own parallelize function does this part (it also owns AC / compile / DDP);
         it is handed a view of `parallel_dims` with tp switched off so that it does not re-apply its DTensor plan.
"""
from __future__ import annotations

from typing import Any, Callable, Optional

import torch

from . import context_parallel, tensor_parallel

_DTYPES = {"float32": torch.float32, "bfloat16": torch.bfloat16, "float16": torch.float16}


class _WithoutTP:
    """Attribute view of the reference's ParallelDims with tensor parallelism reported as already handled."""

    def __init__(self, dims):
        self._dims = dims

    def __getattr__(self, name):
        if name == "tp_enabled":
            return False
        return getattr(self._dims, name)


def _torch_cp_load_balance() -> bool:
    """Whether the train loop's `context_parallel` context (ref: touchnet/utils/distributed.py:292-315) will hand the model
    head-tail load-balanced sequence shards (torch's default) or contiguous ones."""
    try:
        from torch.distributed.tensor.experimental._attention import _cp_options
        return bool(_cp_options.enable_load_balance)
    except Exception:
        return False


def apply_fsdp(model: torch.nn.Module, dp_mesh, param_dtype=torch.bfloat16, reduce_dtype=torch.float32,
               reshard_after_forward_policy: str = "default", pp_enabled: bool = False) -> torch.nn.Module:
    """FSDP2 on every decoder block and on the root (ref: touchnet/models/helper_func.py:134-202, same policies)."""
    from torch.distributed.fsdp import MixedPrecisionPolicy, fully_shard
    mp = MixedPrecisionPolicy(param_dtype=param_dtype, reduce_dtype=reduce_dtype)
    base = getattr(model, getattr(model, "base_model_prefix", "model"))
    base = getattr(base, "model", base) if not hasattr(base, "layers") else base
    layers = list(base.layers)
    for i, block in enumerate(layers):
        if reshard_after_forward_policy == "always":
            reshard = True
        elif reshard_after_forward_policy == "never":
            reshard = False
        elif reshard_after_forward_policy == "default":
            reshard = (not pp_enabled) and i < len(layers) - 1      # last block is needed again at once in backward
        else:
            raise ValueError(f"Invalid reshard_after_forward_policy: {reshard_after_forward_policy}.")
        fully_shard(block, mesh=dp_mesh, mp_policy=mp, reshard_after_forward=reshard)
    fully_shard(model, mesh=dp_mesh, mp_policy=mp, reshard_after_forward=not pp_enabled)
    return model


def make_parallelize_fn(base_fn: Optional[Callable] = None) -> Callable:
    """Build the parallelize_fn of a "*_b200" TrainSpec; `base_fn` is the reference spec's own parallelize_fn (or None
    when the reference is not importable)."""

    def parallelize_b200(model: torch.nn.Module, world_mesh, parallel_dims: Any, job_config: Any) -> torch.nn.Module:
        if getattr(parallel_dims, "pp_enabled", False):
            raise NotImplementedError("pipeline parallelism is outside the B200 hot path (SURVEY 8, out of scope)")
        if getattr(parallel_dims, "tp_enabled", False):
            tensor_parallel.apply_tp(model, world_mesh["tp"])
            if getattr(parallel_dims, "loss_parallel_enabled", False):       # ref: parallelize_llama.py:127-131
                lm = model.language_model if hasattr(model, "language_model") else model
                lm.loss_parallel = True
        if getattr(parallel_dims, "cp_enabled", False):
            context_parallel.enable_context_parallel(model, world_mesh["cp"].get_group(), load_balance=_torch_cp_load_balance())
        if base_fn is not None:
            return base_fn(model, world_mesh, _WithoutTP(parallel_dims), job_config)
        if getattr(parallel_dims, "dp_shard_enabled", False) or getattr(parallel_dims, "cp_enabled", False):
            names = ("dp_replicate", "dp_shard_cp") if getattr(parallel_dims, "dp_replicate_enabled", False) else ("dp_shard_cp",)
            apply_fsdp(model, world_mesh[names],
                       param_dtype=_DTYPES[getattr(job_config, "training_mixed_precision_param", "bfloat16")],
                       reduce_dtype=_DTYPES[getattr(job_config, "training_mixed_precision_reduce", "float32")],
                       reshard_after_forward_policy=getattr(job_config, "training_fsdp_reshard_after_forward", "default"))
        return model

    return parallelize_b200
