"""Checkpoint interchange with the reference trainer.

The reference writes raw `state_dict()`s of its DDP-wrapped models (`worker.py:219-227`:
`gen_model.ckpt`, `gen_ema_model.ckpt`, `disc_model.ckpt`), so every key carries a `module.` prefix;
`generator_ema` is a `copy.deepcopy` of the wrapped generator (`worker.py:40`) and has it too.  The
modules of this package keep the reference's parameter / buffer names, shapes and fp32 dtypes, so a
checkpoint moves in either direction by adding or dropping that prefix - nothing is converted.
"""
from __future__ import annotations

from typing import Dict, Mapping, Union

import torch

PREFIX = "module."


def _as_state(src: Union[str, Mapping[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
    if isinstance(src, (str, bytes)) or hasattr(src, "__fspath__"):
        src = torch.load(src, map_location="cpu")
    return dict(src)


def strip_module_prefix(state: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Keys of a DDP-wrapped model -> keys of the bare module (unprefixed keys pass through)."""
    return {(k[len(PREFIX):] if k.startswith(PREFIX) else k): v for k, v in state.items()}


def add_module_prefix(state: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {(k if k.startswith(PREFIX) else PREFIX + k): v for k, v in state.items()}


def load_reference_checkpoint(module: torch.nn.Module, src: Union[str, Mapping[str, torch.Tensor]]) -> None:
    """Load a reference checkpoint (path or state dict, with or without the DDP prefix) into a bare
    `lcgan_b200.cnn.Generator` / `Discriminator` (or into a DDP-wrapped one).  Strict: a missing,
    unexpected or mis-shaped entry raises, naming it."""
    state = strip_module_prefix(_as_state(src))
    target = module.module if hasattr(module, "module") and isinstance(module.module, torch.nn.Module) else module
    own = target.state_dict()
    missing = sorted(set(own) - set(state))
    unexpected = sorted(set(state) - set(own))
    if missing or unexpected:
        raise KeyError(f"checkpoint does not match {type(target).__name__}: missing {missing[:5]} "
                       f"({len(missing)}), unexpected {unexpected[:5]} ({len(unexpected)})")
    for k, v in state.items():
        if tuple(v.shape) != tuple(own[k].shape):
            raise ValueError(f"{k}: checkpoint shape {tuple(v.shape)} != module shape {tuple(own[k].shape)}")
    target.load_state_dict(state, strict=True)


def reference_state_dict(module: torch.nn.Module) -> Dict[str, torch.Tensor]:
    """State dict in the layout `worker.py:load_model` expects (DDP prefix, fp32, CPU tensors)."""
    target = module.module if hasattr(module, "module") and isinstance(module.module, torch.nn.Module) else module
    return add_module_prefix({k: v.detach().to("cpu", torch.float32) if v.is_floating_point() else v.detach().cpu()
                              for k, v in target.state_dict().items()})


def save_reference_checkpoint(module: torch.nn.Module, path: str) -> None:
    torch.save(reference_state_dict(module), path)
