"""Mirrors of the step-level functions of R/train.py, routed to the B200 path.

Same names and argument meaning as the reference so a maintainer can swap the import:
``length_to_mask`` (R/train.py:34-44), ``process_batch`` (:381-390), ``find_latest_checkpoint`` (:46-79),
``load_checkpoint`` (:81-105), ``save_checkpoint`` (:412-425).  The training driver itself (wandb, accelerate,
dataset download) is out of scope.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from . import _abi
from ._abi import call, ptr
from .model import normalize_state_dict_keys


def length_to_mask(lengths: torch.Tensor) -> torch.Tensor:
    """R/train.py:34-44 — bool [B, max_len], True where ``position + 1 > length`` (padding).

    ``lengths`` on a CUDA device: computed by ``plb_length_to_mask`` (bit-exact integer kernel).  The reference
    calls this on a CPU tensor (R/train.py:384); a CPU input is a host-side integer comparison, not a compute
    fallback, and is kept so existing callers keep working."""
    if lengths.dim() != 1 or lengths.numel() == 0:
        raise ValueError("lengths must be a non-empty 1-D tensor")
    B = lengths.size(0)
    max_len = int(lengths.max().item())
    if lengths.device.type == "cuda":
        li = lengths.to(torch.int32).contiguous()
        valid = torch.empty((B, max_len), dtype=torch.int32, device=lengths.device)
        call("plb_length_to_mask", ptr(li), B, max_len, ptr(valid), torch.cuda.current_stream().cuda_stream)
        return valid == 0
    positions = torch.arange(max_len).expand(B, max_len)
    return (positions + 1) > lengths.unsqueeze(1)


def process_batch(model, batch, criterion=None, accelerator=None) -> torch.Tensor:
    """R/train.py:381-390 — forward + loss for one collate batch.  On the B200 path the backward pass is fused
    with it (``model.loss_and_backward``): the returned loss is a detached device scalar and the gradients are
    already in ``p.grad``, so the caller's ``accelerator.backward(loss)`` / ``loss.backward()`` must be dropped
    (see INTEGRATION.md).  ``criterion`` / ``accelerator`` are accepted for signature compatibility."""
    return model.loss_and_backward(batch)


def find_latest_checkpoint(log_dir: str) -> Tuple[bool, int]:
    """R/train.py:46-79 — newest ``step_<int>.<ext>`` in ``log_dir`` (extension-agnostic)."""
    found, last = False, 0
    try:
        steps = []
        for f in os.listdir(log_dir):
            if f.startswith("step_") and os.path.isfile(os.path.join(log_dir, f)):
                try:
                    steps.append(int(f.split("_")[-1].split(".")[0]))
                except ValueError:
                    continue
        if steps:
            found, last = True, max(steps)
    except Exception as e:  # same forgiving behaviour as the reference
        print(f"Error finding checkpoints: {e}")
        found, last = False, 0
    return found, last


def save_checkpoint(model, optimizer, current_step: int, log_dir: str, current_epoch: int = 0, *,
                    extension: str = "t7", module_prefix: bool = True) -> str:
    """R/train.py:412-425 — ``{'net', 'step', 'epoch', 'optimizer'}`` pickled with torch.save.

    ``extension`` defaults to ``t7`` (north star); the reference writes ``.pth`` (R/train.py:423) and reads any.
    ``module_prefix`` emits the DDP ``module.`` key prefix the StyleTTS loader slices off unconditionally
    (R/README.md:62) and that R/train.py:98 / R/convert_to_hf.py:40 strip when present."""
    net = {("module." + k if module_prefix else k): v.detach().to("cpu").clone() for k, v in model.state_dict().items()}
    opt = optimizer.state_dict()
    state = {"net": net, "step": current_step, "epoch": current_epoch, "optimizer": opt}
    os.makedirs(log_dir, exist_ok=True)
    path = os.path.join(log_dir, f"step_{current_step}.{extension}")
    torch.save(state, path)
    return path


def load_checkpoint(model, optimizer, checkpoint_path: str, accelerator=None, device: Optional[torch.device] = None):
    """R/train.py:81-105 — strict=False weight load (``module.`` stripped, upstream head names aliased) +
    optimizer state."""
    map_location = device or (accelerator.device if accelerator is not None else "cpu")
    checkpoint = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
    model.load_state_dict(normalize_state_dict_keys(checkpoint["net"]), strict=False)
    if optimizer is not None and checkpoint.get("optimizer") is not None:
        optimizer.load_state_dict(checkpoint["optimizer"])
    return model, optimizer
