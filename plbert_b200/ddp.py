"""Data parallelism for the fused step: one process per GPU, bucketed SUM all-reduce of the flat fp32 gradient
arena over NCCL/NVLink, overlapped with the backward pass.

Reference semantics: ``accelerate`` wraps the model in ``DistributedDataParallel`` (R/train.py:160,
218-221), i.e. every rank computes its own loss normalisation (R/train.py:128-129) and gradients are AVERAGED
across ranks.  ALBERT shares one layer across 12 applications, so that layer's gradient is final only at the very
end of backward; the arena is therefore laid out heads-first and reduced in two buckets:

  bucket 0  = head parameters      -> all-reduce launched right after the head backward, hidden under the whole
                                      encoder backward (this is the large one when a 64 k token head is present);
  bucket 1  = shared layer + embeddings -> launched when backward ends; the 1/world scale is folded into AdamW.

The collective runs on a side stream; AdamW waits on it.  With world_size 1 nothing here runs.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def bucket_bounds(head_end: int, trainable_end: int, max_bucket_elems: Optional[int] = None) -> List[Tuple[int, int]]:
    """[begin, end) element ranges of the gradient buckets, in readiness order."""
    bounds: List[Tuple[int, int]] = []
    for lo, hi in ((0, head_end), (head_end, trainable_end)):
        if hi <= lo:
            continue
        if max_bucket_elems is None:
            bounds.append((lo, hi))
        else:
            for s in range(lo, hi, max_bucket_elems):
                bounds.append((s, min(s + max_bucket_elems, hi)))
    return bounds


def all_reduce_buckets(flat: torch.Tensor, bounds: Sequence[Tuple[int, int]], group=None, async_op: bool = False):
    """SUM all-reduce of each bucket of a flat tensor (works on NCCL and on gloo/CPU for tests)."""
    works = []
    for lo, hi in bounds:
        w = dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            works.append(w)
    return works


def broadcast_params(flat: torch.Tensor, src: int = 0, group=None) -> None:
    """DDP's initial parameter broadcast: every rank starts from rank ``src``'s weights."""
    dist.broadcast(flat, src=src, group=group)


class GradientSynchronizer:
    """Hooks an Engine's backward: bucket 0 after the heads, bucket 1 after the encoder."""

    def __init__(self, engine, group=None):
        self.engine, self.group = engine, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        # the collectives run on a side stream (CPU tensors — the gloo tests — are reduced in line)
        self.on_gpu = engine.device.type == "cuda"
        self.stream = torch.cuda.Stream(device=engine.device) if (self.world > 1 and self.on_gpu) else None
        lay = engine.layout
        self.head_bounds = [(0, lay.head_end)] if lay.head_end > 0 else []
        self.tail_bounds = [(lay.head_end, lay.trainable_end)]
        engine._after_head_backward = self._reduce_heads
        engine._after_encoder_backward = self._reduce_tail
        if self.world > 1:
            engine.always_reduce = True   # a rank whose shard has no masked rows still joins the collectives
            broadcast_params(engine.params, 0, group)
            engine.shadow_valid = False

    @property
    def grad_scale(self) -> float:
        return 1.0 / self.world

    def _launch(self, bounds) -> None:
        if self.world == 1 or not bounds:
            return
        if not self.on_gpu:
            all_reduce_buckets(self.engine.grads, bounds, self.group)
            return
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            all_reduce_buckets(self.engine.grads, bounds, self.group)

    def _reduce_heads(self) -> None:
        self._launch(self.head_bounds)

    def _reduce_tail(self) -> None:
        self._launch(self.tail_bounds)
        if self.world > 1 and self.on_gpu:
            torch.cuda.current_stream().wait_stream(self.stream)  # AdamW (next on the main stream) sees reduced grads
