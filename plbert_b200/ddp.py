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


def head_bounds_without(layout, fused_names: Sequence[str]) -> List[Tuple[int, int]]:
    """[begin, end) ranges of the head region [0, head_end) that remain for the NCCL bucket once the gradients of
    ``fused_names`` are reduced inside their WGRAD kernels."""
    bounds, lo = [], 0
    for n in sorted(fused_names, key=lambda k: layout.offsets[k]):
        o = layout.offsets[n]
        if o > lo:
            bounds.append((lo, o))
        lo = o + layout.numel(n)
    if layout.head_end > lo:
        bounds.append((lo, layout.head_end))
    return bounds


def broadcast_params(flat: torch.Tensor, src: int = 0, group=None) -> None:
    """DDP's initial parameter broadcast: every rank starts from rank ``src``'s weights."""
    dist.broadcast(flat, src=src, group=group)


FUSED_MIN_ELEMS = 1 << 20   # head weights at least this large get the fused WGRAD + all-reduce (the 64 k token head)


class GradientSynchronizer:
    """Hooks an Engine's backward: bucket 0 after the heads, bucket 1 after the encoder.

    ``fuse_head_allreduce`` (opt-in; needs NVSwitch multicast on one node): the gradient arena moves into symmetric
    memory and the WGRAD of every large head weight adds straight into ALL ranks' arenas through the switch
    (``plb_gemm_wgrad_allreduce``: multimem.red in the GEMM epilogue), so the 197 MB of the 64 k token head never pass
    through an NCCL kernel.  Ordering: a symmetric-memory barrier on the main stream before the head backward (no rank adds
    into a peer's arena before that peer's AdamW has consumed and zeroed the previous step's sum), and the NCCL all-reduce of
    the last bucket before AdamW (a rank's all-reduce completes only after every rank has enqueued its own, i.e. after every
    rank's head WGRAD kernel has completed).  Falls back to the NCCL bucket when symmetric memory or multicast is
    unavailable on ANY rank.  MEASURED on 2 B200s (profiles/ab_ddp_fused.py, same process, alternating blocks): numerically
    equivalent (profiles/r2g_ddp_equivalence_fused.json) but 57.6 vs 54.1 ms per step — 16-byte multimem.red operations
    move the 197 MB at ~50 GB/s, while the NCCL bucket on the side stream is hidden under the encoder backward — hence
    off by default.  With more than two ranks the arrival order of the adds differs per copy, so replicas would also stop
    being bit-identical."""

    def __init__(self, engine, group=None, fuse_head_allreduce: bool = False):
        self.engine, self.group = engine, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        # the collectives run on a side stream (CPU tensors — the gloo tests — are reduced in line)
        self.on_gpu = engine.device.type == "cuda"
        self.stream = torch.cuda.Stream(device=engine.device) if (self.world > 1 and self.on_gpu) else None
        lay = engine.layout
        self.head_bounds = [(0, lay.head_end)] if lay.head_end > 0 else []
        self.tail_bounds = [(lay.head_end, lay.trainable_end)]
        self.symm = None              # symmetric-memory handle of the gradient arena (fused mode)
        self.fused_heads: List[str] = []
        self.fuse_error: Optional[str] = None
        engine._after_head_backward = self._reduce_heads
        engine._after_encoder_backward = self._reduce_tail
        if self.world > 1:
            engine.always_reduce = True   # a rank whose shard has no masked rows still joins the collectives
            broadcast_params(engine.params, 0, group)
            engine.shadow_valid = False
            if fuse_head_allreduce and self.on_gpu:
                self._setup_fused_heads()

    def _setup_fused_heads(self) -> None:
        """Gradient arena -> symmetric memory + multicast mapping; every rank must succeed or nobody uses it."""
        eng, lay = self.engine, self.engine.layout
        names = [n for n in lay.offsets if n.endswith("_predictor.weight") and lay.offsets[n] < lay.head_end
                 and lay.numel(n) >= FUSED_MIN_ELEMS and lay.shapes[n][1] % 4 == 0 and lay.offsets[n] % 4 == 0]
        ok, handle, arena = bool(names), None, None
        if ok:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                grp = self.group if self.group is not None else dist.group.WORLD
                arena = symm_mem.empty(eng.grads.numel(), dtype=torch.float32, device=eng.device)
                handle = symm_mem.rendezvous(arena, grp)
                ok = int(getattr(handle, "multicast_ptr", 0) or 0) != 0
                if not ok:
                    self.fuse_error = "no multicast mapping (multicast_ptr == 0)"
            except Exception as e:  # noqa: BLE001 - any failure means: keep the NCCL bucket
                ok, self.fuse_error = False, f"{type(e).__name__}: {e}"
        flag = torch.tensor([1 if ok else 0], device=eng.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) != 1:
            if self.fuse_error is None:
                self.fuse_error = "a peer rank could not set up symmetric memory" if names else "no head weight large enough"
            return
        eng.rebind_grads(arena)
        self.symm = handle
        base = int(handle.multicast_ptr)
        self.fused_heads = names
        eng.grad_multicast = {n: base + 4 * lay.offsets[n] for n in names}
        eng._before_head_backward = self._head_barrier
        self.head_bounds = head_bounds_without(lay, names)   # the NCCL head bucket keeps what is left of [0, head_end)

    def _head_barrier(self) -> None:
        self.symm.barrier(channel=0)

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
