"""Host side of the hot path's input contract.

The reference's collate (R/dataloader.py:200-223 ``Collater``, :276-297 ``PhonemeOnlyCollater``) yields
``(phoneme_labels[B,L] int64, masked_phonemes[B,L] int64, input_lengths: list[int], masked_indices: list[list[int]])``
(the 5-tuple form prepends ``token_ids[B,L]``).  This module turns that tuple into the flat integer arrays the
kernels consume — bit-exactly, on the host, once per batch — and stages them through pinned memory.  It also
provides a synthetic generator with the reference pipeline's statistics for benchmarks (the reference dataset
code cannot travel to the GPU box).
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

PAD_ID = 0          # R/char_indexer.py:5,15 — symbols[0] == "P"
MASK_ID = 185       # 'M'
SEPARATOR_ID = 186  # ' '
NUM_SYMBOLS = 188   # len(symbols), probe in SURVEY.md §0


def flatten_masked_indices(input_lengths: Sequence[int], masked_indices: Sequence[Sequence[int]],
                           seq_stride) -> Tuple[np.ndarray, np.ndarray]:
    """Row selection of R/train.py:123-129 as flat arrays: ``rows[k] = b * seq_stride + i`` in the reference's
    iteration order and ``weights[k] = 1 / (count * n_b)`` where ``count`` is the number of samples with at least
    one masked index.  int32 / float32, bit-exact index arithmetic.  ``seq_stride`` is the padded row length, or an
    integer array with the first row of every sample (packed layout: ``rows[k] = seq_stride[b] + i``)."""
    # one pass over the python lists, everything else vectorised (the per-sample loop cost 2-5 ms per 192-sample batch)
    B = len(masked_indices)
    counts = np.fromiter((len(m) for m in masked_indices), dtype=np.int64, count=B)
    total = int(counts.sum())
    if total == 0:
        return np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.float32)
    flat = np.fromiter(itertools.chain.from_iterable(masked_indices), dtype=np.int64, count=total)
    sample = np.repeat(np.arange(B, dtype=np.int64), counts)            # owning sample of every masked index
    length = np.asarray(input_lengths, dtype=np.int64)[sample]
    flat = np.where(flat < 0, flat + length, flat)                      # python-style negative indices
    bad = (flat < 0) | (flat >= length)
    if bad.any():
        b = int(sample[np.argmax(bad)])
        raise IndexError(f"masked index out of range for sample {b} (length {int(input_lengths[b])})")
    count = int(np.count_nonzero(counts))                               # samples with at least one masked index
    weights = 1.0 / (count * counts[sample].astype(np.float64))
    base = sample * int(seq_stride) if np.isscalar(seq_stride) else np.asarray(seq_stride, dtype=np.int64)[sample]
    return (flat + base).astype(np.int32), weights.astype(np.float32)


def flatten_all_valid(input_lengths: Sequence[int], seq_stride, label_stride: int):
    """Row set "every valid position" (upstream's word loss: per-sample mean over ALL positions of the sample, then the
    mean over samples): ``rows[k] = base_b + i`` for i < len_b in the reference's iteration order, the matching index
    into the padded [B, label_stride] label tensors, and ``weights[k] = 1 / (B * len_b)``.  ``seq_stride`` as in
    ``flatten_masked_indices`` (padded row length, or the array of first rows in the packed layout)."""
    lengths = np.asarray(input_lengths, dtype=np.int64)
    B = lengths.shape[0]
    sample = np.repeat(np.arange(B, dtype=np.int64), lengths)
    within = np.arange(int(lengths.sum()), dtype=np.int64) - np.repeat(np.cumsum(lengths) - lengths, lengths)
    base = sample * int(seq_stride) if np.isscalar(seq_stride) else np.asarray(seq_stride, dtype=np.int64)[sample]
    weights = (1.0 / (B * lengths[sample].astype(np.float64))).astype(np.float32)
    return (base + within).astype(np.int32), sample * int(label_stride) + within, weights


@dataclass
class DeviceBatch:
    """One batch resident in HBM: padded [B, L] row layout (row = b * L + j) or, with ``packed``, the sequences back to
    back without padding rows (row = seq_start[b] + j, q_len == kv_len == the sequence length)."""

    B: int
    L: int
    ids: torch.Tensor          # int32 [B*L] input symbols (masked_phonemes)
    pos: torch.Tensor          # int32 [B*L] position ids (restart per sequence)
    seq_start: torch.Tensor    # int32 [B]
    q_len: torch.Tensor        # int32 [B]  (= L)
    kv_len: torch.Tensor       # int32 [B]  (= input_lengths)
    rows: Optional[torch.Tensor] = None       # int32 [M] masked rows
    weights: Optional[torch.Tensor] = None    # fp32 [M]
    labels: List[torch.Tensor] = field(default_factory=list)  # per head: int32 [M] labels at the masked rows
    M: int = 0
    num_valid_tokens: int = 0
    h2d_bytes: int = 0
    packed: bool = False
    n_rows: int = -1           # activation rows: B * L (padded) or sum of the lengths (packed)
    heads: List[str] = field(default_factory=list)   # head of every entry of ``labels`` (set by model.stage_batch)
    # second row set: EVERY valid position, weight 1 / (B * len_b) — the upstream word-loss convention;
    # ``all_valid[i]`` says that ``labels[i]`` (and head i) uses it instead of the masked rows
    rows_all: Optional[torch.Tensor] = None   # int32 [M_all]
    weights_all: Optional[torch.Tensor] = None
    M_all: int = 0
    all_valid: List[bool] = field(default_factory=list)

    @property
    def T(self) -> int:
        return self.n_rows if self.n_rows >= 0 else self.B * self.L


class BatchStager:
    """Packs a collate tuple into ONE pinned host buffer and uploads it with one async copy.

    Replaces the blocking pageable copies and the CPU mask construction at R/train.py:383-386."""

    SLOTS = 2  # a staged batch stays valid until SLOTS more batches have been staged

    def __init__(self, device: torch.device):
        self.device = device
        self._pinned: List[Optional[torch.Tensor]] = [None] * self.SLOTS
        self._dev: List[Optional[torch.Tensor]] = [None] * self.SLOTS
        self._done: List[Optional[torch.cuda.Event]] = [None] * self.SLOTS
        self._next = 0

    def _buffers(self, n_int32: int):
        s = self._next
        self._next = (s + 1) % self.SLOTS
        if self._done[s] is not None:
            self._done[s].synchronize()  # the previous upload out of this pinned slot has finished
        if self._pinned[s] is None or self._pinned[s].numel() < n_int32:
            cap = max(n_int32, 1024) * 2
            self._pinned[s] = torch.empty(cap, dtype=torch.int32, pin_memory=True)
            self._dev[s] = torch.empty(cap, dtype=torch.int32, device=self.device)
        return s, self._pinned[s], self._dev[s]

    def stage(self, masked_phonemes: torch.Tensor, input_lengths: Sequence[int],
              masked_indices: Optional[Sequence[Sequence[int]]] = None,
              label_tensors: Sequence[torch.Tensor] = (), packed: bool = False,
              all_valid_heads: Sequence[int] = ()) -> DeviceBatch:
        """``packed``: drop the padding rows: the sequences are laid out back to back, so every
        per-token kernel (GEMMs, LayerNorm, attention query tiles) only sees valid tokens.  The reference computes the
        padded rows and never reads them (R/train.py:125), so the loss and every gradient are unchanged."""
        ids = masked_phonemes
        if ids.dim() != 2:
            raise ValueError("masked_phonemes must be [B, L]")
        B, L = ids.shape
        lengths = np.asarray([int(x) for x in input_lengths], dtype=np.int32)
        if lengths.shape[0] != B or lengths.min() < 1 or lengths.max() > L:
            raise ValueError("input_lengths must hold one length in [1, L] per sample")
        if packed:
            starts = np.zeros(B, dtype=np.int64)
            np.cumsum(lengths[:-1], out=starts[1:])
            T = int(lengths.sum())
            sample = np.repeat(np.arange(B, dtype=np.int64), lengths)
            pos_np = (np.arange(T, dtype=np.int64) - starts[sample]).astype(np.int32)
            src = sample * L + pos_np                       # padded flat index of every packed row
        else:
            starts = np.arange(B, dtype=np.int64) * L
            T = B * L
            pos_np = np.tile(np.arange(L, dtype=np.int32), B)
            src = None
        on_host = ids.device.type == "cpu"
        ids_np = None
        if on_host:
            ids_np = ids.reshape(-1).to(torch.int32).numpy()
            if src is not None:
                ids_np = ids_np[src]
        if masked_indices is not None:
            rows, weights = flatten_masked_indices(lengths, masked_indices, starts if packed else L)
            label_rows = flatten_masked_indices(lengths, masked_indices, L)[0] if packed else rows  # labels are padded [B, L]
        else:
            rows, weights = np.zeros(0, np.int32), np.zeros(0, np.float32)
            label_rows = rows
        M = rows.shape[0]
        all_valid = [i in set(all_valid_heads) for i in range(len(label_tensors))]
        M_all = 0
        if any(all_valid):
            # every valid position of every sample, in the reference's iteration order (sample by sample, position by position);
            # per-sample mean, then mean over samples: weight 1 / (B * len_b)
            rows_all, label_rows_all, weights_all = flatten_all_valid(lengths, starts if packed else L, L)
            M_all = rows_all.shape[0]
        labs = []
        for i, lt in enumerate(label_tensors):
            flat = lt.reshape(-1)
            sel, m_i = (label_rows_all, M_all) if all_valid[i] else (label_rows.astype(np.int64), M)
            labs.append(flat[torch.from_numpy(sel)].to(torch.int32).cpu().numpy() if m_i else np.zeros(0, np.int32))
        # layout (int32 words): ids[T] | pos[T] | seq_start[B] | q_len[B] | kv_len[B] | rows[M] | weights[M] | labels.. |
        # rows_all[M_all] | weights_all[M_all]
        n = 2 * T + 3 * B + 2 * M + sum(la.shape[0] for la in labs) + 2 * M_all
        slot, pinned, dev = self._buffers(n)
        host = pinned.numpy()
        o = 0
        if ids_np is not None:
            host[o:o + T] = ids_np
        o += T
        host[o:o + T] = pos_np; o += T
        host[o:o + B] = starts.astype(np.int32); o += B
        host[o:o + B] = lengths if packed else L; o += B
        host[o:o + B] = lengths; o += B
        host[o:o + M] = rows; o += M
        host[o:o + M] = weights.view(np.int32); o += M
        for la in labs:
            host[o:o + la.shape[0]] = la; o += la.shape[0]
        if M_all:
            host[o:o + M_all] = rows_all; o += M_all
            host[o:o + M_all] = weights_all.view(np.int32); o += M_all
        dev[:n].copy_(pinned[:n], non_blocking=True)
        if self._done[slot] is None:
            self._done[slot] = torch.cuda.Event()
        self._done[slot].record()
        if not on_host:  # ids already on the device
            flat_ids = ids.reshape(-1).to(torch.int32)
            dev[:T].copy_(flat_ids if src is None else flat_ids[torch.from_numpy(src).to(ids.device)])
        o = 0
        def take(k):
            nonlocal o
            v = dev[o:o + k]; o += k
            return v
        batch = DeviceBatch(B=B, L=int(lengths.max()) if packed else L, ids=take(T), pos=take(T), seq_start=take(B), q_len=take(B),
                            kv_len=take(B))
        batch.rows = take(M)
        batch.weights = take(M).view(torch.float32)
        batch.labels = [take(la.shape[0]) for la in labs]
        if M_all:
            batch.rows_all = take(M_all)
            batch.weights_all = take(M_all).view(torch.float32)
        batch.M_all = M_all
        batch.all_valid = all_valid
        batch.M = M
        batch.num_valid_tokens = int(lengths.sum())
        batch.h2d_bytes = 4 * n
        batch.packed = packed
        batch.n_rows = T
        return batch


class BatchFeeder:
    """Input pipeline off the critical path: keeps one batch staged ahead of the step that runs.

    The reference pulls a batch, builds masks on the CPU and copies with blocking pageable transfers before every step
    (R/train.py:350-353, 383-386; ``num_workers=0`` at :253), so the GPU idles while the host prepares inputs.  Here
    ``step()`` enqueues the kernels of step i on the batch that was staged during step i-1, then flattens / pins /
    uploads batch i+1 while the GPU is busy, and only then hands the loss back — by the time the caller reads it
    (``loss.item()``) the next batch is already resident::

        feed = BatchFeeder(model, loader)
        while not feed.exhausted:
            loss = feed.step(); optimizer.step(); print(loss.item())

    Every batch still takes its own pinned-memory H2D copy; nothing is cached."""

    def __init__(self, model, batches):
        self.model = model
        self._it = iter(batches)
        self._staged: Optional[DeviceBatch] = None
        self._stager = model.new_stager()   # own slots: validation / inference calls in between must not recycle them
        self._advance()

    def _advance(self) -> None:
        try:
            host = next(self._it)
        except StopIteration:
            self._staged = None
            return
        self._staged = self.model.stage_batch(host, stager=self._stager)

    @property
    def exhausted(self) -> bool:
        return self._staged is None

    def step(self) -> torch.Tensor:
        """forward + loss + backward on the staged batch (R/train.py:353-356), then stage the next one."""
        if self._staged is None:
            raise StopIteration("BatchFeeder: the batch iterable is exhausted")
        cur = self._staged
        loss = self.model.loss_and_backward(cur)
        self._advance()
        return loss


def synthetic_collate_batch(batch_size: int, max_seq_length: int = 512, seed: int = 1234, *,
                            variable_length: bool = False, min_length: int = 32, with_token_ids: bool = False,
                            num_tokens: int = 64000, word_pred_prob: float = 0.15, phoneme_mask_prob: float = 0.8,
                            replace_prob: float = 0.1):
    """Synthetic batch with the statistics of the reference pipeline, in the collate tuple format.

    Mirrors R/dataloader.py:59-126: words of 2..9 symbols drawn from the IPA/latin ids 52..184, each followed by the
    separator (id 186); a word is selected with probability ``word_pred_prob`` and then masked (id 185) with
    probability 0.8, replaced by symbols drawn from the sample itself with 0.1, or kept (and NOT recorded) with
    0.1; separators are never selected; the sequence is cropped to ``max_seq_length``; samples are sorted by
    length (descending) and zero-padded (PAD id 0) to the batch maximum.
    """
    rng = np.random.default_rng(seed)
    samples = []
    for _ in range(batch_size):
        target = max_seq_length if not variable_length else int(rng.integers(min_length, max_seq_length + 1))
        labels: List[int] = []
        masked: List[int] = []
        tok: List[int] = []
        midx: List[int] = []
        while len(labels) < target + 16:
            wl = int(rng.integers(2, 10))
            word = rng.integers(52, 185, size=wl)
            t_id = int(rng.integers(5, num_tokens))
            start = len(labels)
            labels.extend(word.tolist()); labels.append(SEPARATOR_ID)
            tok.extend([t_id] * wl); tok.append(87)  # word_separator placeholder, R/configs/config.yml:26
            if rng.random() < word_pred_prob:
                r = rng.random()
                if r < phoneme_mask_prob:
                    masked.extend([MASK_ID] * wl); midx.extend(range(start, start + wl))
                elif r < phoneme_mask_prob + replace_prob:
                    pool = labels if len(labels) > 0 else word.tolist()
                    masked.extend(int(pool[int(k)]) for k in rng.integers(0, len(pool), size=wl))
                    midx.extend(range(start, start + wl))
                else:
                    masked.extend(word.tolist())
            else:
                masked.extend(word.tolist())
            masked.append(SEPARATOR_ID)
        n = len(labels)
        begin = int(rng.integers(0, n - target)) if n > target else 0
        end = begin + target
        samples.append((torch.tensor(tok[begin:end]), torch.tensor(labels[begin:end]), torch.tensor(masked[begin:end]),
                        [i - begin for i in midx if begin <= i < end]))
    samples.sort(key=lambda s: s[1].shape[0], reverse=True)
    L = samples[0][1].shape[0]
    B = batch_size
    token_ids = torch.zeros((B, L), dtype=torch.long)
    phoneme_labels = torch.zeros((B, L), dtype=torch.long)
    masked_phonemes = torch.zeros((B, L), dtype=torch.long)
    lengths = [0] * B
    masked_indices: List[List[int]] = [[] for _ in range(B)]
    for i, (t, l, m, mi) in enumerate(samples):
        n = l.shape[0]
        token_ids[i, :n], phoneme_labels[i, :n], masked_phonemes[i, :n] = t, l, m
        lengths[i] = n
        masked_indices[i] = mi
    if with_token_ids:
        return token_ids, phoneme_labels, masked_phonemes, lengths, masked_indices
    return phoneme_labels, masked_phonemes, lengths, masked_indices
