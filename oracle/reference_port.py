"""CPU port of the reference training step for TIMING — TEST/BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

Used only by ``bench.py`` (``cpu_baseline`` leg and ``--impl reference``) and by tests.  /root/reference cannot
travel to the GPU box and has no installable package (no setup.py / pyproject), so this restates the few lines of
the reference that sit above its third-party libraries and then executes exactly the library code the reference
executes: HF ``AlbertModel`` (the reference builds it at R/train.py:263-265), ``nn.Linear`` heads (R/model.py:10-11,
15-16, 23, 28), the per-sample loss loop (R/train.py:123-129) with ``nn.CrossEntropyLoss`` (R/train.py:215), and
``torch.optim.AdamW(lr)`` (R/train.py:272) — fp32, all host threads.  ``tests/test_reference_port.py`` checks it
against the golden vectors and the numpy oracle.
"""
from __future__ import annotations

import os
import time
from typing import Optional, Sequence

import torch
from torch import nn

CONFIG_YML_MODEL_PARAMS = dict(hidden_size=768, num_attention_heads=12, intermediate_size=2048,
                               max_position_embeddings=512, num_hidden_layers=12)  # R/configs/config.yml:32-39


class PortedModel(nn.Module):
    """R/model.py:5-30 (MultiTaskModel when num_tokens is given, else PhonemeOnlyModel)."""

    def __init__(self, encoder, num_phonemes: int, num_tokens: Optional[int], hidden_size: int):
        super().__init__()
        self.encoder = encoder
        self.phoneme_predictor = nn.Linear(hidden_size, num_phonemes)
        if num_tokens is not None:
            self.token_predictor = nn.Linear(hidden_size, num_tokens)

    def forward(self, phonemes, attention_mask=None):
        hidden = self.encoder(phonemes, attention_mask=attention_mask).last_hidden_state
        out = [self.phoneme_predictor(hidden)]
        if hasattr(self, "token_predictor"):
            out.append(self.token_predictor(hidden))
        return out


def build(num_phonemes: int = 188, num_tokens: Optional[int] = None, model_params: Optional[dict] = None,
          seed: int = 0) -> PortedModel:
    from transformers import AlbertConfig, AlbertModel
    mp = dict(CONFIG_YML_MODEL_PARAMS if model_params is None else model_params)
    torch.manual_seed(seed)
    enc = AlbertModel(AlbertConfig(vocab_size=num_phonemes, **mp))  # R/train.py:263-265
    return PortedModel(enc, num_phonemes, num_tokens, mp["hidden_size"])


def length_to_mask(lengths: torch.Tensor) -> torch.Tensor:  # R/train.py:34-44
    max_len = int(lengths.max().item())
    positions = torch.arange(max_len).expand(lengths.size(0), max_len)
    return (positions + 1) > lengths.unsqueeze(1)


def masked_loss(pred, labels, input_lengths, masked_indices, criterion):  # R/train.py:107-131
    loss, count = 0, 0
    for p, t, n, m in zip(pred, labels, input_lengths, masked_indices):
        if len(m) > 0:
            loss = loss + criterion(p[:n][m], t[:n][m])
            count += 1
    return loss / count if count > 0 else torch.tensor(0.0, requires_grad=True)


def train_step(model: PortedModel, optimizer, batch: Sequence, criterion=None) -> float:
    """R/train.py:381-390 + 355-357 on CPU: mask, forward, loss (both heads summed when present), backward, AdamW."""
    criterion = criterion or nn.CrossEntropyLoss()
    if len(batch) == 5:
        token_ids, labels, masked, lengths, midx = batch
    else:
        labels, masked, lengths, midx = batch
        token_ids = None
    text_mask = length_to_mask(torch.Tensor(lengths))
    preds = model(masked, attention_mask=(~text_mask).int())
    loss = masked_loss(preds[0], labels, lengths, midx, criterion)
    if len(preds) > 1:
        loss = loss + masked_loss(preds[1], token_ids, lengths, midx, criterion)
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return float(loss.item())


def time_steps(batch, num_tokens: Optional[int], steps: int, warmup: int, threads: Optional[int] = None):
    """Returns (seconds per step, valid tokens per step, threads used)."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = build(num_tokens=num_tokens)
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=7e-5)  # R/train.py:272, R/configs/config.yml:20
    lengths = batch[-2]
    for _ in range(warmup):
        train_step(model, opt, batch)
    t0 = time.perf_counter()
    for _ in range(steps):
        train_step(model, opt, batch)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dt, int(sum(lengths)), threads
