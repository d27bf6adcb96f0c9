"""Input pipeline off the critical path: the reference's per-sample masking + collate, run in background
worker processes, feeding ``BatchFeeder``.

The reference builds every batch on the training thread (``num_workers=0``, R/train.py:253): per sample a Python loop over the
words (whole-word selection, 80/10/10 mask / replace / keep, random 512-crop — R/dataloader.py:35-142) and then the collate
(R/dataloader.py:200-223, 276-297).  At B200 step times (≈55 ms for 192 x 512) that loop is as long as the step itself, so
here it runs in ``num_workers`` processes, one whole batch per task, delivered in order.

* ``MaskedPhonemeDataset`` / ``Collater`` / ``PhonemeOnlyCollater`` — host-side mirrors of the reference classes (same
  constructor arguments, same per-sample and per-batch tuples).  The mirror consumes ``numpy.random`` / ``random`` call for
  call like the reference, so with equal seeds it returns identical tensors (tests/test_data_pipeline.py replays a fixture
  produced by the reference's own classes); it exists because the reference tree cannot travel to the GPU box.  Any other
  map-style dataset with the same item format (the reference's own, when importable) can be passed instead.
* ``CollateWorkerLoader`` — the worker pool.  A task is (epoch, batch index, sample indices); the worker seeds both RNGs from
  (seed, epoch, batch index), builds the samples and collates them, so a batch does not depend on the number of workers or
  on scheduling, and epochs differ like the reference's on-the-fly masking.  Built on ``torch.utils.data.DataLoader`` in
  batch-less mode (``batch_size=None``): in-order delivery, prefetching and worker lifetime are the library's plumbing.
* ``synthetic_records`` — SURVEY.md §8d's generator of dataset rows (``{'phonemes': [...], 'token_ids': [...]}``).
"""
from __future__ import annotations

import random
import string
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

# ---- symbol table (R/char_indexer.py:3-15): PAD, sorted punctuation, IPA letters, latin letters, mask, separator, unknown ----
PAD, PHONEME_MASK, PHONEME_SEPARATOR, UNKNOWN = "P", "M", " ", "U"
_PUNCT_EXTRA = ';:,.!?¡¿—…"«»“”‘’،؛؟٫٬٪﴾﴿ـ'
_IPA = ('ɑɐɒæɓʙβɔɕçɗɖðʤəɘɚɛɜɝɞɟʄɡɠɢʛɦɧħɥʜɨɪʝɭɬɫɮʟɱɯɰŋɳɲɴøɵɸθœɶʘɹɺɾɻʀʁɽʂʃʈʧʉʊʋⱱʌɣɤʍχʎʏʑʐʒʔʡʕʢǀǁǂǃˈˌːˑʼʴʰʱʲʷˠˤ˞'
        '↓↑→↗↘̩ᵻ')


def build_symbols() -> List[str]:
    punctuation = "".join(sorted(set(_PUNCT_EXTRA + string.punctuation)))
    return [PAD] + list(punctuation) + list(_IPA) + list(string.ascii_lowercase) + [PHONEME_MASK, PHONEME_SEPARATOR, UNKNOWN]


symbols = build_symbols()
_INDEX: Dict[str, int] = {s: i for i, s in enumerate(symbols)}
_UNK, _MASK, _SEP = _INDEX[UNKNOWN], _INDEX[PHONEME_MASK], _INDEX[PHONEME_SEPARATOR]


def index_chars(text: str) -> List[int]:
    """``CharacterIndexer.__call__`` (R/char_indexer.py:19-24): symbol ids, unknown characters -> id of 'U'."""
    return [_INDEX.get(ch, _UNK) for ch in text]


class MaskedPhonemeDataset(torch.utils.data.Dataset):
    """Mirror of R/dataloader.py:19-142.  Item: ``(phoneme_labels, masked_phonemes, masked_index)`` or, with
    ``use_token_ids``, ``(token_ids, phoneme_labels, masked_phonemes, masked_index)`` — LongTensors + a python list.

    Works on symbol ids instead of strings (one table lookup per character instead of three), but draws from the global
    ``numpy.random`` / ``random`` generators exactly where the reference does: one ``np.random.rand()`` per word (:85), one
    ``np.random.choice(3, p=...)`` per selected word (:89), ``random.choices`` over the sample's own phonemes for a
    replacement (:94), one ``np.random.randint`` for the crop of an over-long sample (:114)."""

    def __init__(self, dataset, word_pred_prob, phoneme_mask_prob, replace_prob, word_separator, max_seq_length, use_token_ids):
        self.data = dataset
        self.max_seq_length = max_seq_length
        self.word_pred_prob = word_pred_prob
        self.phoneme_mask_prob = phoneme_mask_prob
        self.replace_prob = replace_prob
        self.word_separator = word_separator
        self.use_token_ids = use_token_ids

    def __len__(self) -> int:
        return len(self.data)

    def __getitem__(self, idx):
        words = self.data[idx]["phonemes"]
        word_ids = [index_chars(w) for w in words]
        pool = [i for w in word_ids for i in w]                       # ''.join(words): what a replacement draws from
        tokens = self.data[idx]["token_ids"] if self.use_token_ids else [self.word_separator] * len(words)
        # np.random.choice(3, p=probs) draws ONE uniform and looks it up in the normalised cumulative distribution
        # (side='right'); done by hand here because choice() re-validates p on every call (20 us per selected word)
        cdf = np.cumsum(np.asarray([self.phoneme_mask_prob, self.replace_prob,
                                    1 - (self.phoneme_mask_prob + self.replace_prob)], dtype=np.float64))
        cdf /= cdf[-1]
        c0, c1 = float(cdf[0]), float(cdf[1])
        rand, sample = np.random.rand, np.random.random_sample
        labels: List[int] = []
        masked: List[int] = []
        out_tokens: List[int] = []
        masked_index: List[int] = []
        for ids, tok in zip(word_ids, tokens):
            n = len(ids)
            start = len(masked)
            out_tokens.extend([tok] * n)
            out_tokens.append(self.word_separator)
            labels.extend(ids)
            labels.append(_SEP)
            if rand() < self.word_pred_prob:
                u = sample()
                action = 0 if u < c0 else (1 if u < c1 else 2)          # 0 mask, 1 replace, 2 keep (keep is NOT recorded)
                if action == 1:
                    masked.extend(random.choices(pool, k=n))
                elif action == 0:
                    masked.extend([_MASK] * n)
                else:
                    masked.extend(ids)
                if action != 2:
                    masked_index.extend(range(start, start + n))
            else:
                masked.extend(ids)
            masked.append(_SEP)
        length = len(masked)
        if length > self.max_seq_length:
            begin = int(np.random.randint(0, length - self.max_seq_length))
            end = begin + self.max_seq_length
            masked, out_tokens, labels = masked[begin:end], out_tokens[begin:end], labels[begin:end]
            masked_index = [i - begin for i in masked_index if begin <= i < end]
        masked_t, labels_t = torch.LongTensor(masked), torch.LongTensor(labels)
        if self.use_token_ids:
            return torch.LongTensor(out_tokens), labels_t, masked_t, masked_index
        return labels_t, masked_t, masked_index


def _collate(batch, n_tensors: int):
    """Shared body of the two collaters: stable sort by length (descending), zero padding (PAD id 0) to the batch maximum."""
    batch = sorted(batch, key=lambda item: item[0].shape[0], reverse=True)
    width = batch[0][0].shape[0]
    out = [torch.zeros((len(batch), width), dtype=torch.long) for _ in range(n_tensors)]
    lengths, masked_indices = [], []
    for row, item in enumerate(batch):
        n = item[n_tensors - 1].size(0)
        for dst, src in zip(out, item[:n_tensors]):
            dst[row, :n] = src
        lengths.append(n)
        masked_indices.append(item[n_tensors])
    return tuple(out) + (lengths, masked_indices)


class Collater:
    """R/dataloader.py:200-223 -> ``(token_ids, phoneme_labels, masked_phonemes, input_lengths, masked_indices)``."""

    def __call__(self, batch):
        return _collate(batch, 3)


class PhonemeOnlyCollater:
    """R/dataloader.py:276-297 -> ``(phoneme_labels, masked_phonemes, input_lengths, masked_indices)``."""

    def __call__(self, batch):
        return _collate(batch, 2)


def synthetic_records(n: int, seed: int = 1234, num_tokens: int = 64000, min_words: int = 130, max_words: int = 150,
                      min_word_len: int = 2, max_word_len: int = 9) -> List[dict]:
    """Dataset rows of SURVEY.md §8d: words of 2..9 characters drawn uniformly from the IPA / latin part of the symbol table
    (ids 52..184), one token id ~U{5..num_tokens-1} per word, enough words (>= 120) that every sample exceeds 512 symbols."""
    rng = np.random.default_rng(seed)
    alphabet = np.array(symbols[52:185])
    records = []
    for _ in range(n):
        nw = int(rng.integers(min_words, max_words + 1))
        lens = rng.integers(min_word_len, max_word_len + 1, size=nw)
        chars = alphabet[rng.integers(0, alphabet.shape[0], size=int(lens.sum()))]
        cuts = np.cumsum(lens)[:-1]
        records.append({"phonemes": ["".join(w) for w in np.split(chars, cuts)],
                        "token_ids": [int(t) for t in rng.integers(5, num_tokens, size=nw)]})
    return records


class _BatchTasks(torch.utils.data.Dataset):
    """Map-style view of one epoch: item k = the collated batch k, built under a seed derived from (seed, epoch, k)."""

    def __init__(self, dataset, collate_fn, seed: int):
        self.dataset, self.collate_fn, self.seed = dataset, collate_fn, seed

    def __getitem__(self, task):
        epoch, k, indices = task
        s = (self.seed * 1000003 + epoch * 7919 + k) % (2 ** 32)
        np.random.seed(s)
        random.seed(s)
        return self.collate_fn([self.dataset[i] for i in indices])


class _TaskSampler:
    """Sampler over batch tasks.  ``epochs`` = 1: the tasks of the loader's next epoch (advances the epoch counter every time
    it is iterated); ``epochs`` = None: all epochs back to back, endlessly (keeps the workers busy across epoch boundaries)."""

    def __init__(self, owner: "CollateWorkerLoader", endless: bool):
        self.owner, self.endless = owner, endless

    def __iter__(self):
        while True:
            tasks = self.owner._epoch_tasks(self.owner.epoch)
            self.owner.epoch += 1
            yield from tasks
            if not self.endless:
                return

    def __len__(self) -> int:
        return len(self.owner)


class CollateWorkerLoader:
    """Epoch-style iterable of collate tuples built by ``num_workers`` background processes (0 = in the calling thread,
    same batches).  ``shuffle`` / ``drop_last`` as in R/dataloader.py:244-273 (train: True / True; validation: False / False).
    The worker processes are started once and live as long as the loader (``persistent_workers``).

        loader = CollateWorkerLoader(MaskedPhonemeDataset(records, **dataset_params, use_token_ids=True), 192, Collater(), 8)
        feed = BatchFeeder(model, loader)            # stages batch i+1 (flatten + pinned H2D) under step i

    ``endless()`` iterates over all epochs back to back without draining the workers at the epoch boundaries."""

    def __init__(self, dataset, batch_size: int, collate_fn, num_workers: int = 4, shuffle: bool = True, drop_last: bool = True,
                 seed: int = 1234, prefetch_factor: int = 2):
        self.dataset, self.batch_size, self.shuffle, self.drop_last, self.seed = dataset, batch_size, shuffle, drop_last, seed
        self.num_workers = num_workers
        self._tasks = _BatchTasks(dataset, collate_fn, seed)
        self._kw = dict(batch_size=None, shuffle=False, num_workers=num_workers)
        if num_workers > 0:
            self._kw.update(prefetch_factor=prefetch_factor, persistent_workers=True)
        self._dl = None   # created on first use: worker start-up is paid once
        self.epoch = 0

    def __len__(self) -> int:
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _epoch_tasks(self, epoch: int):
        n = len(self.dataset)
        order = np.random.default_rng([self.seed, epoch]).permutation(n) if self.shuffle else np.arange(n)
        return [(epoch, k, order[k * self.batch_size:(k + 1) * self.batch_size].tolist()) for k in range(len(self))]

    def __iter__(self):
        # the sampler yields whole tasks; with batch_size=None the DataLoader hands each one to dataset.__getitem__ in a worker
        # and returns the results in task order
        if self._dl is None:
            self._dl = torch.utils.data.DataLoader(self._tasks, sampler=_TaskSampler(self, endless=False), **self._kw)
        return iter(self._dl)

    def endless(self):
        return iter(torch.utils.data.DataLoader(self._tasks, sampler=_TaskSampler(self, endless=True), **self._kw))


def build_dataloader(df, batch_size, device, dataset_config, use_token_ids, num_workers=4, seed: int = 1234, **kwargs):
    """R/dataloader.py:225-274 with background workers: 5 % (at most 10 000) of the rows held out for validation, train
    loader shuffled with ``drop_last``, validation in order without.  Returns ``(train_loader, val_loader)``."""
    dataset = MaskedPhonemeDataset(df, use_token_ids=use_token_ids, **dataset_config)
    total = len(dataset)
    val_size = min(int(total * 0.05), 10000)
    order = np.random.default_rng(seed).permutation(total).tolist()
    train_set = torch.utils.data.Subset(dataset, order[:total - val_size])
    val_set = torch.utils.data.Subset(dataset, order[total - val_size:])
    collate = Collater() if use_token_ids else PhonemeOnlyCollater()
    return (CollateWorkerLoader(train_set, batch_size, collate, num_workers, shuffle=True, drop_last=True, seed=seed),
            CollateWorkerLoader(val_set, batch_size, collate, num_workers, shuffle=False, drop_last=False, seed=seed + 1))
