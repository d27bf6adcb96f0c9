"""Host-side input pipeline: the mirror of the reference's dataset / collate classes against outputs of the
reference's own classes (tests/golden/dataset_pipeline.npz, oracle/make_golden.py:dataset_case), and the worker pool's
ordering and determinism."""
import os
import random

import numpy as np
import pytest
import torch

from plbert_b200 import data as D


def _ragged(z, name):
    offs = z[name + "/midx_offs"]
    return [z[name + "/midx_flat"][offs[i]:offs[i + 1]].tolist() for i in range(len(offs) - 1)]


def test_symbol_table_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "dataset_pipeline.npz"))
    assert D.symbols == z["symbols"].tolist() and len(D.symbols) == 188
    assert (D.symbols[0], D.symbols[185], D.symbols[186], D.symbols[187]) == ("P", "M", " ", "U")
    assert D.index_chars("a M☃") == [159, 186, 185, 187]            # unknown characters map to 'U'


@pytest.mark.parametrize("name,use_tok,max_len,words", [("short", True, 48, (3, 12)), ("long", False, 512, (100, 150))])
def test_dataset_and_collate_equal_the_reference_bit_for_bit(golden_dir, name, use_tok, max_len, words):
    """Same records, same ``np.random`` / ``random`` seeds -> identical tensors, lengths and masked index lists: the mirror
    draws random numbers exactly where the reference does (selection, action, replacement, crop)."""
    z = np.load(os.path.join(golden_dir, "dataset_pipeline.npz"))
    recs = D.synthetic_records(6, seed=5, num_tokens=997, min_words=words[0], max_words=words[1])
    ds = D.MaskedPhonemeDataset(recs, word_pred_prob=0.3, phoneme_mask_prob=0.8, replace_prob=0.1, word_separator=87,
                                max_seq_length=max_len, use_token_ids=use_tok)
    np.random.seed(99)
    random.seed(99)
    items = [ds[i] for i in range(len(ds))]
    *tensors, lengths, midx = (D.Collater() if use_tok else D.PhonemeOnlyCollater())(items)
    assert len(tensors) == (3 if use_tok else 2)
    for k, t in enumerate(tensors):
        assert t.dtype == torch.int64 and np.array_equal(t.numpy(), z[f"{name}/tensor{k}"]), k
    assert lengths == z[name + "/lengths"].tolist() and [list(m) for m in midx] == _ragged(z, name)
    assert sorted(lengths, reverse=True) == lengths and max(lengths) <= max_len
    if name == "long":
        assert lengths == [512] * 6                                       # every sample was cropped


def test_reference_classes_live_when_present():
    """In the build container the reference tree is importable: compare against it directly on fresh seeds."""
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present (GPU box)")
    import sys
    sys.path.insert(0, ref)
    import dataloader as R
    recs = D.synthetic_records(5, seed=11, num_tokens=500, min_words=20, max_words=60)
    kw = dict(word_pred_prob=0.15, phoneme_mask_prob=0.8, replace_prob=0.1, word_separator=87, max_seq_length=128, use_token_ids=True)
    outs = []
    for mod in (R, D):
        ds = mod.MaskedPhonemeDataset(recs, **kw)
        np.random.seed(3)
        random.seed(3)
        outs.append(mod.Collater()([ds[i] for i in range(len(ds))]))
    for a, b in zip(outs[0][:3], outs[1][:3]):
        assert torch.equal(a, b)
    assert list(outs[0][3]) == list(outs[1][3]) and [list(x) for x in outs[0][4]] == [list(x) for x in outs[1][4]]


def _epoch(loader):
    return [tuple(t.clone() if torch.is_tensor(t) else t for t in b) for b in loader]


@pytest.mark.timeout(300)
def test_worker_pool_order_and_determinism():
    """Batches do not depend on the number of workers or on scheduling, arrive in batch order, differ between epochs (masking
    is redrawn like the reference's on-the-fly masking) and repeat exactly for the same (seed, epoch)."""
    recs = D.synthetic_records(37, seed=2, num_tokens=300, min_words=10, max_words=40)
    ds = D.MaskedPhonemeDataset(recs, word_pred_prob=0.15, phoneme_mask_prob=0.8, replace_prob=0.1, word_separator=87,
                                max_seq_length=96, use_token_ids=False)
    runs = {}
    for workers in (0, 3):
        loader = D.CollateWorkerLoader(ds, 8, D.PhonemeOnlyCollater(), num_workers=workers, shuffle=True, drop_last=True, seed=7)
        assert len(loader) == 4
        runs[workers] = (_epoch(loader), _epoch(loader))
    for e in range(2):
        for a, b in zip(runs[0][e], runs[3][e]):
            assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and a[2] == b[2] and a[3] == b[3]
    first, second = runs[3]
    assert len(first) == 4 and any(not torch.equal(a[1], b[1]) for a, b in zip(first, second) if a[1].shape == b[1].shape)
    # validation flavour: in order, last batch kept, every sample exactly once
    val = D.CollateWorkerLoader(ds, 8, D.PhonemeOnlyCollater(), num_workers=2, shuffle=False, drop_last=False, seed=7)
    sizes = [len(b[2]) for b in val]
    assert sizes == [8, 8, 8, 8, 5]
    tr, va = D.build_dataloader(recs, 4, "cpu", dict(word_pred_prob=0.15, phoneme_mask_prob=0.8, replace_prob=0.1, word_separator=87,
                                                     max_seq_length=96), use_token_ids=True, num_workers=0)
    assert len(tr) == (37 - 1) // 4 and sum(len(b[3]) for b in va) == 1 and len(next(iter(tr))) == 5
