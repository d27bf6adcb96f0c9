"""Host-side logic on CPU: row flattening (bit-exact vs the oracle and the golden collate fixture), arena layout,
synthetic batch contract, checkpoint file discovery, FusedAdamW state-dict layout, no-fallback behaviour."""
import os

import numpy as np
import pytest
import torch

from oracle import plbert_oracle as O
from plbert_b200 import batch as B
from plbert_b200.engine import ALIGN, EngineConfig, ParamLayout


def test_flatten_matches_oracle_and_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "collate_contract.npz"))
    offs = z["midx_offs"]
    midx = [z["midx_flat"][offs[i]:offs[i + 1]].tolist() for i in range(len(offs) - 1)]
    L = z["masked"].shape[1]
    rows, w = B.flatten_masked_indices(z["lengths"], midx, L)
    orows, ow, _ = O.flatten_masked_indices(z["lengths"], midx, L)
    assert rows.dtype == np.int32 and np.array_equal(rows, orows)
    assert np.array_equal(w, ow.astype(np.float32))
    assert (B.PAD_ID, B.MASK_ID, B.SEPARATOR_ID, B.NUM_SYMBOLS) == (int(z["pad_id"]), int(z["mask_id"]), int(z["sep_id"]),
                                                                   int(z["num_symbols"]))
    # every selected row holds a label that is a real phoneme (never PAD / separator), as in the reference pipeline
    labels = z["labels"].reshape(-1)[rows]
    assert (labels != B.PAD_ID).all() and (labels != B.SEPARATOR_ID).all()


def test_flatten_edge_cases():
    rows, w = B.flatten_masked_indices([4, 3], [[], []], 8)
    assert rows.size == 0 and w.size == 0
    rows, w = B.flatten_masked_indices([4, 3], [[-1], [0, 0]], 8)  # negative index and duplicates, as torch indexing allows
    assert rows.tolist() == [3, 8, 8] and np.allclose(w, [0.5, 0.25, 0.25])
    with pytest.raises(IndexError):
        B.flatten_masked_indices([4], [[4]], 8)
    # packed layout: per-sample first rows instead of a common stride
    rows, w = B.flatten_masked_indices([4, 3, 5], [[1, -1], [], [0]], np.array([0, 4, 7]))
    assert rows.tolist() == [1, 3, 7] and np.allclose(w, [0.25, 0.25, 0.5])


def test_flatten_all_valid_rows_weights_and_order():
    """Row set of the upstream word loss: every valid position, sample by sample; weights sum to 1 (mean of per-sample means)."""
    from plbert_b200.batch import flatten_all_valid
    lengths, L = [5, 3, 1], 7
    rows, label_rows, w = flatten_all_valid(lengths, L, L)
    want = [b * L + i for b, n in enumerate(lengths) for i in range(n)]
    assert rows.tolist() == want and label_rows.tolist() == want and rows.dtype == np.int32
    assert np.allclose(w, [1 / (3 * n) for n in lengths for _ in range(n)]) and abs(w.sum() - 1.0) < 1e-6
    starts = np.array([0, 5, 8])                      # packed layout: sequences back to back
    prow, plab, pw = flatten_all_valid(lengths, starts, L)
    assert prow.tolist() == list(range(9)) and plab.tolist() == want and np.array_equal(pw, w)


def test_synthetic_batch_contract_and_statistics():
    tok, lab, msk, lens, midx = B.synthetic_collate_batch(64, 512, seed=5, with_token_ids=True)
    assert lab.dtype == torch.int64 and lab.shape == msk.shape == tok.shape == (64, 512)
    assert lens == sorted(lens, reverse=True) and all(l == 512 for l in lens)
    n_rows = sum(len(m) for m in midx)
    frac = n_rows / (64 * 512)
    assert 0.08 < frac < 0.16  # 15 % of words x 90 % recorded, separators never selected (~11.6-13.8 %)
    for b, m in enumerate(midx):
        m = np.asarray(m)
        assert len(set(m.tolist())) == len(m) and (m >= 0).all() and (m < lens[b]).all()
        assert (lab[b][m] != B.SEPARATOR_ID).all()
    masked_pos = msk == B.MASK_ID
    assert masked_pos.sum() > 0.6 * n_rows  # 80 % of the recorded words are replaced by the mask symbol
    # variable-length: sorted by length, zero padded
    lab, msk, lens, midx = B.synthetic_collate_batch(16, 512, seed=6, variable_length=True, min_length=32)
    assert lens == sorted(lens, reverse=True) and min(lens) >= 32 and msk.shape[1] == max(lens)
    for b, l in enumerate(lens):
        assert not msk[b, l:].any() and not lab[b, l:].any()


def _layout(num_tokens=64000):
    return ParamLayout(EngineConfig(188, 128, 768, 12, 2048, 12, 512, 2, 1e-12, 0, 188, num_tokens))


def test_param_layout_matches_reference_state_dict():
    lay = _layout()
    shapes = O.param_shapes(O.OracleConfig(), 188, 64000)
    assert set(lay.shapes) == set(shapes) and all(tuple(lay.shapes[k]) == tuple(v) for k, v in shapes.items())
    # fused QKV views need the three tensors adjacent, in q|k|v order
    a = O._LAYER + "attention."
    assert lay.offsets[a + "key.weight"] == lay.offsets[a + "query.weight"] + 768 * 768
    assert lay.offsets[a + "value.weight"] == lay.offsets[a + "key.weight"] + 768 * 768
    assert lay.offsets[a + "value.bias"] == lay.offsets[a + "query.bias"] + 2 * 768
    # no overlap, every group aligned for TMA (16-byte) and vector access
    spans = sorted((o, o + lay.numel(n)) for n, o in lay.offsets.items())
    assert all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1))
    assert lay.offsets["phoneme_predictor.weight"] % ALIGN == 0 and lay.offsets[a + "dense.weight"] % ALIGN == 0
    # readiness order: heads, then the shared layer + embeddings, then the frozen pooler
    assert lay.offsets["token_predictor.weight"] < lay.head_end <= lay.offsets[a + "query.weight"]
    assert lay.trainable_end <= lay.offsets["encoder.pooler.weight"] and lay.total % 4 == 0
    assert _layout(None).head_end < lay.head_end


def test_find_latest_checkpoint(tmp_path):
    from plbert_b200.train import find_latest_checkpoint
    assert find_latest_checkpoint(str(tmp_path)) == (False, 0)
    for name in ("step_1000.t7", "step_12000.pth", "step_bad.pth", "config.yml"):
        (tmp_path / name).write_text("x")
    assert find_latest_checkpoint(str(tmp_path)) == (True, 12000)
    assert find_latest_checkpoint(str(tmp_path / "missing")) == (False, 0)


def test_state_dict_key_normalisation():
    from plbert_b200 import normalize_state_dict_keys
    sd = {"module.encoder.pooler.bias": 1, "module.mask_predictor.weight": 2, "word_predictor.bias": 3, "phoneme_predictor.bias": 4}
    assert normalize_state_dict_keys(sd) == {"encoder.pooler.bias": 1, "phoneme_predictor.weight": 2, "token_predictor.bias": 3,
                                             "phoneme_predictor.bias": 4}


def test_drop_in_constructor_and_state_dict_and_no_cpu_fallback(tmp_path):
    from transformers import AlbertConfig, AlbertModel
    import plbert_b200 as P
    model_params = dict(hidden_size=128, num_attention_heads=2, intermediate_size=256, max_position_embeddings=64,
                        num_hidden_layers=2, embedding_size=32, pretrained_model="", dropout=0.1)  # extra keys as in config.yml
    enc = AlbertModel(AlbertConfig(vocab_size=188, **model_params))
    model = P.MultiTaskModel(enc, num_phonemes=188, num_tokens=300, hidden_size=128)
    want = O.param_shapes(O.OracleConfig(188, 32, 128, 2, 256, 2, 64), 188, 300)
    got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert got == {k: tuple(v) for k, v in want.items()}
    assert list(got) == list(want)  # parameter order = optimizer-state index space
    model.encoder.save_pretrained(str(tmp_path))  # R/convert_to_hf.py:49
    again = AlbertModel.from_pretrained(str(tmp_path))
    assert torch.equal(again.embeddings.word_embeddings.weight, enc.embeddings.word_embeddings.weight)
    with pytest.raises(P.PlbError):
        model(torch.zeros(1, 8, dtype=torch.long))  # CPU tensors: loud failure, never a PyTorch fallback
    with pytest.raises(P.PlbError):
        P.FusedAdamW(model.parameters(), lr=7e-5).step()


def test_unsupported_configs_are_rejected():
    from transformers import AlbertConfig
    base = dict(hidden_size=128, num_attention_heads=2, intermediate_size=256, embedding_size=32)
    for kw in (dict(hidden_act="gelu"), dict(hidden_dropout_prob=1.0), dict(num_attention_heads=4), dict(num_hidden_groups=2)):
        cfg = AlbertConfig(vocab_size=188, **{**base, **kw})
        with pytest.raises(Exception):
            EngineConfig.from_albert(cfg, 188, None)
    # dropout is a kernel parameter (identity at the reference's p = 0): probabilities in [0, 1) are carried into the engine
    ec = EngineConfig.from_albert(AlbertConfig(vocab_size=188, hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.2, **base), 188, None)
    assert (ec.hidden_dropout_prob, ec.attention_probs_dropout_prob) == (0.1, 0.2)


def test_fused_albert_keeps_the_container_and_fails_loudly_on_cpu():
    """``fuse_albert`` only swaps ``forward`` on the caller's AlbertModel instance: names, shapes, ``save_pretrained`` inputs
    stay; the stock implementation stays reachable; a CPU call raises instead of silently computing through PyTorch."""
    import pytest
    import torch
    from transformers import AlbertConfig, AlbertModel
    import plbert_b200 as P
    cfg = AlbertConfig(vocab_size=188, embedding_size=32, hidden_size=128, num_attention_heads=2, intermediate_size=256,
                       num_hidden_layers=2, max_position_embeddings=64)
    torch.manual_seed(0)
    bert = AlbertModel(cfg)
    keys = list(bert.state_dict().keys())
    assert P.fuse_albert(bert) is bert and P.FusedAlbertEncoder(bert) is bert       # idempotent
    assert list(bert.state_dict().keys()) == keys                                     # no wrapper / head leaks in
    ids = torch.randint(1, 185, (2, 16))
    with pytest.raises(P.PlbError):
        bert(ids, attention_mask=torch.ones_like(ids))
    out = bert.hf_forward(ids, attention_mask=torch.ones_like(ids))                   # the stock path is still there
    assert out.last_hidden_state.shape == (2, 16, 128)
    # the drop-in models fuse their encoder and report themselves as its owner
    model = P.PhonemeOnlyModel(AlbertModel(cfg), 188, 128)
    assert model.encoder._plb_fused and model.encoder._plb_owner is model
    assert set(model.state_dict()) == {"encoder." + k for k in keys} | {"phoneme_predictor.weight", "phoneme_predictor.bias"}


def test_dropout_rng_replica_statistics():
    """oracle/dropout_rng.py (the numpy replica of csrc/rng.cuh): keep rate, scale, independence of streams, pair layout."""
    import numpy as np
    from oracle import dropout_rng as R
    m = R.hidden_mask(512, 768, 0.1, 1234, 3)
    assert set(np.unique(m)) == {0.0, 1 / 0.9}
    assert abs((m != 0).mean() - 0.9) < 3e-3
    assert abs(m.mean() - 1.0) < 4e-3
    m2 = R.hidden_mask(512, 768, 0.1, 1234, 4)
    agree = ((m != 0) == (m2 != 0)).mean()
    assert abs(agree - (0.81 + 0.01)) < 5e-3          # independent masks agree with probability 0.9^2 + 0.1^2
    assert abs(np.corrcoef(m[:, :-1].ravel() != 0, m[:, 1:].ravel() != 0)[0, 1]) < 0.01   # the two halves of a pair are independent
    a = R.attention_mask(6, 40, 33, 0.25, 99, 1)
    assert a.shape == (6, 40, 33) and abs((a != 0).mean() - 0.75) < 0.02
    assert np.array_equal(R.hidden_mask(4, 8, 0.0, 1, 1), np.ones((4, 8)))
