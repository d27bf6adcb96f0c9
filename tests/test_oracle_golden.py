"""The oracle (oracle/plbert_oracle.py) against the golden vectors produced by executing the reference
(oracle/make_golden.py: R/model.py + HF AlbertModel + R/train.py loss + torch AdamW, float64).

Tolerances: float64 vs float64, so only summation order differs -> rel 1e-9 (north star fp32 bar is 1e-5)."""
import os

import numpy as np
import pytest

from oracle import plbert_oracle as O


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    cfgv = z["config"]
    cfg = O.OracleConfig(vocab_size=int(cfgv[0]), embedding_size=int(cfgv[1]), hidden_size=int(cfgv[2]),
                         num_attention_heads=int(cfgv[3]), intermediate_size=int(cfgv[4]),
                         num_hidden_layers=int(cfgv[5]), max_position_embeddings=int(cfgv[6]))
    params = {k[len("param/"):]: z[k].astype(np.float64) for k in z.files if k.startswith("param/")}
    offs = z["midx_offs"]
    midx = [z["midx_flat"][offs[i]:offs[i + 1]].tolist() for i in range(len(offs) - 1)]
    return z, cfg, params, midx


def _rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def test_length_to_mask_bit_exact(golden_dir):
    z = np.load(os.path.join(golden_dir, "length_to_mask.npz"))
    n = len([k for k in z.files if k.startswith("lengths_")])
    assert n >= 4
    for i in range(n):
        lens = z[f"lengths_{i}"]
        assert np.array_equal(O.length_to_mask(lens), z[f"pad_mask_{i}"])
        got = O.attention_mask_from_lengths(lens)
        assert got.dtype == np.int32 and np.array_equal(got, z[f"attn_mask_{i}"])


def test_multitask_step_matches_reference(golden_dir):
    z, cfg, params, midx = _load(golden_dir, "multitask_step.npz")
    assert any(len(m) == 0 for m in midx) and any(len(m) > 0 for m in midx)
    total, losses, grads, hidden = O.loss_and_grads(params, cfg, z["masked"], z["labels"], z["lengths"], midx,
                                                    token_labels=z["token_ids"])
    assert abs(losses["phoneme_predictor"] - z["loss_phoneme"]) <= 1e-10 * abs(z["loss_phoneme"])
    assert abs(losses["token_predictor"] - z["loss_token"]) <= 1e-10 * abs(z["loss_token"])
    # hidden states: the reference defines padded query rows too (computed, never used) -> compare everything
    assert _rel(hidden, z["last_hidden"]) < 1e-10
    logits = O.head_logits(params, hidden, "phoneme_predictor")
    assert _rel(logits, z["phoneme_logits"]) < 1e-10
    ref_grads = {k[len("grad/"):]: z[k] for k in z.files if k.startswith("grad/")}
    assert set(ref_grads) == set(grads), set(ref_grads) ^ set(grads)
    assert {k[len("nograd/"):] for k in z.files if k.startswith("nograd/")} == {"encoder.pooler.weight",
                                                                                 "encoder.pooler.bias"}
    qb = np.linalg.norm(ref_grads[O._LAYER + "attention.query.bias"])
    for name, g in ref_grads.items():
        if name.endswith("attention.key.bias"):
            # true gradient is identically zero (softmax shift invariance) -> both sides are rounding noise
            assert np.linalg.norm(grads[name]) <= 1e-9 * qb and np.linalg.norm(g) <= 1e-9 * qb
            continue
        assert _rel(grads[name], g) < 1e-9, name
    # rows that can never receive gradient
    assert not grads[O._EMB + "word_embeddings.weight"][0].any()
    assert not grads[O._EMB + "token_type_embeddings.weight"][1].any()
    L = z["masked"].shape[1]
    assert not grads[O._EMB + "position_embeddings.weight"][L:].any()


def test_adamw_two_steps_match_reference(golden_dir):
    z, cfg, params, midx = _load(golden_dir, "multitask_step.npz")
    state = {}
    for _ in range(2):
        total, _, grads, _ = O.loss_and_grads(params, cfg, z["masked"], z["labels"], z["lengths"], midx,
                                              token_labels=z["token_ids"])
        O.adamw_step(params, grads, state)
    assert abs(total - z["loss_step2"]) <= 1e-9 * abs(z["loss_step2"])
    for k in z.files:
        if not k.startswith("after2/"):
            continue
        name = k[len("after2/"):]
        # AdamW's first steps move every weight by ~lr regardless of gradient size, so compare the UPDATE
        upd_ref = z[k] - z["param/" + name]
        upd = params[name] - z["param/" + name]
        if "pooler" in name:
            assert not upd.any() and not upd_ref.any()  # no grad -> untouched, no decay (torch skips it)
            continue
        if name.endswith("attention.key.bias"):
            continue  # sign(noise) updates: gradient is rounding noise on both sides
        assert _rel(upd, upd_ref) < 1e-6, name
    # optimizer state exists exactly for the parameters that received a gradient
    names = list(O.param_shapes(cfg, int(z["config"][0]), int(z["config"][7])).keys())
    with_state = sorted(names.index(n) for n in state)
    assert with_state == z["opt_state_keys"].tolist()


def test_phoneme_only_process_batch(golden_dir):
    z, cfg, params, midx = _load(golden_dir, "phoneme_only_step.npz")
    total, _, grads, _ = O.loss_and_grads(params, cfg, z["masked"], z["labels"], z["lengths"], midx)
    assert abs(total - z["loss"]) <= 1e-10 * abs(z["loss"])
    for k in z.files:
        if k.startswith("grad/") and not k.endswith("attention.key.bias"):
            assert _rel(grads[k[len("grad/"):]], z[k]) < 1e-9, k


def test_all_empty_batch_gives_zero_loss():
    logits = np.random.default_rng(0).normal(size=(2, 3, 7))
    loss, dl = O.calculate_loss(logits, np.zeros((2, 3), dtype=np.int64), [3, 2], [[], []])
    assert loss == 0.0 and not dl.any()


def test_flatten_masked_indices_weights_and_order():
    rows, w, owner = O.flatten_masked_indices([5, 4, 3], [[4, 0], [], [2]], seq_stride=8)
    assert rows.tolist() == [4, 0, 18] and owner.tolist() == [0, 0, 2]
    assert np.allclose(w, [0.25, 0.25, 0.5])
    with pytest.raises(IndexError):
        O.flatten_masked_indices([3], [[3]], seq_stride=8)


def test_float32_oracle_within_north_star_tolerance(golden_dir):
    """fp32 evaluation of the same math stays within the north star's fp32 bar (rel 1e-5 on the loss)."""
    z, cfg, params, midx = _load(golden_dir, "phoneme_only_step.npz")
    p32 = {k: v.astype(np.float32) for k, v in params.items()}
    total, _, grads, _ = O.loss_and_grads(p32, cfg, z["masked"], z["labels"], z["lengths"], midx, dtype=np.float32)
    assert abs(float(total) - z["loss"]) <= 1e-5 * abs(z["loss"])
    for k in z.files:
        if k.startswith("grad/") and not k.endswith("attention.key.bias"):
            assert _rel(grads[k[len("grad/"):]].astype(np.float64), z[k]) < 1e-4, k


def load_full_size_case(golden_dir):
    """Case D (oracle/make_golden.py:full_size_case): the benchmarked architecture evaluated by the reference.  Returns
    (npz, oracle config, float32-rounded parameters as float64, masked index lists)."""
    from oracle.make_golden import subsample_index  # noqa: F401  (same element subsets as the generator)
    z = np.load(os.path.join(golden_dir, "full_size_step.npz"))
    c = [int(x) for x in z["config"]]
    cfg = O.OracleConfig(vocab_size=c[0], embedding_size=c[1], hidden_size=c[2], num_attention_heads=c[3], intermediate_size=c[4],
                         num_hidden_layers=c[5], max_position_embeddings=c[6])
    params = {k: v.astype(np.float32).astype(np.float64)
              for k, v in O.init_params(cfg, c[0], c[7], seed=int(z["param_seed"]), ln_jitter=float(z["ln_jitter"])).items()}
    offs = z["midx_offs"]
    midx = [z["midx_flat"][offs[i]:offs[i + 1]].tolist() for i in range(len(offs) - 1)]
    return z, cfg, params, midx


def test_full_size_step_matches_reference(golden_dir):
    """The oracle at the BENCHMARKED dimensions (H 768, 12 heads, 12 shared-layer applications, L 512, two heads) against the
    reference's own float64 outputs: losses, every gradient tensor's norm and a fixed 8192-element subsample of each."""
    from oracle.make_golden import subsample_index
    z, cfg, params, midx = load_full_size_case(golden_dir)
    assert (cfg.hidden_size, cfg.num_attention_heads, cfg.num_hidden_layers, cfg.intermediate_size) == (768, 12, 12, 2048)
    assert z["masked"].shape == (3, 512) and min(z["lengths"]) < 512 and any(len(m) == 0 for m in midx)
    total, losses, grads, hidden = O.loss_and_grads(params, cfg, z["masked"], z["labels"], z["lengths"], midx,
                                                    token_labels=z["token_ids"])
    assert abs(losses["phoneme_predictor"] - z["loss_phoneme"]) <= 1e-10 * abs(z["loss_phoneme"])
    assert abs(losses["token_predictor"] - z["loss_token"]) <= 1e-10 * abs(z["loss_token"])
    h = hidden.reshape(-1)
    assert _rel(h[subsample_index("last_hidden", h.size)], z["last_hidden_sub"]) < 1e-9
    names = {k[len("gradsub/"):] for k in z.files if k.startswith("gradsub/")}
    assert names == set(grads), names ^ set(grads)
    qb = float(z["gradnorm/" + O._LAYER + "attention.query.bias"])
    for name in sorted(names):
        g = grads[name].reshape(-1)
        if name.endswith("attention.key.bias"):
            assert np.linalg.norm(g) <= 1e-9 * qb and float(z["gradnorm/" + name]) <= 1e-9 * qb
            continue
        assert abs(np.linalg.norm(g) - z["gradnorm/" + name]) <= 1e-9 * z["gradnorm/" + name], name
        assert _rel(g[subsample_index(name, g.size)], z["gradsub/" + name]) < 1e-8, name
