"""End-to-end parity of the CUDA training step (public API -> C ABI -> sm_100a kernels).

Compared against (1) the golden vectors produced by executing the reference (tests/golden/cuda_config_step.npz,
see oracle/make_golden.py) and (2) the oracle on fresh seeded inputs.  Tolerances are the north star's bf16 bars:
loss rel <= 2e-2, per-tensor gradient cosine >= 0.999; ``attention.key.bias`` is exempt because its true gradient is
identically zero (softmax shift invariance) and is asserted to be negligible instead; index / gather
work is bit-exact (rows that cannot receive gradient are exactly zero)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import plbert_oracle as O  # noqa: E402  (the checker; never on the product path)

LAYER = O._LAYER
EMB = O._EMB


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def build_model(cfgv, two_heads, dev, params=None, seed=0):
    from transformers import AlbertConfig, AlbertModel
    import plbert_b200 as P
    Vp, E, H, nh, I, nl, maxpos, Vt = [int(x) for x in cfgv]
    torch.manual_seed(seed)
    enc = AlbertModel(AlbertConfig(vocab_size=Vp, embedding_size=E, hidden_size=H, num_attention_heads=nh,
                                   intermediate_size=I, num_hidden_layers=nl, max_position_embeddings=maxpos))
    model = P.MultiTaskModel(enc, Vp, Vt, H) if two_heads else P.PhonemeOnlyModel(enc, Vp, H)
    if params is not None:
        missing, unexpected = model.load_state_dict({k: torch.from_numpy(np.asarray(v, dtype=np.float32)) for k, v in params.items()},
                                                    strict=False)
        assert not unexpected and all("position_ids" in m or "token_type_ids" in m for m in missing)
    else:
        with torch.no_grad():
            for n, p in model.named_parameters():
                if n.endswith("bias") or "LayerNorm" in n or "layer_norm" in n:
                    p.add_(0.05 * torch.randn_like(p))
    return model.to(dev)


def oracle_cfg(cfgv):
    return O.OracleConfig(vocab_size=int(cfgv[0]), embedding_size=int(cfgv[1]), hidden_size=int(cfgv[2]),
                          num_attention_heads=int(cfgv[3]), intermediate_size=int(cfgv[4]),
                          num_hidden_layers=int(cfgv[5]), max_position_embeddings=int(cfgv[6]))


def cosine(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-300))


def check_grads(model, ref_grads, L):
    got = {n: p.grad.detach().float().cpu().numpy() for n, p in model.named_parameters() if p.grad is not None}
    assert set(got) == set(ref_grads), set(got) ^ set(ref_grads)
    assert model.encoder.pooler.weight.grad is None and model.encoder.pooler.bias.grad is None
    qb = np.linalg.norm(ref_grads[LAYER + "attention.query.bias"])
    worst = 1.0
    for n, g in ref_grads.items():
        if n.endswith("attention.key.bias"):
            assert np.linalg.norm(got[n]) <= 0.05 * qb, "key.bias gradient must be (numerically) zero"
            continue
        c = cosine(got[n], g)
        worst = min(worst, c)
        assert c >= 0.999, (n, c)
        ratio = np.linalg.norm(got[n]) / np.linalg.norm(g)
        assert 0.97 < ratio < 1.03, (n, ratio)
    # bit-exact index work: rows no token can reach are exactly zero
    assert not got[EMB + "word_embeddings.weight"][0].any()          # padding_idx
    assert not got[EMB + "token_type_embeddings.weight"][1].any()    # only type 0 is used
    assert not got[EMB + "position_embeddings.weight"][L:].any()     # positions beyond the batch length
    return worst


def ragged(z):
    offs = z["midx_offs"]
    return [z["midx_flat"][offs[i]:offs[i + 1]].tolist() for i in range(len(offs) - 1)]


def test_step_matches_reference_golden(dev, golden_dir):
    """CUDA path vs outputs of the reference itself (R/model.py + HF AlbertModel + R/train.py loss, float64)."""
    z = np.load(os.path.join(golden_dir, "cuda_config_step.npz"))
    params = {k[len("param/"):]: z[k] for k in z.files if k.startswith("param/")}
    model = build_model(z["config"], True, dev, params)
    midx = ragged(z)
    batch = (torch.from_numpy(z["token_ids"]), torch.from_numpy(z["labels"]), torch.from_numpy(z["masked"]),
             z["lengths"].tolist(), midx)
    loss = model.loss_and_backward(batch)
    ref_loss = float(z["loss_phoneme"] + z["loss_token"])
    assert abs(loss.item() - ref_loss) <= 2e-2 * abs(ref_loss), (loss.item(), ref_loss)
    ref_grads = {k[len("grad/"):]: z[k] for k in z.files if k.startswith("grad/")}
    worst = check_grads(model, ref_grads, z["masked"].shape[1])
    print(f"golden: loss {loss.item():.5f} vs {ref_loss:.5f}; worst gradient cosine {worst:.6f}")
    # encoder-only forward (config 5 path) on valid positions
    hid = model.encode(torch.from_numpy(z["masked"]).to(dev),
                       torch.from_numpy((np.arange(z["masked"].shape[1])[None, :] < z["lengths"][:, None]).astype(np.int32)).to(dev))
    for b, n in enumerate(z["lengths"]):
        assert cosine(hid[b, :n].float().cpu().numpy(), z["last_hidden"][b, :n]) > 0.9995


@pytest.mark.parametrize("two_heads,varlen", [(False, False), (True, True), (False, True)])
def test_step_matches_oracle(dev, two_heads, varlen):
    from plbert_b200.batch import synthetic_collate_batch
    import plbert_b200 as P
    cfgv = [188, 32, 128, 2, 256, 4, 128, 211]
    model = build_model(cfgv, two_heads, dev, seed=3)
    batch = synthetic_collate_batch(5, 128, seed=11, variable_length=varlen, min_length=9, with_token_ids=two_heads,
                                    num_tokens=211, word_pred_prob=0.25)
    if two_heads:
        token_ids, labels, masked, lengths, midx = batch
    else:
        labels, masked, lengths, midx = batch
        token_ids = None
    midx[1] = []  # a sample without masked rows is skipped by the loss (R/train.py:124)
    params = {k: v.detach().cpu().double().numpy() for k, v in model.state_dict().items()}
    total, losses, grads, hidden = O.loss_and_grads(params, oracle_cfg(cfgv), masked.numpy(), labels.numpy(), lengths, midx,
                                                    token_labels=None if token_ids is None else token_ids.numpy())
    opt = P.FusedAdamW(model.parameters(), lr=7e-5)
    loss = model.loss_and_backward(batch)
    assert abs(loss.item() - total) <= 2e-2 * abs(total), (loss.item(), total)
    worst = check_grads(model, grads, masked.shape[1])
    print(f"oracle ({'2 heads' if two_heads else '1 head'}, varlen={varlen}): loss {loss.item():.5f} vs {total:.5f}; "
          f"worst cosine {worst:.6f}")
    # AdamW: first-step update is -lr * sign-like(m/sqrt(v)) - lr*wd*p; compare update directions
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    opt.step()
    state = {}
    O.adamw_step(params, grads, state)
    for n, p in model.named_parameters():
        upd = (p.detach() - before[n]).double().cpu().numpy()
        if "pooler" in n:
            assert not upd.any()
            continue
        if n.endswith("attention.key.bias"):
            continue
        ref_upd = params[n] - before[n].double().cpu().numpy()
        # step-1 Adam normalises every element to ~lr * sign(g): elements whose gradient is below the bf16 noise floor
        # may flip, so this is a direction check only (the kernel itself is pinned in test_kernels_gpu.py)
        assert cosine(upd, ref_upd) > 0.85, n
        assert np.abs(upd).max() <= 7e-5 * 1.05 + 7e-5 * 0.01 * np.abs(before[n].cpu().numpy()).max(), n
    sd = opt.state_dict()
    assert len(sd["state"]) == len(list(model.parameters())) - 2  # no state for the pooler (never gets a gradient)
    assert all(set(v) == {"step", "exp_avg", "exp_avg_sq"} for v in sd["state"].values())


@pytest.mark.parametrize("varlen", [False, True])
def test_step_with_dropout_matches_oracle_with_the_same_masks(dev, varlen):
    """hidden_dropout_prob / attention_probs_dropout_prob > 0 (HF :109, :134, :201; the reference config leaves them at 0): the
    step must equal the fp64 oracle run with the SAME masks, rebuilt by the numpy replica of the kernels' generator — in the
    padded and in the packed row layout; a second step draws new masks; eval() switches dropout off."""
    from transformers import AlbertConfig, AlbertModel
    from plbert_b200.batch import synthetic_collate_batch
    from oracle import dropout_rng as R
    import plbert_b200 as P
    Vp, E, H, nh, I, nl, L, p_h, p_a = 188, 32, 128, 2, 256, 3, 128, 0.1, 0.15
    torch.manual_seed(5)
    enc = AlbertModel(AlbertConfig(vocab_size=Vp, embedding_size=E, hidden_size=H, num_attention_heads=nh, intermediate_size=I,
                                   num_hidden_layers=nl, max_position_embeddings=L, hidden_dropout_prob=p_h,
                                   attention_probs_dropout_prob=p_a))
    model = P.PhonemeOnlyModel(enc, Vp, H).to(dev)
    model.train()
    batch = synthetic_collate_batch(4, L, seed=21, variable_length=varlen, min_length=40, word_pred_prob=0.25)
    labels, masked, lengths, midx = batch
    B, Lb = masked.shape
    params = {k: v.detach().cpu().double().numpy() for k, v in model.state_dict().items()}
    cfg = O.OracleConfig(vocab_size=Vp, embedding_size=E, hidden_size=H, num_attention_heads=nh, intermediate_size=I,
                         num_hidden_layers=nl, max_position_embeddings=L)

    def masks(step):
        eng = model._engine
        seed, base = eng.dropout_seed, 64 * step
        packed = varlen                                    # loss_and_backward packs variable-length batches
        starts = np.concatenate([[0], np.cumsum(lengths)[:-1]]) if packed else np.arange(B) * Lb
        T = int(sum(lengths)) if packed else B * Lb

        def rows(cols, stream):                            # [T, cols] in the engine's row order -> [B, Lb, cols], padding rows = 1
            m = R.hidden_mask(T, cols, p_h, seed, stream)
            out = np.ones((B, Lb, cols))
            for b in range(B):
                n = lengths[b] if packed else Lb
                out[b, :n] = m[starts[b]:starts[b] + n]
            return out
        attn = [R.attention_mask(B * nh, Lb, Lb, p_a, seed, base + 1 + 2 * l).reshape(B, nh, Lb, Lb) for l in range(nl)]
        return dict(emb=rows(E, base), out=[rows(H, base + 2 + 2 * l) for l in range(nl)], attn=attn)

    model._ensure_engine(dev)
    losses = []
    for step in range(2):
        total, _, grads, _ = O.loss_and_grads(params, cfg, masked.numpy(), labels.numpy(), lengths, midx, dropout=masks(step))
        loss = model.loss_and_backward(batch)
        assert abs(loss.item() - total) <= 2e-2 * abs(total), (step, loss.item(), total)
        worst = check_grads(model, grads, Lb)
        losses.append(loss.item())
        print(f"dropout step {step} (varlen={varlen}): loss {loss.item():.5f} vs {total:.5f}; worst cosine {worst:.6f}")
    assert abs(losses[0] - losses[1]) > 1e-4               # same weights, same batch, new masks
    total0, _, _, _ = O.loss_and_grads(params, cfg, masked.numpy(), labels.numpy(), lengths, midx)
    assert abs(losses[0] - total0) > 1e-3 * abs(total0)   # and dropout really changed the loss
    model.eval()
    l_eval = model.loss(batch).item()
    assert abs(l_eval - total0) <= 2e-2 * abs(total0)
    model.train()


def test_training_trajectory_tracks_oracle(dev):
    """Eight optimizer steps (lr 1e-3 so that the weights really move) on eight different batches: the CUDA path's loss at
    every step stays within the north star's bf16 bar (rel 2e-2) of the fp64 oracle trained on the same batches from the
    same weights — forward, backward, in-place shared-layer accumulation and AdamW compose correctly over time."""
    from plbert_b200.batch import BatchFeeder, synthetic_collate_batch
    import plbert_b200 as P
    cfgv = [188, 32, 128, 2, 256, 3, 128, 211]
    model = build_model(cfgv, True, dev, seed=13)
    params = {k: v.detach().cpu().double().numpy() for k, v in model.state_dict().items()}
    model_init = {k: v.detach().clone() for k, v in model.state_dict().items()}
    batches = [synthetic_collate_batch(4, 128, seed=50 + i, variable_length=(i % 2 == 1), min_length=16, with_token_ids=True,
                                       num_tokens=211, word_pred_prob=0.3) for i in range(8)]
    opt = P.FusedAdamW(model.parameters(), lr=1e-3)
    feed = BatchFeeder(model, batches)
    got, want, state = [], [], {}
    for tok, lab, msk, lens, midx in batches:
        loss = feed.step()
        opt.step()
        got.append(loss.item())
        total, _, grads, _ = O.loss_and_grads(params, oracle_cfg(cfgv), msk.numpy(), lab.numpy(), lens, midx, token_labels=tok.numpy())
        O.adamw_step(params, grads, state, lr=1e-3)
        want.append(float(total))
    moved = max(float(np.abs(params[n] - v.detach().cpu().double().numpy()).max()) for n, v in model_init.items())
    assert moved > 4e-3, moved                                  # eight steps at lr 1e-3 really moved the weights
    for i, (a, b) in enumerate(zip(got, want)):
        assert abs(a - b) <= 2e-2 * abs(b), (i, got, want)
    # the weights after eight steps agree as well (AdamW normalises tiny gradients to +-lr, hence the loose direction bar)
    for n, p in model.named_parameters():
        if "pooler" in n or n.endswith("attention.key.bias"):
            continue
        assert cosine(p.detach().cpu().double().numpy(), params[n]) > 0.999, n


@pytest.mark.parametrize("packed", [False, True])
def test_token_head_over_all_valid_positions(dev, packed):
    """``MultiTaskModel.token_rows = "all_valid"``: the token-head loss runs over every valid position (upstream's word
    loss) while the phoneme loss keeps the masked rows; the all-valid head is processed in row chunks."""
    from plbert_b200.batch import synthetic_collate_batch
    cfgv = [188, 32, 128, 2, 256, 3, 128, 211]
    model = build_model(cfgv, True, dev, seed=17)
    model.token_rows = "all_valid"
    model.pack_sequences = packed
    batch = synthetic_collate_batch(5, 128, seed=61, variable_length=True, min_length=9, with_token_ids=True, num_tokens=211,
                                    word_pred_prob=0.25)
    token_ids, labels, masked, lengths, midx = batch
    midx[2] = []
    params = {k: v.detach().cpu().double().numpy() for k, v in model.state_dict().items()}
    total, losses, grads, _ = O.loss_and_grads(params, oracle_cfg(cfgv), masked.numpy(), labels.numpy(), lengths, midx,
                                               token_labels=token_ids.numpy(), token_rows="all_valid")
    model._ensure_engine(dev).ROW_CHUNK = 100           # several chunks, the last one ragged
    loss = model.loss_and_backward(batch)
    assert model.last_batch.M_all == sum(lengths) and model.last_batch.packed == packed
    assert abs(loss.item() - total) <= 2e-2 * abs(total), (loss.item(), total, losses)
    check_grads(model, grads, masked.shape[1])
    assert abs(model.loss(batch).item() - total) <= 2e-2 * abs(total)          # validation path, same row sets
    # all masks empty: the phoneme loss is 0, the token loss still runs
    empty = (token_ids, labels, masked, lengths, [[] for _ in lengths])
    t2, l2, _, _ = O.loss_and_grads(params, oracle_cfg(cfgv), masked.numpy(), labels.numpy(), lengths, empty[4],
                                    token_labels=token_ids.numpy(), token_rows="all_valid")
    got = model.loss_and_backward(empty).item()
    assert abs(got - t2) <= 2e-2 * abs(t2) and abs(t2 - l2["token_predictor"]) < 1e-12


def test_launch_overlap_does_not_change_results(dev):
    """plb_set_launch_overlap: with programmatic dependent launch the next kernel's CTAs start under the previous kernel's tail
    and block in griddepcontrol.wait before their first global access — same loss, same gradients, same weights after AdamW
    as with plain stream-ordered launches (only the fp32 atomics' summation order may differ)."""
    import plbert_b200 as P
    from plbert_b200._abi import call
    from plbert_b200.batch import synthetic_collate_batch
    cfgv = [188, 32, 128, 2, 256, 4, 128, 211]
    batches = [synthetic_collate_batch(6, 128, seed=30 + i, variable_length=bool(i), min_length=9, with_token_ids=True,
                                       num_tokens=211, word_pred_prob=0.25) for i in range(2)]
    out = {}
    try:
        for on in (0, 1):
            call("plb_set_launch_overlap", on)
            model = build_model(cfgv, True, dev, seed=5)
            opt = P.FusedAdamW(model.parameters(), lr=1e-5)   # small steps: elements at the summation-noise floor may step either way
            losses = []
            for step in range(4):
                losses.append(model.loss_and_backward(batches[step % 2]).item())
                if step == 0:
                    grads = {n: p.grad.detach().float().clone() for n, p in model.named_parameters() if p.grad is not None}
                opt.step()
            torch.cuda.synchronize()
            out[on] = (losses, grads, {n: p.detach().float().clone() for n, p in model.named_parameters()})
    finally:
        call("plb_set_launch_overlap", 1)
    for a, b in zip(out[0][0], out[1][0]):
        assert abs(a - b) <= 1e-4 * abs(a), (out[0][0], out[1][0])
    for n, g in out[0][1].items():
        assert (out[1][1][n] - g).abs().max().item() <= 2e-3 * g.abs().max().item() + 1e-7, n
    for n, w in out[0][2].items():
        assert (out[1][2][n] - w).abs().max().item() <= 5e-3 * w.abs().max().item() + 1e-6, n


def test_packed_layout_equals_padded_layout(dev):
    """Variable-length batches run without their padding rows.  The reference computes the padded
    rows and never reads them, so loss and gradients must not change; only the fp32 summation order of the weight
    gradients (split-K over a different row count) may."""
    from plbert_b200.batch import synthetic_collate_batch
    cfgv = [188, 32, 128, 2, 256, 4, 128, 211]
    batch = synthetic_collate_batch(6, 128, seed=21, variable_length=True, min_length=9, with_token_ids=True, num_tokens=211,
                                    word_pred_prob=0.25)
    out = {}
    for packed in (False, True):
        model = build_model(cfgv, True, dev, seed=5)
        model.pack_sequences = packed
        loss = model.loss_and_backward(batch)
        assert model.last_batch.packed == packed
        if packed:
            assert model.last_batch.T == sum(batch[3]) < 6 * 128
        out[packed] = (loss.item(), {n: p.grad.detach().float().clone() for n, p in model.named_parameters() if p.grad is not None})
    assert abs(out[True][0] - out[False][0]) <= 1e-5 * abs(out[False][0])
    for n, g in out[False][1].items():
        ref = g.abs().max().item()
        assert (out[True][1][n] - g).abs().max().item() <= 2e-3 * ref + 1e-7, n


def test_batch_feeder_matches_direct_calls(dev):
    """The look-ahead input pipeline stages batch i+1 while step i runs; it must feed every step the
    same integers as staging each batch right before its own step: the first loss is bit-identical (same weights, and the
    forward pass is deterministic), later ones differ only through the fp32 atomic summation order of the gradients."""
    import plbert_b200 as P
    from plbert_b200.batch import BatchFeeder, synthetic_collate_batch
    cfgv = [188, 32, 128, 2, 256, 2, 128, 211]
    batches = [synthetic_collate_batch(4, 128, seed=30 + i, variable_length=(i == 1), min_length=9, with_token_ids=True,
                                       num_tokens=211, word_pred_prob=0.25) for i in range(3)]
    out = []
    for fed in (False, True):
        model = build_model(cfgv, True, dev, seed=6)
        build_ref = {n: p.detach().clone() for n, p in model.named_parameters()}
        opt = P.FusedAdamW(model.parameters(), lr=7e-5)
        losses = []
        if fed:
            feed = BatchFeeder(model, batches)
            while not feed.exhausted:
                loss = feed.step()
                opt.step()
                losses.append(loss.item())
                for vb in batches:   # validation-style calls in between must not disturb the batch staged ahead
                    model.loss(vb)
            with pytest.raises(StopIteration):
                feed.step()
        else:
            for b in batches:
                loss = model.loss_and_backward(b)
                opt.step()
                losses.append(loss.item())
        out.append((losses, {n: p.detach().clone() for n, p in model.named_parameters()}))
    assert out[0][0][0] == out[1][0][0]
    assert np.allclose(out[0][0], out[1][0], rtol=1e-4, atol=0)
    for n, p in out[0][1].items():
        assert (p - out[1][1][n]).abs().max().item() <= 3 * 7e-5, n   # three AdamW steps move an element by <= 3 lr
        if "pooler" in n or "key.bias" in n:   # never updated / true gradient identically zero
            continue
        assert cosine((p - build_ref[n]).cpu().numpy(), (out[1][1][n] - build_ref[n]).cpu().numpy()) > 0.9, n


def test_encoder_graph_replay_equals_eager(dev):
    """Encoder-only inference through a captured CUDA graph returns bit-identical hidden states to the
    eager launch sequence — for new inputs, padded batches and after the weights changed."""
    import plbert_b200 as P
    from plbert_b200.batch import synthetic_collate_batch
    cfgv = [188, 32, 128, 2, 256, 3, 128, 0]
    model = build_model(cfgv, False, dev, seed=8)
    g = torch.Generator().manual_seed(0)
    for trial in range(3):
        ids = torch.randint(1, 185, (5, 96), generator=g).to(dev)
        lens = torch.tensor([96, 96, 70, 33, 9]) if trial else torch.full((5,), 96)
        mask = (torch.arange(96)[None, :] < lens[:, None]).int().to(dev)
        eager = model.encode(ids, mask if trial else None)
        graphed = model.encode(ids, mask if trial else None, use_graph=True)
        assert graphed.shape == (5, 96, 128) and graphed.dtype == torch.bfloat16
        for b, n in enumerate(lens.tolist()):
            assert torch.equal(eager[b, :n], graphed[b, :n]), (trial, b)
    assert len(model._graphs) == 1
    # weights move (one optimizer step): the replay must see them
    opt = P.FusedAdamW(model.parameters(), lr=1e-2)
    model.loss_and_backward(synthetic_collate_batch(4, 128, seed=5, word_pred_prob=0.3))
    opt.step()
    after = model.encode(ids, mask, use_graph=True)
    assert not torch.equal(after, graphed)
    assert torch.equal(after[0], model.encode(ids, mask)[0])


def test_reference_style_forward_backward(dev):
    """Unchanged reference call pattern (R/train.py:381-390, 355-357): full logits -> python loss loop -> backward."""
    import plbert_b200 as P
    from plbert_b200.batch import synthetic_collate_batch
    cfgv = [188, 32, 128, 2, 256, 3, 64, 0]
    model = build_model(cfgv, False, dev, seed=5)
    labels, masked, lengths, midx = synthetic_collate_batch(4, 64, seed=2, variable_length=True, min_length=12,
                                                            word_pred_prob=0.3)
    params = {k: v.detach().cpu().double().numpy() for k, v in model.state_dict().items()}
    total, _, grads, hidden = O.loss_and_grads(params, oracle_cfg(cfgv), masked.numpy(), labels.numpy(), lengths, midx)
    from plbert_b200.train import length_to_mask
    text_mask = length_to_mask(torch.Tensor(lengths)).to(dev)
    pred = model(masked.to(dev), attention_mask=(~text_mask).int())
    assert pred.shape == (4, masked.shape[1], 188) and pred.dtype == torch.float32
    ref_logits = O.head_logits(params, hidden, "phoneme_predictor")
    for b, n in enumerate(lengths):
        assert cosine(pred[b, :n].detach().cpu().numpy(), ref_logits[b, :n]) > 0.9995
    criterion = torch.nn.CrossEntropyLoss()
    loss, count = 0, 0
    for p_, t_, n_, m_ in zip(pred, labels.to(dev), lengths, midx):  # the loop of R/train.py:123-129
        if len(m_) > 0:
            loss = loss + criterion(p_[:n_][m_], t_[:n_][m_])
            count += 1
    loss = loss / count
    opt = P.FusedAdamW(model.parameters(), lr=7e-5)
    opt.zero_grad()
    loss.backward()
    assert abs(loss.item() - total) <= 2e-2 * abs(total)
    check_grads(model, grads, masked.shape[1])
    opt.step()
    # p.grad aliasing the gradient arena (the fused path publishes arena views; zero_grad(set_to_none=False) keeps them):
    # the autograd node must not add the arena onto itself — d(sum of 8 positions' logits)/d(bias) = 8
    model.loss_and_backward((labels, masked, lengths, midx))
    assert model.phoneme_predictor.bias.grad.data_ptr() == model._engine.grad_views["phoneme_predictor.bias"].data_ptr()
    opt.zero_grad(set_to_none=False)
    pred = model(masked.to(dev), attention_mask=(~text_mask).int())
    pred[0, :8].sum().backward()
    g = model.phoneme_predictor.bias.grad
    assert torch.allclose(g, torch.full_like(g, 8.0), rtol=1e-2), g[:4]
    opt.step()


def test_styletts_style_finetuning_with_a_stock_torch_optimizer(dev):
    """R/README.md:46-117 (StyleTTS second-stage training): ``bert`` is finetuned through ``bert(texts, attention_mask=...)
    .last_hidden_state`` -> ``bert_encoder = nn.Linear(768, 512)`` with a STOCK ``torch.optim.AdamW`` whose param group is edited
    in place (betas (0.9, 0.99), lr 1e-5, weight_decay 0.01).  The fused encoder must follow the weights such an optimizer writes
    through torch (the bf16 shadow is refreshed when the parameters' version counters move) — compared with the stock HF
    encoder taking the same steps on the same data."""
    import plbert_b200 as P
    from transformers import AlbertConfig, AlbertModel
    cfg = AlbertConfig(vocab_size=188, embedding_size=32, hidden_size=128, num_attention_heads=2, intermediate_size=256,
                       num_hidden_layers=3, max_position_embeddings=128)
    torch.manual_seed(4)
    stock = AlbertModel(cfg)
    sd = {k: v.clone() for k, v in stock.state_dict().items()}
    proj0 = torch.nn.Linear(128, 64)
    g = torch.Generator().manual_seed(9)
    ids = torch.randint(1, 185, (4, 96), generator=g).to(dev)
    lens = torch.tensor([96, 70, 41, 12])
    text_mask = (torch.arange(96)[None, :] >= lens[:, None]).to(dev)                  # StyleTTS: True = padding
    target = torch.randn(4, 96, 64, generator=g).to(dev)
    outs = {}
    for kind in ("stock", "fused"):
        bert = AlbertModel(cfg) if kind == "stock" else P.fuse_albert(AlbertModel(cfg))
        bert.load_state_dict(sd, strict=True)
        bert = bert.to(dev).train()
        proj = torch.nn.Linear(128, 64)
        proj.load_state_dict(proj0.state_dict())
        proj = proj.to(dev)
        params = [p for n, p in bert.named_parameters() if "pooler" not in n]
        opt = torch.optim.AdamW(params, lr=1e-4)
        for grp in opt.param_groups:                                                   # R/README.md:79-86
            grp["betas"], grp["lr"], grp["weight_decay"] = (0.9, 0.99), 2e-3, 0.01     # large lr: three steps must move the output
        opt_p = torch.optim.AdamW(proj.parameters(), lr=1e-3)
        losses = []
        for step in range(3):
            hid = bert(ids, attention_mask=(~text_mask).int()).last_hidden_state       # R/README.md:91
            d_en = proj(hid.float())
            loss = ((d_en - target) ** 2 * (~text_mask)[:, :, None]).mean()
            opt.zero_grad()
            opt_p.zero_grad()
            loss.backward()
            opt_p.step()                                                               # R/README.md:116-117
            opt.step()
            losses.append(loss.item())
        with torch.no_grad():
            final = bert(ids, attention_mask=(~text_mask).int()).last_hidden_state.float()
        outs[kind] = (losses, final[~text_mask].cpu().numpy(), {n: p.detach().float().cpu().numpy() for n, p in bert.named_parameters()})
    for a, b in zip(outs["stock"][0], outs["fused"][0]):
        assert abs(a - b) <= 2e-2 * abs(a), (outs["stock"][0], outs["fused"][0])
    assert outs["stock"][0][2] < outs["stock"][0][0]                                   # the steps did train
    before = AlbertModel(cfg)
    before.load_state_dict(sd)
    with torch.no_grad():
        h0 = before.to(dev)(ids, attention_mask=(~text_mask).int()).last_hidden_state[~text_mask].cpu().numpy()
    assert cosine(outs["stock"][1], h0) < 0.99                                         # ... and moved the encoder output
    assert cosine(outs["fused"][1], outs["stock"][1]) > 0.999                          # the fused encoder followed the optimizer
    w = "encoder.albert_layer_groups.0.albert_layers.0.ffn.weight"
    assert cosine(outs["fused"][2][w] - sd[w].numpy(), outs["stock"][2][w] - sd[w].numpy()) > 0.9


def test_empty_batch_and_checkpoint_roundtrip(dev, tmp_path):
    import plbert_b200 as P
    from plbert_b200 import train as T
    from plbert_b200.batch import synthetic_collate_batch
    cfgv = [188, 32, 128, 2, 256, 2, 64, 50]
    model = build_model(cfgv, True, dev, seed=7)
    opt = P.FusedAdamW(model.parameters(), lr=7e-5)
    token_ids, labels, masked, lengths, midx = synthetic_collate_batch(3, 64, seed=4, with_token_ids=True, num_tokens=50)
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    loss = model.loss_and_backward((token_ids, labels, masked, lengths, [[] for _ in lengths]))
    assert loss.item() == 0.0 and all(p.grad is None for p in model.parameters())  # R/train.py:129
    opt.step()
    assert all(torch.equal(p.detach(), before[n]) for n, p in model.named_parameters())
    for _ in range(2):
        model.loss_and_backward((token_ids, labels, masked, lengths, midx))
        opt.step()
    path = T.save_checkpoint(model, opt, 2, str(tmp_path), 1)
    assert path.endswith("step_2.t7") and T.find_latest_checkpoint(str(tmp_path)) == (True, 2)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert set(ck) == {"net", "step", "epoch", "optimizer"} and all(k.startswith("module.") for k in ck["net"])
    # the reference's own loaders strip the prefix and load into the stock classes (R/convert_to_hf.py:40-43)
    from transformers import AlbertConfig, AlbertModel
    stock = AlbertModel(AlbertConfig(vocab_size=188, embedding_size=32, hidden_size=128, num_attention_heads=2,
                                     intermediate_size=256, num_hidden_layers=2, max_position_embeddings=64))
    enc_sd = {k[len("module.encoder."):]: v for k, v in ck["net"].items() if k.startswith("module.encoder.")}
    stock.load_state_dict(enc_sd, strict=True)  # StyleTTS-style strict load (R/README.md:57-66)
    # resume into a fresh model: identical next step
    model2 = build_model(cfgv, True, dev, seed=99)
    opt2 = P.FusedAdamW(model2.parameters(), lr=7e-5)
    model2.loss_and_backward((token_ids, labels, masked, lengths, midx))  # creates the arenas
    T.load_checkpoint(model2, opt2, path, device=dev)
    l1 = model.loss_and_backward((token_ids, labels, masked, lengths, midx)); opt.step()
    l2 = model2.loss_and_backward((token_ids, labels, masked, lengths, midx)); opt2.step()
    assert abs(l1.item() - l2.item()) < 1e-5 * abs(l1.item())
    for (n, a), (_, b) in zip(model.named_parameters(), model2.named_parameters()):
        assert torch.allclose(a, b, rtol=0, atol=1e-6), n


def test_cpu_call_fails_loudly():
    import plbert_b200 as P
    from transformers import AlbertConfig, AlbertModel
    enc = AlbertModel(AlbertConfig(vocab_size=188, embedding_size=32, hidden_size=128, num_attention_heads=2,
                                   intermediate_size=256, num_hidden_layers=2, max_position_embeddings=64))
    model = P.PhonemeOnlyModel(enc, 188, 128)
    with pytest.raises(P.PlbError):
        model(torch.zeros(2, 8, dtype=torch.long))


def test_fused_albert_encoder_call_shape_and_autograd(dev):
    """``encoder(ids, attention_mask=m).last_hidden_state`` (R/model.py:14-15; StyleTTS, R/README.md:91,107) runs the kernels:
    inside the drop-in models and stand-alone (``fuse_albert(AlbertModel(cfg))`` + the StyleTTS strict state-dict load), with
    and without gradients, against the stock HF implementation in fp32 on the same weights."""
    import copy
    import plbert_b200 as P
    from transformers import AlbertConfig, AlbertModel
    cfg = AlbertConfig(vocab_size=188, embedding_size=32, hidden_size=128, num_attention_heads=2, intermediate_size=256,
                       num_hidden_layers=3, max_position_embeddings=128)
    torch.manual_seed(21)
    stock = AlbertModel(cfg)
    with torch.no_grad():
        for n, p in stock.named_parameters():
            if n.endswith("bias") or "LayerNorm" in n or "layer_norm" in n:
                p.add_(0.05 * torch.randn_like(p))
    sd = {k: v.clone() for k, v in stock.state_dict().items()}
    stock = stock.to(dev).train()
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(1, 185, (4, 96), generator=g).to(dev)
    lens = torch.tensor([96, 80, 33, 7])
    mask = (torch.arange(96)[None, :] < lens[:, None]).int().to(dev)
    valid = mask.bool()
    ref = stock(ids, attention_mask=mask).last_hidden_state
    wgt = torch.randn(ref.shape, generator=g).to(dev) * valid[:, :, None]          # padded rows never reach a loss
    (ref * wgt).sum().backward()

    # stand-alone encoder, loaded the StyleTTS way (strict load of the encoder keys)
    bert = P.fuse_albert(AlbertModel(cfg))
    bert.load_state_dict(sd, strict=True)
    bert = bert.to(dev)
    with torch.no_grad():
        out = bert(ids, attention_mask=mask)
    assert out.last_hidden_state.dtype == torch.float32 and out.last_hidden_state.shape == (4, 96, 128) and out.pooler_output is None
    assert cosine(out.last_hidden_state[valid].cpu().numpy(), ref[valid].detach().cpu().numpy()) > 0.9995
    assert cosine(out[0][valid].cpu().numpy(), ref[valid].detach().cpu().numpy()) > 0.9995       # tuple-style access
    bert.plb_use_graph = True
    with torch.no_grad():
        graphed = bert(ids, attention_mask=mask).last_hidden_state
    assert torch.equal(graphed[valid], out.last_hidden_state[valid])
    # differentiable: finetuning through the encoder call
    hid = bert(ids, attention_mask=mask).last_hidden_state
    assert hid.requires_grad
    (hid * wgt).sum().backward()
    qb = stock.encoder.albert_layer_groups[0].albert_layers[0].attention.query.bias.grad.norm().item()
    for (n, p), (_, q) in zip(bert.named_parameters(), stock.named_parameters()):
        if "pooler" in n:
            assert p.grad is None
            continue
        if n.endswith("attention.key.bias"):
            assert p.grad.norm().item() <= 0.05 * qb
            continue
        assert cosine(p.grad.float().cpu().numpy(), q.grad.cpu().numpy()) >= 0.999, n
    # the same call on the encoder of a drop-in model uses the model's engine (one set of weights)
    model = P.PhonemeOnlyModel(AlbertModel(cfg), 188, 128)
    model.encoder.load_state_dict(sd, strict=True)
    model = model.to(dev)
    with torch.no_grad():
        h2 = model.encoder(ids, attention_mask=mask).last_hidden_state
    assert model.encoder._plb_owner is model and model._engine is not None
    assert torch.equal(h2[valid], out.last_hidden_state[valid])
    assert torch.equal(h2.to(torch.bfloat16)[valid], model.encode(ids, mask)[valid])
