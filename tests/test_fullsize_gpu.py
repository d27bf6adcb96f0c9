"""Parity of the CUDA step AT THE BENCHMARKED ARCHITECTURE (BASELINE.json configs[1]: hidden 768, 12 heads, 12 applications
of the shared layer, embedding 128, FFN 2048, 512 positions, Vp 188): the only size at which the 12-deep in-place fp32
accumulation of the shared-layer gradients, four resident K/V boxes per (sequence, head) and the 2-CTA 256x256 GEMM tiles
act together.  Checked against

  * golden case D — outputs of the reference itself (R/model.py + HF AlbertModel + R/train.py loss, float64) on a 3 x 512
    batch with one padded and one empty-mask sample (tests/golden/full_size_step.npz, oracle/make_golden.py), and
  * the float64 oracle on the same and on fresh inputs (the oracle is pinned to case D on CPU: test_oracle_golden.py).

Bars (north star, bf16): loss rel <= 2e-2, per-tensor gradient cosine >= 0.999, index work bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import plbert_oracle as O  # noqa: E402  (the checker; never on the product path)
from oracle.make_golden import subsample_index  # noqa: E402
from test_oracle_golden import load_full_size_case  # noqa: E402
from test_step_gpu import build_model, check_grads, cosine  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def case_d(golden_dir):
    z, cfg, params, midx = load_full_size_case(golden_dir)
    total, losses, grads, hidden = O.loss_and_grads(params, cfg, z["masked"], z["labels"], z["lengths"], midx,
                                                    token_labels=z["token_ids"])
    return dict(z=z, cfg=cfg, params=params, midx=midx, total=total, losses=losses, grads=grads, hidden=hidden)


def _batch(z, midx):
    return (torch.from_numpy(z["token_ids"]), torch.from_numpy(z["labels"]), torch.from_numpy(z["masked"]),
            z["lengths"].tolist(), midx)


@pytest.mark.parametrize("packed", [False, True])
def test_full_size_step_matches_reference_golden_and_oracle(dev, case_d, packed):
    z, midx = case_d["z"], case_d["midx"]
    model = build_model(z["config"], True, dev, case_d["params"])
    model.pack_sequences = packed
    loss = model.loss_and_backward(_batch(z, midx))
    assert model.last_batch.packed == packed and model.last_batch.T == (int(z["lengths"].sum()) if packed else 3 * 512)
    ref_loss = float(z["loss_phoneme"] + z["loss_token"])                 # the reference's own number
    assert abs(loss.item() - ref_loss) <= 2e-2 * abs(ref_loss), (loss.item(), ref_loss)
    worst = check_grads(model, case_d["grads"], 512)                       # every element, vs the oracle
    # ... and directly against the reference's gradients on the elements the fixture keeps
    got = {n: p.grad.detach().float().cpu().numpy().reshape(-1) for n, p in model.named_parameters() if p.grad is not None}
    worst_sub = 1.0
    for n, g in got.items():
        if n.endswith("attention.key.bias"):
            continue
        sub = g[subsample_index(n, g.size)]
        c = cosine(sub, z["gradsub/" + n])
        worst_sub = min(worst_sub, c)
        assert c >= 0.999, (n, c)
        assert 0.97 < np.linalg.norm(g) / float(z["gradnorm/" + n]) < 1.03, n
    print(f"full size (packed={packed}): loss {loss.item():.5f} vs reference {ref_loss:.5f}; worst gradient cosine "
          f"{worst:.6f} (oracle, all elements) / {worst_sub:.6f} (reference, subsample)")
    # encoder-only forward on the same weights: last_hidden_state at valid positions
    lens = z["lengths"]
    mask = torch.from_numpy((np.arange(512)[None, :] < lens[:, None]).astype(np.int32)).to(dev)
    hid = model.encode(torch.from_numpy(z["masked"]).to(dev), mask).float().cpu().numpy()
    for b, n in enumerate(lens):
        assert cosine(hid[b, :n], case_d["hidden"][b, :n]) > 0.9995
    idx = subsample_index("last_hidden", hid.size)
    valid = (np.arange(512)[None, :] < lens[:, None])[:, :, None].repeat(768, 2).reshape(-1)[idx]
    assert cosine(hid.reshape(-1)[idx][valid], z["last_hidden_sub"][valid]) > 0.9995


def test_full_size_reference_style_two_head_forward_backward(dev, case_d):
    """``MultiTaskModel.forward`` (R/model.py:13-18) as the reference calls it: full [B, L, V] logits of BOTH heads with
    autograd, the python loss loop of R/train.py:123-129 on each head, ``loss.backward()``."""
    import plbert_b200 as P
    from plbert_b200.train import length_to_mask
    z, midx = case_d["z"], case_d["midx"]
    model = build_model(z["config"], True, dev, case_d["params"])
    tok, lab, msk, lengths, _ = _batch(z, midx)
    text_mask = length_to_mask(torch.Tensor(lengths)).to(dev)
    p_pred, t_pred = model(msk.to(dev), attention_mask=(~text_mask).int())
    assert p_pred.shape == (3, 512, 188) and t_pred.shape == (3, 512, 5003) and p_pred.dtype == t_pred.dtype == torch.float32
    for head, pred in (("phoneme_predictor", p_pred), ("token_predictor", t_pred)):
        ref = O.head_logits(case_d["params"], case_d["hidden"], head)
        for b, n in enumerate(lengths):
            assert cosine(pred[b, :n].detach().cpu().numpy(), ref[b, :n]) > 0.9995, head
    criterion = torch.nn.CrossEntropyLoss()
    total = 0
    for pred, target in ((p_pred, lab.to(dev)), (t_pred, tok.to(dev))):
        loss, count = 0, 0
        for p_, t_, n_, m_ in zip(pred, target, lengths, midx):
            if len(m_) > 0:
                loss = loss + criterion(p_[:n_][m_], t_[:n_][m_])
                count += 1
        total = total + loss / count
    opt = P.FusedAdamW(model.parameters(), lr=7e-5)
    opt.zero_grad()
    total.backward()
    assert abs(total.item() - case_d["total"]) <= 2e-2 * abs(case_d["total"])
    check_grads(model, case_d["grads"], 512)
    opt.step()


def test_full_size_64k_token_head(dev):
    """Same architecture with the benchmarked 64 000-word token head on a few hundred masked rows (B = 2, L = 512)."""
    from plbert_b200.batch import synthetic_collate_batch
    cfgv = [188, 128, 768, 12, 2048, 12, 512, 64000]
    cfg = O.OracleConfig(vocab_size=188)
    params = {k: v.astype(np.float32).astype(np.float64) for k, v in O.init_params(cfg, 188, 64000, seed=77, ln_jitter=0.05).items()}
    model = build_model(cfgv, True, dev, params)
    tok, lab, msk, lengths, midx = synthetic_collate_batch(2, 512, seed=5, with_token_ids=True, num_tokens=64000)
    assert lengths == [512, 512] and 60 < sum(len(m) for m in midx) <= 256
    total, losses, grads, _ = O.loss_and_grads(params, cfg, msk.numpy(), lab.numpy(), lengths, midx, token_labels=tok.numpy())
    loss = model.loss_and_backward((tok, lab, msk, lengths, midx))
    assert abs(loss.item() - total) <= 2e-2 * abs(total), (loss.item(), total, losses)
    worst = check_grads(model, grads, 512)
    print(f"full size, 64k head: loss {loss.item():.5f} vs {total:.5f} ({losses}); worst gradient cosine {worst:.6f}")
