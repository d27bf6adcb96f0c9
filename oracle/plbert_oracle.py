"""CPU oracle for the PL-BERT training hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module.  The product path (``plbert_b200``) never does and fails loudly
when its CUDA library is missing.

What it is: an independent numpy restatement (explicit forward AND hand-derived backward, any float
dtype) of the arithmetic the reference executes for one ``MultiTaskModel`` / ``PhonemeOnlyModel``
training step.  The reference keeps that arithmetic in third-party libraries that are not under
/root/reference (transformers 5.5.0 ``AlbertModel``, torch ``CrossEntropyLoss`` / ``AdamW``;
the reference README pins ``transformers<4.33.3``, no lockfile), so each function cites the call
site in the reference (R/ = /root/reference) and the library lines it restates
(HF/ = transformers/models/albert/modeling_albert.py, TORCH/ = torch/optim/adam.py).

Pinning: the reference ships no tests, golden vectors or fixtures, so the oracle is
pinned against outputs of the reference itself, executed in the build container by
``oracle/make_golden.py`` (R/model.py + HF AlbertModel + R/train.py loss + torch.optim.AdamW) and
committed under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays them.  Parity for the
token/word-head loss is UNPINNED by the reference (no such loss exists in R/):
it follows the phoneme-loss template of R/train.py:107-131.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

Params = Dict[str, np.ndarray]

_LAYER = "encoder.encoder.albert_layer_groups.0.albert_layers.0."
_EMB = "encoder.embeddings."


@dataclass
class OracleConfig:
    """Subset of AlbertConfig the path depends on (HF configuration_albert.py:56-75 defaults)."""

    vocab_size: int = 188
    embedding_size: int = 128
    hidden_size: int = 768
    num_attention_heads: int = 12
    intermediate_size: int = 2048
    num_hidden_layers: int = 12
    max_position_embeddings: int = 512
    type_vocab_size: int = 2
    layer_norm_eps: float = 1e-12

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads


# ----------------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------------
def param_shapes(cfg: OracleConfig, num_phonemes: int, num_tokens: Optional[int]) -> Dict[str, Tuple[int, ...]]:
    """state_dict names and shapes, in ``model.parameters()`` order."""
    E, H, I = cfg.embedding_size, cfg.hidden_size, cfg.intermediate_size
    s: Dict[str, Tuple[int, ...]] = {}
    s[_EMB + "word_embeddings.weight"] = (cfg.vocab_size, E)
    s[_EMB + "position_embeddings.weight"] = (cfg.max_position_embeddings, E)
    s[_EMB + "token_type_embeddings.weight"] = (cfg.type_vocab_size, E)
    s[_EMB + "LayerNorm.weight"] = (E,)
    s[_EMB + "LayerNorm.bias"] = (E,)
    s["encoder.encoder.embedding_hidden_mapping_in.weight"] = (H, E)
    s["encoder.encoder.embedding_hidden_mapping_in.bias"] = (H,)
    s[_LAYER + "full_layer_layer_norm.weight"] = (H,)
    s[_LAYER + "full_layer_layer_norm.bias"] = (H,)
    for n in ("query", "key", "value", "dense"):
        s[_LAYER + f"attention.{n}.weight"] = (H, H)
        s[_LAYER + f"attention.{n}.bias"] = (H,)
    s[_LAYER + "attention.LayerNorm.weight"] = (H,)
    s[_LAYER + "attention.LayerNorm.bias"] = (H,)
    s[_LAYER + "ffn.weight"] = (I, H)
    s[_LAYER + "ffn.bias"] = (I,)
    s[_LAYER + "ffn_output.weight"] = (H, I)
    s[_LAYER + "ffn_output.bias"] = (H,)
    s["encoder.pooler.weight"] = (H, H)
    s["encoder.pooler.bias"] = (H,)
    s["phoneme_predictor.weight"] = (num_phonemes, H)
    s["phoneme_predictor.bias"] = (num_phonemes,)
    if num_tokens is not None:
        s["token_predictor.weight"] = (num_tokens, H)
        s["token_predictor.bias"] = (num_tokens,)
    return s


def init_params(cfg: OracleConfig, num_phonemes: int, num_tokens: Optional[int], seed: int = 0,
                dtype=np.float64, ln_jitter: float = 0.0) -> Params:
    """Random parameters with the reference's init *statistics* (HF :304-318: normal(0, 0.02), LN w=1/b=0,
    word_embeddings[0] = 0).  Not bit-identical to torch's RNG — parity tests copy weights instead.
    ``ln_jitter`` > 0 draws the LayerNorm weights from normal(1, ln_jitter) so that their gradient paths are exercised.
    numpy's ``Generator.normal`` stream is stable across numpy versions, so (cfg, seed) names one parameter set that the
    fixture generator (which feeds it to the reference) and the tests can both rebuild without storing 10^7 numbers."""
    rng = np.random.default_rng(seed)
    p: Params = {}
    for name, shape in param_shapes(cfg, num_phonemes, num_tokens).items():
        if "LayerNorm.weight" in name or "layer_norm.weight" in name:
            v = np.ones(shape) if ln_jitter == 0.0 else rng.normal(1.0, ln_jitter, shape)
        elif name.endswith("bias"):
            v = rng.normal(0.0, 0.02, shape)  # non-zero biases exercise every bias-gradient path
        else:
            v = rng.normal(0.0, 0.02, shape)
        p[name] = v.astype(dtype)
    p[_EMB + "word_embeddings.weight"][0] = 0
    return p


# ----------------------------------------------------------------------------------------------
# a1 / a8 integer work (bit-exact)
# ----------------------------------------------------------------------------------------------
def length_to_mask(lengths: Sequence[int]) -> np.ndarray:
    """R/train.py:34-44 — bool [B, max_len], True = padding: (position + 1) > length."""
    lengths = np.asarray(lengths, dtype=np.int64)
    max_len = int(lengths.max())
    positions = np.arange(max_len, dtype=np.int64)[None, :]
    return (positions + 1) > lengths[:, None]


def attention_mask_from_lengths(lengths: Sequence[int]) -> np.ndarray:
    """R/train.py:386 — ``(~text_mask).int()``: int32 [B, max_len], 1 = valid token."""
    return (~length_to_mask(lengths)).astype(np.int32)


def flatten_masked_indices(input_lengths: Sequence[int], masked_indices: Sequence[Sequence[int]],
                           seq_stride: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Row selection of R/train.py:123-129 as flat arrays.

    For sample b with n_b = len(masked_indices[b]) > 0 the reference takes rows
    ``pred[b][:len_b][masked_indices[b]]`` and averages their CE; samples with n_b == 0 are skipped and
    the result is divided by ``count`` = number of non-empty samples.  Returned: flat row index
    ``b * seq_stride + i`` (int64, reference order), per-row weight ``1 / (count * n_b)`` (float64) and the
    owning sample of each row.  Python negative indices are resolved against ``len_b`` as numpy/torch would.
    """
    rows: List[int] = []
    weights: List[float] = []
    owner: List[int] = []
    count = sum(1 for m in masked_indices if len(m) > 0)
    for b, (length, idx) in enumerate(zip(input_lengths, masked_indices)):
        n_b = len(idx)
        if n_b == 0:
            continue
        for i in idx:
            i = int(i)
            if i < 0:
                i += int(length)
            if not (0 <= i < int(length)):
                raise IndexError(f"masked index {i} out of range for sample {b} of length {length}")
            rows.append(b * seq_stride + i)
            weights.append(1.0 / (count * n_b))
            owner.append(b)
    return np.asarray(rows, dtype=np.int64), np.asarray(weights, dtype=np.float64), np.asarray(owner, dtype=np.int64)


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------
def layer_norm_fwd(x: np.ndarray, w: np.ndarray, b: np.ndarray, eps: float):
    """torch.nn.LayerNorm: biased variance, eps inside the sqrt."""
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + x.dtype.type(eps))
    xhat = (x - mu) * rstd
    return xhat * w + b, xhat, rstd


def layer_norm_bwd(dy: np.ndarray, xhat: np.ndarray, rstd: np.ndarray, w: np.ndarray):
    """dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * w."""
    g = dy * w
    dx = rstd * (g - g.mean(axis=-1, keepdims=True) - xhat * (g * xhat).mean(axis=-1, keepdims=True))
    red = tuple(range(dy.ndim - 1))
    return dx, (dy * xhat).sum(axis=red), dy.sum(axis=red)


_GELU_C = 0.7978845608028654  # sqrt(2 / pi)


def gelu_new(u: np.ndarray) -> np.ndarray:
    """HF activations.py:59-66 NewGELUActivation (hidden_act='gelu_new', HF configuration_albert.py:64)."""
    return 0.5 * u * (1.0 + np.tanh(_GELU_C * (u + 0.044715 * u ** 3)))


def gelu_new_grad(u: np.ndarray) -> np.ndarray:
    t = np.tanh(_GELU_C * (u + 0.044715 * u ** 3))
    return 0.5 * (1.0 + t) + 0.5 * u * (1.0 - t * t) * _GELU_C * (1.0 + 0.134145 * u * u)


def _split_heads(x: np.ndarray, nh: int) -> np.ndarray:  # [B,L,H] -> [B,nh,L,d]
    B, L, H = x.shape
    return x.reshape(B, L, nh, H // nh).transpose(0, 2, 1, 3)


def _merge_heads(x: np.ndarray) -> np.ndarray:  # [B,nh,L,d] -> [B,L,H]
    B, nh, L, d = x.shape
    return x.transpose(0, 2, 1, 3).reshape(B, L, nh * d)


def _lin(p: Params, name: str, dtype):
    return p[name + ".weight"].astype(dtype, copy=False), p[name + ".bias"].astype(dtype, copy=False)


# ----------------------------------------------------------------------------------------------
# encoder forward / backward
# ----------------------------------------------------------------------------------------------
def encoder_forward(params: Params, input_ids: np.ndarray, attention_mask: Optional[np.ndarray], cfg: OracleConfig,
                    dtype=np.float64, keep_cache: bool = True, dropout: Optional[dict] = None):
    """``AlbertModel.forward`` → last_hidden_state (HF :387-423), restated per SURVEY.md §3.2.

    embeddings (HF :71-110) → 128→H projection (HF :276) → ``num_hidden_layers`` applications of the ONE shared
    layer (HF :278-286): attention with key-padding mask (HF :170-204, mask semantics HF masking_utils.py:495-520:
    padded KEYS get -inf, padded queries are still computed), post-LN residual blocks (HF :200-202, :235),
    gelu_new FFN (HF :239-241).  The pooler (HF :418) is skipped: PL-BERT never reads it (R/model.py:14-16).

    ``dropout`` (training with hidden_dropout_prob / attention_probs_dropout_prob > 0; the reference config has 0): explicit
    multiplier arrays (0 or 1/(1-p)) for the three nn.Dropout sites — ``emb`` [B,L,E] after the embedding LayerNorm (HF :109),
    ``attn`` list per layer application of [B,nh,L,L] on the attention probabilities (HF :134), ``out`` list of [B,L,H] on the
    attention output projection before the residual add (HF :201).  Tests build them with oracle/dropout_rng.py.
    """
    ids = np.asarray(input_ids)
    B, L = ids.shape
    nh, d = cfg.num_attention_heads, cfg.head_dim
    eps = cfg.layer_norm_eps
    g = lambda n: params[n].astype(dtype, copy=False)

    word, pos, typ = g(_EMB + "word_embeddings.weight"), g(_EMB + "position_embeddings.weight"), g(
        _EMB + "token_type_embeddings.weight")
    e_sum = (word[ids] + typ[0][None, None, :]) + pos[:L][None, :, :]
    e, e_hat, e_rstd = layer_norm_fwd(e_sum, g(_EMB + "LayerNorm.weight"), g(_EMB + "LayerNorm.bias"), eps)
    if dropout is not None and dropout.get("emb") is not None:
        e = e * dropout["emb"]
    Wm, bm = _lin(params, "encoder.encoder.embedding_hidden_mapping_in", dtype)
    x = e @ Wm.T + bm

    if attention_mask is None:
        key_bias = np.zeros((B, 1, 1, L), dtype=dtype)
    else:
        am = np.asarray(attention_mask)
        key_bias = np.where(am[:, None, None, :] != 0, 0.0, -np.inf).astype(dtype)

    Wq, bq = _lin(params, _LAYER + "attention.query", dtype)
    Wk, bk = _lin(params, _LAYER + "attention.key", dtype)
    Wv, bv = _lin(params, _LAYER + "attention.value", dtype)
    Wo, bo = _lin(params, _LAYER + "attention.dense", dtype)
    W1, b1 = _lin(params, _LAYER + "ffn", dtype)
    W2, b2 = _lin(params, _LAYER + "ffn_output", dtype)
    ln1w, ln1b = g(_LAYER + "attention.LayerNorm.weight"), g(_LAYER + "attention.LayerNorm.bias")
    ln2w, ln2b = g(_LAYER + "full_layer_layer_norm.weight"), g(_LAYER + "full_layer_layer_norm.bias")
    scale = dtype(1.0 / np.sqrt(d)) if isinstance(dtype, type) else 1.0 / np.sqrt(d)

    layers = []
    for li in range(cfg.num_hidden_layers):
        m_attn = dropout["attn"][li] if dropout is not None and dropout.get("attn") is not None else None
        m_out = dropout["out"][li] if dropout is not None and dropout.get("out") is not None else None
        q = _split_heads(x @ Wq.T + bq, nh)
        k = _split_heads(x @ Wk.T + bk, nh)
        v = _split_heads(x @ Wv.T + bv, nh)
        s = (q @ k.transpose(0, 1, 3, 2)) * scale + key_bias
        s_max = s.max(axis=-1, keepdims=True)
        pexp = np.exp(s - s_max)
        prob = pexp / pexp.sum(axis=-1, keepdims=True)
        ctx = _merge_heads((prob if m_attn is None else prob * m_attn) @ v)
        proj = ctx @ Wo.T + bo
        z1 = x + (proj if m_out is None else proj * m_out)
        a, a_hat, a_rstd = layer_norm_fwd(z1, ln1w, ln1b, eps)
        u = a @ W1.T + b1
        gact = gelu_new(u)
        z2 = a + (gact @ W2.T + b2)
        y, y_hat, y_rstd = layer_norm_fwd(z2, ln2w, ln2b, eps)
        if keep_cache:
            layers.append(dict(x=x, q=q, k=k, v=v, prob=prob, ctx=ctx, a=a, a_hat=a_hat, a_rstd=a_rstd, u=u, g=gact,
                               y_hat=y_hat, y_rstd=y_rstd, m_attn=m_attn, m_out=m_out))
        x = y
    cache = dict(ids=ids, e=e, e_hat=e_hat, e_rstd=e_rstd, layers=layers, L=L,
                 m_emb=dropout.get("emb") if dropout is not None else None) if keep_cache else None
    return x, cache


def encoder_backward(params: Params, cache: dict, d_last_hidden: np.ndarray, cfg: OracleConfig, dtype=np.float64) -> Params:
    """Hand-derived backward of :func:`encoder_forward` (formulas: SURVEY.md appendix A).  Gradients of the
    shared layer are the SUM over its ``num_hidden_layers`` applications (autograd AccumulateGrad in the
    reference, R/train.py:356).  Pooler parameters receive no gradient and are absent from the result."""
    nh, d = cfg.num_attention_heads, cfg.head_dim
    g = lambda n: params[n].astype(dtype, copy=False)
    Wq, Wk, Wv, Wo = (g(_LAYER + f"attention.{n}.weight") for n in ("query", "key", "value", "dense"))
    W1, W2 = g(_LAYER + "ffn.weight"), g(_LAYER + "ffn_output.weight")
    ln1w, ln2w = g(_LAYER + "attention.LayerNorm.weight"), g(_LAYER + "full_layer_layer_norm.weight")
    scale = 1.0 / np.sqrt(d)
    H = cfg.hidden_size

    grads: Params = {}

    def acc(name, val):
        grads[name] = grads[name] + val if name in grads else val

    def flat(t):
        return t.reshape(-1, t.shape[-1])

    dy = d_last_hidden.astype(dtype, copy=False)
    for c in reversed(cache["layers"]):
        # y = LN2(a + g W2^T + b2)
        dz2, dw, db = layer_norm_bwd(dy, c["y_hat"], c["y_rstd"], ln2w)
        acc(_LAYER + "full_layer_layer_norm.weight", dw)
        acc(_LAYER + "full_layer_layer_norm.bias", db)
        acc(_LAYER + "ffn_output.weight", flat(dz2).T @ flat(c["g"]))
        acc(_LAYER + "ffn_output.bias", flat(dz2).sum(0))
        du = (dz2 @ W2) * gelu_new_grad(c["u"])
        acc(_LAYER + "ffn.weight", flat(du).T @ flat(c["a"]))
        acc(_LAYER + "ffn.bias", flat(du).sum(0))
        da = dz2 + du @ W1
        # a = LN1(x + ctx Wo^T + bo)
        dz1, dw, db = layer_norm_bwd(da, c["a_hat"], c["a_rstd"], ln1w)
        acc(_LAYER + "attention.LayerNorm.weight", dw)
        acc(_LAYER + "attention.LayerNorm.bias", db)
        dproj = dz1 if c["m_out"] is None else dz1 * c["m_out"]      # z1 = x + dropout(proj)
        acc(_LAYER + "attention.dense.weight", flat(dproj).T @ flat(c["ctx"]))
        acc(_LAYER + "attention.dense.bias", flat(dproj).sum(0))
        dctx = _split_heads(dproj @ Wo, nh)
        # softmax attention (ctx = dropout(prob) v)
        prob, q, k, v = c["prob"], c["q"], c["k"], c["v"]
        prob_d = prob if c["m_attn"] is None else prob * c["m_attn"]
        dv = prob_d.transpose(0, 1, 3, 2) @ dctx
        dp = dctx @ v.transpose(0, 1, 3, 2)
        if c["m_attn"] is not None:
            dp = dp * c["m_attn"]
        ds = prob * (dp - (dp * prob).sum(axis=-1, keepdims=True))
        dq = (ds @ k) * scale
        dk = (ds.transpose(0, 1, 3, 2) @ q) * scale
        dq, dk, dv = _merge_heads(dq), _merge_heads(dk), _merge_heads(dv)
        x = c["x"]
        dx = dz1.copy()
        for n, dt, W in (("query", dq, Wq), ("key", dk, Wk), ("value", dv, Wv)):
            acc(_LAYER + f"attention.{n}.weight", flat(dt).T @ flat(x))
            acc(_LAYER + f"attention.{n}.bias", flat(dt).sum(0))
            dx = dx + dt @ W
        dy = dx
    # x0 = e Wm^T + bm ; e = LN(word[ids] + type[0] + pos[:L])
    Wm = g("encoder.encoder.embedding_hidden_mapping_in.weight")
    grads["encoder.encoder.embedding_hidden_mapping_in.weight"] = flat(dy).T @ flat(cache["e"])
    grads["encoder.encoder.embedding_hidden_mapping_in.bias"] = flat(dy).sum(0)
    de = dy @ Wm
    if cache.get("m_emb") is not None:
        de = de * cache["m_emb"]
    dsum, dw, db = layer_norm_bwd(de, cache["e_hat"], cache["e_rstd"], g(_EMB + "LayerNorm.weight"))
    grads[_EMB + "LayerNorm.weight"] = dw
    grads[_EMB + "LayerNorm.bias"] = db
    ids, L = cache["ids"], cache["L"]
    dword = np.zeros_like(g(_EMB + "word_embeddings.weight"))
    np.add.at(dword, ids.reshape(-1), flat(dsum))
    dword[0] = 0  # padding_idx=0 (HF :56): the pad row never receives gradient
    grads[_EMB + "word_embeddings.weight"] = dword
    dpos = np.zeros_like(g(_EMB + "position_embeddings.weight"))
    dpos[:L] = dsum.sum(axis=0)
    grads[_EMB + "position_embeddings.weight"] = dpos
    dtyp = np.zeros_like(g(_EMB + "token_type_embeddings.weight"))
    dtyp[0] = flat(dsum).sum(0)
    grads[_EMB + "token_type_embeddings.weight"] = dtyp
    assert dy.shape[-1] == H
    return grads


# ----------------------------------------------------------------------------------------------
# heads + loss
# ----------------------------------------------------------------------------------------------
def head_logits(params: Params, hidden: np.ndarray, head: str, dtype=np.float64) -> np.ndarray:
    """R/model.py:15-16,28 — ``nn.Linear(hidden_size, vocab)`` over every position."""
    W, b = _lin(params, head, dtype)
    return hidden @ W.T + b


def masked_ce(logits_rows: np.ndarray, labels_rows: np.ndarray, weights: np.ndarray):
    """Weighted sum of row-wise ``-log softmax(z)[y]`` (nn.CrossEntropyLoss, R/train.py:215) and its gradient
    ``w (softmax(z) - onehot(y))``."""
    z = logits_rows
    zmax = z.max(axis=-1, keepdims=True)
    ez = np.exp(z - zmax)
    se = ez.sum(axis=-1, keepdims=True)
    lse = (np.log(se) + zmax)[:, 0]
    n = z.shape[0]
    picked = z[np.arange(n), labels_rows]
    w = weights.astype(z.dtype)
    loss = (w * (lse - picked)).sum()
    dz = ez / se
    dz[np.arange(n), labels_rows] -= 1.0
    dz *= w[:, None]
    return loss, dz


def calculate_loss(logits: np.ndarray, labels: np.ndarray, input_lengths: Sequence[int],
                   masked_indices: Sequence[Sequence[int]]):
    """R/train.py:107-131 ``calculate_phoneme_loss``: mean over non-empty samples of the per-sample mean CE over
    that sample's masked rows; 0 when every sample is empty (R/train.py:129).  Returns (loss, dlogits)."""
    B, L, V = logits.shape
    rows, w, _ = flatten_masked_indices(input_lengths, masked_indices, L)
    dlogits = np.zeros_like(logits)
    if rows.size == 0:
        return logits.dtype.type(0.0), dlogits
    loss, dz = masked_ce(logits.reshape(B * L, V)[rows], np.asarray(labels).reshape(-1)[rows], w)
    np.add.at(dlogits.reshape(B * L, V), rows, dz)
    return loss, dlogits


# ----------------------------------------------------------------------------------------------
# whole step
# ----------------------------------------------------------------------------------------------
def loss_and_grads(params: Params, cfg: OracleConfig, masked_phonemes: np.ndarray, phoneme_labels: np.ndarray,
                   input_lengths: Sequence[int], masked_indices: Sequence[Sequence[int]],
                   token_labels: Optional[np.ndarray] = None, dtype=np.float64, token_rows: str = "masked",
                   dropout: Optional[dict] = None):
    """One forward+backward of the reference step: R/train.py:381-390 ``process_batch`` followed by
    ``loss.backward()`` (R/train.py:356).  With ``token_labels`` the same loss template is applied to
    ``token_predictor`` and the two losses are summed (unpinned half, see module docstring); ``token_rows="all_valid"``
    runs the token loss over every valid position of a sample instead of its masked ones (upstream's word loss)."""
    lengths = [int(x) for x in input_lengths]
    attn = attention_mask_from_lengths(lengths)
    assert attn.shape[1] == np.asarray(masked_phonemes).shape[1], "batch must be padded to max(input_lengths)"
    hidden, cache = encoder_forward(params, masked_phonemes, attn, cfg, dtype, dropout=dropout)
    B, L, H = hidden.shape
    grads: Params = {}
    losses = {}
    dhidden = np.zeros_like(hidden)
    heads = [("phoneme_predictor", phoneme_labels)]
    if token_labels is not None:
        heads.append(("token_predictor", token_labels))
    for head, labels in heads:
        logits = head_logits(params, hidden, head, dtype)
        rows = masked_indices
        if head == "token_predictor" and token_rows == "all_valid":
            rows = [list(range(n)) for n in lengths]
        loss, dlogits = calculate_loss(logits, labels, lengths, rows)
        W = params[head + ".weight"].astype(dtype, copy=False)
        grads[head + ".weight"] = dlogits.reshape(B * L, -1).T @ hidden.reshape(B * L, H)
        grads[head + ".bias"] = dlogits.reshape(B * L, -1).sum(0)
        dhidden = dhidden + dlogits @ W
        losses[head] = loss
    grads.update(encoder_backward(params, cache, dhidden, cfg, dtype))
    total = sum(losses.values())
    return total, losses, grads, hidden


def adamw_step(params: Params, grads: Params, state: Dict[str, dict], lr: float = 7e-5,
               betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01) -> None:
    """torch.optim.AdamW as the reference constructs it (R/train.py:272: one param group, defaults) — update rule
    TORCH/optim/adam.py:417-419,457,476,531-547.  In place.  Parameters without a gradient (the pooler) are skipped
    and get no state, exactly like torch."""
    b1, b2 = betas
    for name, gval in grads.items():
        p = params[name]
        st = state.setdefault(name, dict(step=0, exp_avg=np.zeros_like(p), exp_avg_sq=np.zeros_like(p)))
        st["step"] += 1
        t = st["step"]
        gval = gval.astype(p.dtype, copy=False)
        p *= p.dtype.type(1.0 - lr * weight_decay)
        st["exp_avg"] = st["exp_avg"] * p.dtype.type(b1) + gval * p.dtype.type(1.0 - b1)
        st["exp_avg_sq"] = st["exp_avg_sq"] * p.dtype.type(b2) + (gval * gval) * p.dtype.type(1.0 - b2)
        bc1 = 1.0 - b1 ** t
        bc2 = 1.0 - b2 ** t
        step_size = lr / bc1
        denom = np.sqrt(st["exp_avg_sq"]) / p.dtype.type(np.sqrt(bc2)) + p.dtype.type(eps)
        p -= p.dtype.type(step_size) * (st["exp_avg"] / denom)
