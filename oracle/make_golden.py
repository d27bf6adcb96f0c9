"""Generate the golden fixtures under tests/golden/ by EXECUTING THE REFERENCE in this container.

TEST INFRASTRUCTURE.  Run from the repo root (CPU only, needs /root/reference, torch, transformers):

    python oracle/make_golden.py

What runs is the reference's own code, imported from /root/reference (never copied):
``model.MultiTaskModel`` / ``model.PhonemeOnlyModel`` (R/model.py:5-30) around the installed HF ``AlbertModel``,
``train.length_to_mask`` / ``train.process_batch`` / ``train.calculate_phoneme_loss`` (R/train.py:34-44,107-131,
381-390), ``dataloader.MaskedPhonemeDataset`` + ``Collater`` / ``PhonemeOnlyCollater`` (R/dataloader.py:19-142,
200-297) for the batches, and ``torch.optim.AdamW`` as constructed at R/train.py:272.  ``accelerate`` is not
installed here, so a stub module is inserted before ``import train``; nothing of
accelerate is on the arithmetic path.  Everything is evaluated in float64 so the fixtures pin the math, not
one machine's float32 rounding.
"""
from __future__ import annotations

import os
import random
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from transformers import AlbertConfig, AlbertModel  # noqa: F401  (must precede the accelerate stub)
    if "accelerate" not in sys.modules:
        acc = types.ModuleType("accelerate")
        acc.Accelerator = object
        acc.DistributedDataParallelKwargs = lambda **k: None
        sys.modules["accelerate"] = acc
    import train as ref_train
    import model as ref_model
    import dataloader as ref_data
    import char_indexer as ref_sym
    return ref_train, ref_model, ref_data, ref_sym


def synthetic_records(n, rng, words_lo, words_hi, symbols, num_tokens):
    """Records shaped like the reference dataset rows: {'phonemes': [word strings], 'token_ids': [ints]}."""
    alphabet = symbols[52:185]  # IPA + latin letters (never PAD / mask / separator / punctuation)
    recs = []
    for _ in range(n):
        nw = int(rng.integers(words_lo, words_hi + 1))
        words = ["".join(rng.choice(alphabet, size=int(rng.integers(2, 8)))) for _ in range(nw)]
        recs.append({"phonemes": words, "token_ids": [int(t) for t in rng.integers(5, num_tokens, size=nw)]})
    return recs


def ragged(list_of_lists):
    flat = np.asarray([i for l in list_of_lists for i in l], dtype=np.int64)
    offs = np.cumsum([0] + [len(l) for l in list_of_lists]).astype(np.int64)
    return flat, offs


def state_to_np(sd):
    return {k: v.detach().cpu().numpy().copy() for k, v in sd.items()}


FULL = dict(hidden_size=768, num_attention_heads=12, intermediate_size=2048, max_position_embeddings=512,
            num_hidden_layers=12, embedding_size=128)   # R/configs/config.yml:32-39 (embedding_size: HF default 128)
SUBSAMPLE = 8192   # gradient elements kept per tensor in the full-size fixture


def subsample_index(name: str, numel: int) -> np.ndarray:
    """Fixed pseudo-random element subset of a tensor (all of it when small), shared by the generator and the tests."""
    if numel <= SUBSAMPLE:
        return np.arange(numel)
    import zlib
    return np.sort(np.random.default_rng(zlib.crc32(name.encode())).choice(numel, SUBSAMPLE, replace=False))


def full_size_case():
    """Case D: the BENCHMARKED architecture (H 768, 12 heads, 12 applications of the shared layer, E 128, I 2048, 512
    positions, Vp 188) with a 5003-word token head, B = 3, L = 512 — one sample shorter than the others (padded), one
    sample without masked rows — evaluated by the reference in float64.  The 10^7 parameters are not stored: they are
    ``plbert_oracle.init_params(cfg, 188, 5003, seed=2024, ln_jitter=0.05)`` rounded to float32; of the gradients the
    fixture keeps every tensor's norm and a fixed 8192-element subsample, which pins the oracle at this size (and the
    oracle's full gradients then pin the CUDA path)."""
    ref_train, ref_model, ref_data, ref_sym = import_reference()
    from transformers import AlbertConfig, AlbertModel
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle import plbert_oracle as O
    symbols = ref_sym.symbols
    Vp, Vt, seed = len(symbols), 5003, 2024
    np.random.seed(31)
    random.seed(31)
    rng = np.random.default_rng(31)
    recs = synthetic_records(2, rng, 130, 150, symbols, Vt) + synthetic_records(1, rng, 60, 60, symbols, Vt)
    ds = ref_data.MaskedPhonemeDataset(recs, use_token_ids=True, word_separator=87, max_seq_length=512, word_pred_prob=0.15,
                                       phoneme_mask_prob=0.8, replace_prob=0.1)
    token_ids, labels, masked, lengths, midx = ref_data.Collater()([ds[i] for i in range(len(ds))])
    midx = [list(m) for m in midx]
    assert lengths[0] == 512 and lengths[-1] < 512 and all(len(m) > 0 for m in midx)
    midx[1] = []   # a sample without masked rows is skipped by the loss (R/train.py:124)
    cfg = O.OracleConfig(vocab_size=Vp, **FULL)
    params = {k: v.astype(np.float32) for k, v in O.init_params(cfg, Vp, Vt, seed=seed, ln_jitter=0.05).items()}
    net = ref_model.MultiTaskModel(AlbertModel(AlbertConfig(vocab_size=Vp, **FULL)), num_phonemes=Vp, num_tokens=Vt,
                                   hidden_size=FULL["hidden_size"])
    missing, unexpected = net.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()}, strict=False)
    assert not unexpected and all("pooler" in m or "position_ids" in m or "token_type_ids" in m for m in missing), missing
    net = net.double().train()
    criterion = torch.nn.CrossEntropyLoss()
    attn = (~ref_train.length_to_mask(torch.Tensor(lengths))).int()
    p_pred, t_pred = net(masked, attention_mask=attn)
    loss_p = ref_train.calculate_phoneme_loss(p_pred, labels, lengths, midx, criterion)
    loss_t = ref_train.calculate_phoneme_loss(t_pred, token_ids, lengths, midx, criterion)
    (loss_p + loss_t).backward()
    out = {"config": np.asarray([Vp, FULL["embedding_size"], FULL["hidden_size"], FULL["num_attention_heads"],
                                 FULL["intermediate_size"], FULL["num_hidden_layers"], FULL["max_position_embeddings"], Vt],
                                dtype=np.int64),
           "param_seed": np.int64(seed), "ln_jitter": np.float64(0.05),
           "token_ids": token_ids.numpy(), "labels": labels.numpy(), "masked": masked.numpy(),
           "lengths": np.asarray(lengths, dtype=np.int64), "loss_phoneme": np.float64(loss_p.item()),
           "loss_token": np.float64(loss_t.item())}
    out["midx_flat"], out["midx_offs"] = ragged(midx)
    hidden = net.encoder(masked, attention_mask=attn).last_hidden_state.detach().numpy().reshape(-1)
    out["last_hidden_sub"] = hidden[subsample_index("last_hidden", hidden.size)]
    for n, p in net.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.detach().numpy().reshape(-1)
        out[f"gradnorm/{n}"] = np.float64(np.linalg.norm(g))
        out[f"gradsub/{n}"] = g[subsample_index(n, g.size)]
    np.savez_compressed(os.path.join(OUT, "full_size_step.npz"), **out)
    print("case D: loss_p %.6f loss_t %.6f lengths %s masked rows %s" % (loss_p.item(), loss_t.item(), lengths,
                                                                         [len(m) for m in midx]))


def dataset_case():
    """Outputs of the reference's own ``MaskedPhonemeDataset`` + collaters (R/dataloader.py:19-142,200-223,276-297) under
    fixed ``np.random`` / ``random`` seeds, for the host-side mirror in plbert_b200/data.py.  Records are regenerated by the
    test from ``plbert_b200.data.synthetic_records`` (same seed), only the outputs are stored."""
    ref_train, ref_model, ref_data, ref_sym = import_reference()
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from plbert_b200.data import synthetic_records as make_records
    out = {"symbols": np.asarray(ref_sym.symbols)}
    for name, use_tok, max_len, words in (("short", True, 48, (3, 12)), ("long", False, 512, (100, 150))):
        recs = make_records(6, seed=5, num_tokens=997, min_words=words[0], max_words=words[1])
        ds = ref_data.MaskedPhonemeDataset(recs, word_pred_prob=0.3, phoneme_mask_prob=0.8, replace_prob=0.1, word_separator=87,
                                           max_seq_length=max_len, use_token_ids=use_tok)
        np.random.seed(99)
        random.seed(99)
        items = [ds[i] for i in range(len(ds))]
        batch = (ref_data.Collater() if use_tok else ref_data.PhonemeOnlyCollater())(items)
        *tensors, lengths, midx = batch
        for k, t in enumerate(tensors):
            out[f"{name}/tensor{k}"] = t.numpy()
        out[f"{name}/lengths"] = np.asarray(lengths, dtype=np.int64)
        out[f"{name}/midx_flat"], out[f"{name}/midx_offs"] = ragged([list(m) for m in midx])
    np.savez_compressed(os.path.join(OUT, "dataset_pipeline.npz"), **out)
    print("dataset case: lengths", out["short/lengths"].tolist(), out["long/lengths"].tolist())


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "dataset":
        os.makedirs(OUT, exist_ok=True)
        dataset_case()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "D":   # only the full-size case (a minute of float64 CPU work)
        os.makedirs(OUT, exist_ok=True)
        full_size_case()
        return
    ref_train, ref_model, ref_data, ref_sym = import_reference()
    from transformers import AlbertConfig, AlbertModel

    os.makedirs(OUT, exist_ok=True)
    symbols = ref_sym.symbols
    Vp = len(symbols)
    Vt = 97
    dataset_params = dict(word_separator=87, max_seq_length=40, word_pred_prob=0.15, phoneme_mask_prob=0.8,
                          replace_prob=0.1)
    model_params = dict(hidden_size=32, num_attention_heads=2, intermediate_size=64, max_position_embeddings=48,
                        num_hidden_layers=3, embedding_size=16)
    acc = types.SimpleNamespace(device=torch.device("cpu"))
    criterion = torch.nn.CrossEntropyLoss()

    # ---------------- integer fixtures: a1 length_to_mask --------------------------------------
    lens_cases = [[5, 3, 1], [40, 40], [1], [7, 7, 2, 9, 4]]
    mask_out = {}
    for i, lens in enumerate(lens_cases):
        m = ref_train.length_to_mask(torch.Tensor(lens))
        mask_out[f"lengths_{i}"] = np.asarray(lens, dtype=np.int64)
        mask_out[f"pad_mask_{i}"] = m.numpy()
        mask_out[f"attn_mask_{i}"] = (~m).int().numpy()
    np.savez_compressed(os.path.join(OUT, "length_to_mask.npz"), **mask_out)

    # ---------------- case A: MultiTaskModel, two heads, padded variable-length batch ----------
    np.random.seed(7)
    random.seed(7)
    rng = np.random.default_rng(7)
    recs = synthetic_records(6, rng, 2, 9, symbols, Vt)
    ds = ref_data.MaskedPhonemeDataset(recs, use_token_ids=True, **{**dataset_params, "word_pred_prob": 0.35})
    batch = ref_data.Collater()([ds[i] for i in range(len(ds))])
    token_ids, labels, masked, lengths, midx = batch
    # make sure the edge cases the loss handles are present: one sample without masked rows
    midx = [list(m) for m in midx]
    midx[-1] = []
    if all(len(m) == 0 for m in midx[:-1]):
        midx[0] = [0, 1]

    torch.manual_seed(0)
    cfg = AlbertConfig(vocab_size=Vp, **model_params)
    net = ref_model.MultiTaskModel(AlbertModel(cfg), num_phonemes=Vp, num_tokens=Vt,
                                   hidden_size=model_params["hidden_size"]).double()
    with torch.no_grad():  # non-trivial biases / LN affine so every gradient path is exercised
        for n, p in net.named_parameters():
            if n.endswith("bias") or "LayerNorm" in n or "layer_norm" in n:
                p.add_(0.05 * torch.randn_like(p))
    net.train()
    init_sd = state_to_np(net.state_dict())
    opt = torch.optim.AdamW(net.parameters(), lr=7e-5)

    out = {f"param/{k}": v for k, v in init_sd.items()}
    out["token_ids"], out["labels"], out["masked"] = token_ids.numpy(), labels.numpy(), masked.numpy()
    out["lengths"] = np.asarray(lengths, dtype=np.int64)
    out["midx_flat"], out["midx_offs"] = ragged(midx)
    out["config"] = np.asarray([Vp, model_params["embedding_size"], model_params["hidden_size"],
                                model_params["num_attention_heads"], model_params["intermediate_size"],
                                model_params["num_hidden_layers"], model_params["max_position_embeddings"], Vt],
                               dtype=np.int64)
    for step in range(2):
        text_mask = ref_train.length_to_mask(torch.Tensor(lengths))
        attn = (~text_mask).int()
        p_pred, t_pred = net(masked, attention_mask=attn)
        loss_p = ref_train.calculate_phoneme_loss(p_pred, labels, lengths, midx, criterion)
        loss_t = ref_train.calculate_phoneme_loss(t_pred, token_ids, lengths, midx, criterion)
        loss = loss_p + loss_t
        opt.zero_grad()
        loss.backward()
        if step == 0:
            out["loss_phoneme"] = np.float64(loss_p.item())
            out["loss_token"] = np.float64(loss_t.item())
            hidden = net.encoder(masked, attention_mask=attn).last_hidden_state
            out["last_hidden"] = hidden.detach().numpy()
            out["phoneme_logits"] = p_pred.detach().numpy()
            for n, p in net.named_parameters():
                if p.grad is not None:
                    out[f"grad/{n}"] = p.grad.detach().numpy().copy()
                else:
                    out[f"nograd/{n}"] = np.zeros(1)
        opt.step()
    out["loss_step2"] = np.float64(loss.item())
    for k, v in state_to_np(net.state_dict()).items():
        out[f"after2/{k}"] = v
    osd = opt.state_dict()
    out["opt_state_keys"] = np.asarray(sorted(osd["state"].keys()), dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "multitask_step.npz"), **out)
    print("case A: loss_p %.6f loss_t %.6f lengths %s masked rows %s" %
          (out["loss_phoneme"], out["loss_token"], lengths, [len(m) for m in midx]))

    # ---------------- case B: PhonemeOnlyModel through R/train.py process_batch -----------------
    np.random.seed(11)
    random.seed(11)
    rng = np.random.default_rng(11)
    recs = synthetic_records(5, rng, 3, 12, symbols, Vt)
    ds = ref_data.MaskedPhonemeDataset(recs, use_token_ids=False, **dataset_params)
    batch = ref_data.PhonemeOnlyCollater()([ds[i] for i in range(len(ds))])
    labels, masked, lengths, midx = batch
    midx = [list(m) for m in midx]
    if all(len(m) == 0 for m in midx):
        midx[0] = [1]
    torch.manual_seed(1)
    net = ref_model.PhonemeOnlyModel(AlbertModel(AlbertConfig(vocab_size=Vp, **model_params)), num_phonemes=Vp,
                                     hidden_size=model_params["hidden_size"]).double()
    net.train()
    out = {f"param/{k}": v for k, v in state_to_np(net.state_dict()).items()}
    out["labels"], out["masked"] = labels.numpy(), masked.numpy()
    out["lengths"] = np.asarray(lengths, dtype=np.int64)
    out["midx_flat"], out["midx_offs"] = ragged(midx)
    out["config"] = np.asarray([Vp, model_params["embedding_size"], model_params["hidden_size"],
                                model_params["num_attention_heads"], model_params["intermediate_size"],
                                model_params["num_hidden_layers"], model_params["max_position_embeddings"], 0],
                               dtype=np.int64)
    loss = ref_train.process_batch(net, (labels, masked, lengths, midx), criterion, acc)
    loss.backward()
    out["loss"] = np.float64(loss.item())
    for n, p in net.named_parameters():
        if p.grad is not None:
            out[f"grad/{n}"] = p.grad.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "phoneme_only_step.npz"), **out)
    print("case B: loss %.6f lengths %s masked rows %s" % (out["loss"], lengths, [len(m) for m in midx]))

    # ---------------- case C: a configuration the CUDA path supports (head_dim 64), fp32 storage ---------
    # Pins the CUDA kernels DIRECTLY to reference outputs (tests/test_step_gpu.py), not only through the oracle.
    np.random.seed(23)
    random.seed(23)
    rng = np.random.default_rng(23)
    gp = dict(hidden_size=128, num_attention_heads=2, intermediate_size=256, max_position_embeddings=64,
              num_hidden_layers=3, embedding_size=32)
    recs = synthetic_records(6, rng, 3, 14, symbols, Vt)
    ds = ref_data.MaskedPhonemeDataset(recs, use_token_ids=True, **{**dataset_params, "max_seq_length": 56,
                                                                    "word_pred_prob": 0.3})
    token_ids, labels, masked, lengths, midx = ref_data.Collater()([ds[i] for i in range(len(ds))])
    midx = [list(m) for m in midx]
    midx[2] = []
    torch.manual_seed(2)
    net = ref_model.MultiTaskModel(AlbertModel(AlbertConfig(vocab_size=Vp, **gp)), num_phonemes=Vp, num_tokens=Vt,
                                   hidden_size=gp["hidden_size"])
    with torch.no_grad():
        for n, p in net.named_parameters():
            if n.endswith("bias") or "LayerNorm" in n or "layer_norm" in n:
                p.add_(0.05 * torch.randn_like(p))
    out = {f"param/{k}": v.astype(np.float32) for k, v in state_to_np(net.state_dict()).items()}  # exact: fp32 init
    net = net.double().train()
    out["token_ids"], out["labels"], out["masked"] = token_ids.numpy(), labels.numpy(), masked.numpy()
    out["lengths"] = np.asarray(lengths, dtype=np.int64)
    out["midx_flat"], out["midx_offs"] = ragged(midx)
    out["config"] = np.asarray([Vp, gp["embedding_size"], gp["hidden_size"], gp["num_attention_heads"],
                                gp["intermediate_size"], gp["num_hidden_layers"], gp["max_position_embeddings"], Vt],
                               dtype=np.int64)
    attn = (~ref_train.length_to_mask(torch.Tensor(lengths))).int()
    p_pred, t_pred = net(masked, attention_mask=attn)
    loss_p = ref_train.calculate_phoneme_loss(p_pred, labels, lengths, midx, criterion)
    loss_t = ref_train.calculate_phoneme_loss(t_pred, token_ids, lengths, midx, criterion)
    (loss_p + loss_t).backward()
    out["loss_phoneme"], out["loss_token"] = np.float64(loss_p.item()), np.float64(loss_t.item())
    out["last_hidden"] = net.encoder(masked, attention_mask=attn).last_hidden_state.detach().numpy().astype(np.float32)
    for n, p in net.named_parameters():
        if p.grad is not None:
            out[f"grad/{n}"] = p.grad.detach().numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "cuda_config_step.npz"), **out)
    print("case C: loss_p %.6f loss_t %.6f lengths %s masked rows %s" %
          (out["loss_phoneme"], out["loss_token"], lengths, [len(m) for m in midx]))

    # all-empty batch: R/train.py:129 returns tensor(0.)
    empty = ref_train.calculate_phoneme_loss(torch.zeros(2, 3, Vp), torch.zeros(2, 3).long(), [3, 2], [[], []], criterion)
    assert float(empty) == 0.0

    # ---------------- batch-contract fixture: collate output types/values (a11) ----------------
    np.savez_compressed(os.path.join(OUT, "collate_contract.npz"), labels=labels.numpy(), masked=masked.numpy(),
                        lengths=np.asarray(lengths, dtype=np.int64), midx_flat=ragged(midx)[0],
                        midx_offs=ragged(midx)[1], pad_id=np.int64(0),
                        mask_id=np.int64(symbols.index(ref_sym.PHONEME_MASK)),
                        sep_id=np.int64(symbols.index(ref_sym.PHONEME_SEPARATOR)), num_symbols=np.int64(Vp))
    full_size_case()
    dataset_case()
    print("wrote fixtures to", os.path.abspath(OUT))


if __name__ == "__main__":
    main()
