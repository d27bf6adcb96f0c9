"""Export a PL-BERT training checkpoint to a Hugging Face directory.

Same command line and output files as the reference converter (R/convert_to_hf.py:9-64)::

    python -m plbert_b200.convert_to_hf --checkpoint_path runs/x/step_1000.t7 --config_path config.yml --output_dir out/

writes ``out/config.json`` + weights (``AlbertModel.save_pretrained``, R/convert_to_hf.py:49), ``out/pl_bert_full_model.pt``
(state dict of the full model incl. heads, :52), ``out/training_metadata.txt`` (:55-58) and ``out/config.yml`` (:61-62);
``load_pl_bert_model(out)`` is the loader of R/convert_to_hf.py:66-102.

Differences, all forced by facts recorded in SURVEY.md §0:

* the reference builds ``AlbertConfig(**model_params)`` without ``vocab_size`` (R/convert_to_hf.py:27), which yields a
  30 000-row embedding that cannot take a 188-row checkpoint; here every size that the checkpoint determines
  (``vocab_size``, ``num_phonemes``, ``num_tokens``) is read from the checkpoint tensors themselves;
* the reference needs the AraBERT tokenizer only for ``tokenizer.vocab_size`` (:24,32); it is optional here
  (``--num_tokens`` or the ``token_predictor.weight`` shape replace it), so the tool works offline;
* phoneme-only checkpoints (what R/train.py actually writes, R/train.py:266-270) are accepted: the exported full model
  then has no ``token_predictor``;
* upstream yl4579/PL-BERT head names (``mask_predictor`` / ``word_predictor``) and the DDP ``module.`` prefix are
  normalised on the way in.

Pure state-dict work: nothing here computes, so it runs without a GPU.
"""
from __future__ import annotations

import argparse
import os
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch
import yaml

from .model import MultiTaskModel, PhonemeOnlyModel, normalize_state_dict_keys

_WORD_EMB = "encoder.embeddings.word_embeddings.weight"
# keys of config['model_params'] that are not AlbertConfig fields in this fork (R/configs/config.yml:32-39)
_NON_ALBERT_KEYS = ("pretrained_model", "dropout")


def read_checkpoint(path: str) -> Tuple[Dict[str, torch.Tensor], Dict[str, object]]:
    """``(normalised state dict, {'step', 'epoch'})`` of a ``step_N.<ext>`` file (R/train.py:416-421)."""
    ck = torch.load(path, map_location="cpu", weights_only=False)
    if "net" not in ck:
        raise KeyError(f"{path}: not a PL-BERT checkpoint (no 'net' entry)")
    meta = {"step": ck.get("step"), "epoch": ck.get("epoch")}
    return normalize_state_dict_keys(ck["net"]), meta


def albert_config_from(model_params: Dict[str, object], vocab_size: int):
    from transformers import AlbertConfig
    kwargs = {k: v for k, v in model_params.items() if k not in _NON_ALBERT_KEYS and k != "vocab_size"}
    return AlbertConfig(vocab_size=vocab_size, **kwargs)


def build_model_for(state: Dict[str, torch.Tensor], model_params: Dict[str, object], num_tokens: Optional[int] = None):
    """Model skeleton whose sizes come from the checkpoint tensors (vocab, head widths) and from ``model_params``."""
    from transformers import AlbertModel
    if _WORD_EMB not in state or "phoneme_predictor.weight" not in state:
        raise KeyError("checkpoint lacks encoder.embeddings.word_embeddings.weight / phoneme_predictor.weight")
    vocab = int(state[_WORD_EMB].shape[0])
    hidden = int(model_params["hidden_size"])
    num_phonemes = int(state["phoneme_predictor.weight"].shape[0])
    encoder = AlbertModel(albert_config_from(model_params, vocab))
    if "token_predictor.weight" in state:
        vt = int(state["token_predictor.weight"].shape[0])
        if num_tokens is not None and num_tokens != vt:
            raise ValueError(f"--num_tokens {num_tokens} does not match the checkpoint's token head ({vt} rows)")
        return MultiTaskModel(encoder, num_phonemes, vt, hidden)
    return PhonemeOnlyModel(encoder, num_phonemes, hidden)


def convert(checkpoint_path: str, config_path: str, output_dir: str, num_tokens: Optional[int] = None) -> str:
    with open(config_path) as f:
        config = yaml.safe_load(f)
    state, meta = read_checkpoint(checkpoint_path)
    model = build_model_for(state, config["model_params"], num_tokens)
    missing, unexpected = model.load_state_dict(state, strict=False)   # R/convert_to_hf.py:43
    # buffers that old transformers versions persisted are the only tolerated leftovers
    bad = [k for k in unexpected if not k.endswith(("position_ids", "token_type_ids"))]
    if missing or bad:
        raise RuntimeError(f"checkpoint does not fit the model: missing {sorted(missing)}, unexpected {sorted(bad)}")
    os.makedirs(output_dir, exist_ok=True)
    model.encoder.save_pretrained(output_dir)
    torch.save(OrderedDict((k, v.detach().clone()) for k, v in model.state_dict().items()),
               os.path.join(output_dir, "pl_bert_full_model.pt"))
    with open(os.path.join(output_dir, "training_metadata.txt"), "w") as f:
        f.write(f"Original checkpoint: {checkpoint_path}\nStep: {meta['step']}\nEpoch: {meta['epoch']}\n")
    with open(os.path.join(output_dir, "config.yml"), "w") as f:
        yaml.dump(config, f)
    return output_dir


def load_pl_bert_model(model_dir: str):
    """R/convert_to_hf.py:66-102 without the tokenizer download: returns the full model in eval mode."""
    from transformers import AlbertModel
    with open(os.path.join(model_dir, "config.yml")) as f:
        config = yaml.safe_load(f)
    full = torch.load(os.path.join(model_dir, "pl_bert_full_model.pt"), map_location="cpu", weights_only=True)
    encoder = AlbertModel.from_pretrained(model_dir)
    hidden = int(config["model_params"]["hidden_size"])
    num_phonemes = int(full["phoneme_predictor.weight"].shape[0])
    if "token_predictor.weight" in full:
        model = MultiTaskModel(encoder, num_phonemes, int(full["token_predictor.weight"].shape[0]), hidden)
    else:
        model = PhonemeOnlyModel(encoder, num_phonemes, hidden)
    model.load_state_dict(full)
    return model.eval()


def styletts_encoder_state_dict(checkpoint_path: str) -> "OrderedDict[str, torch.Tensor]":
    """The ``AlbertModel`` state dict the StyleTTS snippet extracts (R/README.md:57-66), for checkpoints written with or
    without the ``module.`` prefix (the snippet's ``k[7:]`` silently mangles the latter)."""
    state, _ = read_checkpoint(checkpoint_path)
    return OrderedDict((k[len("encoder."):], v) for k, v in state.items() if k.startswith("encoder."))


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(description="Convert a PL-BERT checkpoint to Hugging Face format")
    ap.add_argument("--checkpoint_path", required=True, help="step_N.t7 / .pth written by training")
    ap.add_argument("--config_path", required=True, help="the training config.yml")
    ap.add_argument("--output_dir", required=True)
    ap.add_argument("--num_tokens", type=int, default=None, help="token-head vocabulary (checked against the checkpoint)")
    a = ap.parse_args(argv)
    out = convert(a.checkpoint_path, a.config_path, a.output_dir, a.num_tokens)
    print(f"Model successfully converted and saved to {out}")


if __name__ == "__main__":
    main()
