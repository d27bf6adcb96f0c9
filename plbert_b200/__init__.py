"""plbert_b200 — B200-native (sm_100a) implementation of PL-BERT's training hot path.

Public surface (mirrors the reference, Fadi987/PL-BERT):
  MultiTaskModel, PhonemeOnlyModel   drop-ins for R/model.py
  fuse_albert / FusedAlbertEncoder   AlbertModel whose forward(ids, attention_mask=...).last_hidden_state runs the kernels
                                     (R/model.py:14-15, StyleTTS R/README.md:91); differentiable
  FusedAdamW                         drop-in for torch.optim.AdamW as used at R/train.py:272
  train.process_batch / length_to_mask / save_checkpoint / load_checkpoint / find_latest_checkpoint
  BatchFeeder                        look-ahead input staging (host flatten + pinned H2D of batch i+1 under step i)
  ddp.GradientSynchronizer           bucketed NCCL gradient all-reduce overlapped with backward
The compute lives in libplbert_b200.so (hand-written CUDA, C ABI in include/plbert_b200.h); importing this
package does not need a GPU, computing with it does — there is no CPU fallback.
"""
from . import _abi
from ._abi import PlbError
from .batch import BatchFeeder
from .model import FusedAlbertEncoder, MultiTaskModel, PhonemeOnlyModel, fuse_albert, normalize_state_dict_keys
from .optim import FusedAdamW

__all__ = ["MultiTaskModel", "PhonemeOnlyModel", "fuse_albert", "FusedAlbertEncoder", "FusedAdamW", "BatchFeeder", "PlbError", "normalize_state_dict_keys", "_abi"]
