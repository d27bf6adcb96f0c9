"""oracle/reference_port.py (the CPU timing port used by bench.py) computes the same step as the numpy oracle and,
when /root/reference is mounted (build container only), as the live reference."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import plbert_oracle as O
from oracle import reference_port as RP
from plbert_b200.batch import synthetic_collate_batch

MP = dict(hidden_size=32, num_attention_heads=2, intermediate_size=64, max_position_embeddings=64, num_hidden_layers=2,
          embedding_size=16)
CFG = O.OracleConfig(vocab_size=188, embedding_size=16, hidden_size=32, num_attention_heads=2, intermediate_size=64,
                     num_hidden_layers=2, max_position_embeddings=64)


def _batch():
    return synthetic_collate_batch(4, 48, seed=3, variable_length=True, min_length=8, with_token_ids=True, num_tokens=60,
                                   word_pred_prob=0.4)


def test_port_step_matches_oracle():
    batch = _batch()
    tok, lab, msk, lens, midx = batch
    model = RP.build(188, 60, MP, seed=1).double()
    params = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
    opt = torch.optim.AdamW(model.parameters(), lr=7e-5)
    loss = RP.train_step(model, opt, batch)
    want, _, grads, _ = O.loss_and_grads(params, CFG, msk.numpy(), lab.numpy(), lens, midx, token_labels=tok.numpy())
    assert abs(loss - want) <= 1e-10 * abs(want)
    O.adamw_step(params, grads, {})
    for k, v in model.state_dict().items():
        if "key.bias" in k:
            continue
        assert np.allclose(v.detach().numpy(), params[k], rtol=0, atol=1e-9), k


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference is only mounted in the build container")
def test_port_matches_live_reference():
    sys.path.insert(0, "/root/reference")
    from transformers import AlbertConfig, AlbertModel
    if "accelerate" not in sys.modules:
        acc = types.ModuleType("accelerate")
        acc.Accelerator = object
        acc.DistributedDataParallelKwargs = lambda **k: None
        sys.modules["accelerate"] = acc
    import model as ref_model
    import train as ref_train
    batch = _batch()
    tok, lab, msk, lens, midx = batch
    port = RP.build(188, 60, MP, seed=2).double()
    torch.manual_seed(0)
    ref = ref_model.MultiTaskModel(AlbertModel(AlbertConfig(vocab_size=188, **MP)), 188, 60, 32).double()
    ref.load_state_dict(port.state_dict())
    crit = torch.nn.CrossEntropyLoss()
    attn = (~ref_train.length_to_mask(torch.Tensor(lens))).int()
    p, t = ref(msk, attention_mask=attn)
    want = ref_train.calculate_phoneme_loss(p, lab, lens, midx, crit) + ref_train.calculate_phoneme_loss(t, tok, lens, midx, crit)
    opt = torch.optim.SGD(port.parameters(), lr=0.0)
    got = RP.train_step(port, opt, batch)
    assert abs(got - want.item()) <= 1e-12 * abs(want.item())
