"""Informational "PyTorch-library bar" (SURVEY.md §8d): the reference training step executed on ONE B200 through stock
libraries — HF AlbertModel under torch.autocast(bf16) (SDPA attention), nn.Linear heads on every position, the reference's
per-sample loss loop (R/train.py:107-131), torch.optim.AdamW — at the BASELINE size (192 x 512).  Not part of the product
and not the reference arm of bench.py (that one is the CPU port); it only says what cuBLAS/SDPA/foreach-AdamW give on the
same box.  python profiles/torch_library_bar.py [--model multitask] [--steps 5]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch import nn  # noqa: E402

import bench  # noqa: E402


class Ref(nn.Module):  # R/model.py:5-30
    def __init__(self, enc, two):
        super().__init__()
        self.encoder = enc
        self.phoneme_predictor = nn.Linear(768, bench.NUM_PHONEMES)
        if two:
            self.token_predictor = nn.Linear(768, bench.NUM_TOKENS)

    def forward(self, ids, attention_mask=None):
        h = self.encoder(ids, attention_mask=attention_mask).last_hidden_state
        return [self.phoneme_predictor(h)] + ([self.token_predictor(h)] if hasattr(self, "token_predictor") else [])


def masked_loss(pred, labels, lengths, midx, crit):  # R/train.py:107-131
    loss, count = 0, 0
    for p, t, n, m in zip(pred, labels, lengths, midx):
        if len(m) > 0:
            loss = loss + crit(p[:n][m], t[:n][m])
            count += 1
    return loss / count


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="phoneme_only")
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    from transformers import AlbertConfig, AlbertModel
    dev = torch.device("cuda:0")
    two = args.model == "multitask"
    torch.manual_seed(0)
    model = Ref(AlbertModel(AlbertConfig(vocab_size=bench.NUM_PHONEMES, **bench.MODEL_PARAMS)), two).to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=7e-5)
    crit = nn.CrossEntropyLoss()
    hb = bench.make_batch(1234, bench.BATCH, two)
    if two:
        tok, lab, msk, lens, midx = hb
        tok = tok.to(dev)
    else:
        lab, msk, lens, midx = hb
    lab, msk = lab.to(dev), msk.to(dev)

    def step():
        mask = (torch.arange(msk.shape[1], device=dev)[None, :] < torch.tensor(lens, device=dev)[:, None]).int()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            preds = model(msk, attention_mask=mask)
        loss = masked_loss(preds[0].float(), lab, lens, midx, crit)
        if two:
            loss = loss + masked_loss(preds[1].float(), tok, lens, midx, crit)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"what": "reference step through PyTorch libraries on 1 B200 (autocast bf16, SDPA, torch AdamW)",
                      "model": args.model, "ms_per_step": ms, "tokens_per_s": sum(lens) / (ms * 1e-3), "loss": float(loss),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}))


if __name__ == "__main__":
    main()
