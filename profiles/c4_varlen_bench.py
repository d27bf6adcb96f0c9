"""BASELINE.json C4: 192 sequences with lengths ~U[32, 512] (phoneme-only step), padded [B, L] layout vs the packed
(cu_seqlens) layout that loss_and_backward uses for variable-length batches.  Reports valid tokens/s.
    python profiles/c4_varlen_bench.py [--steps 8]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=8)
    args = ap.parse_args()
    import plbert_b200 as P
    from plbert_b200.batch import synthetic_collate_batch
    from transformers import AlbertConfig, AlbertModel
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    enc = AlbertModel(AlbertConfig(vocab_size=bench.NUM_PHONEMES, **bench.MODEL_PARAMS))
    model = P.PhonemeOnlyModel(enc, bench.NUM_PHONEMES, 768).to(dev)
    opt = P.FusedAdamW(model.parameters(), lr=7e-5)
    batches = [synthetic_collate_batch(bench.BATCH, bench.SEQ, seed=77 + i, variable_length=True, min_length=32) for i in range(2)]
    tokens = [int(sum(b[2])) for b in batches]
    for packed in (False, True):
        model.pack_sequences = packed
        for i in range(3):
            model.loss_and_backward(batches[i % 2])
            opt.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            loss = model.loss_and_backward(batches[i % 2])
            opt.step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        tok = sum(tokens[i % 2] for i in range(args.steps)) / args.steps
        print(json.dumps({"config": "C4: 192 sequences, lengths U[32,512], phoneme-only step, host batch -> step (e2e path)",
                          "layout": "packed" if packed else "padded", "ms_per_step": ms, "valid_tokens_per_step": tok,
                          "rows_per_step": model.last_batch.T, "valid_tokens_per_s": tok / (ms * 1e-3), "loss": float(loss)}))


if __name__ == "__main__":
    main()
