"""One profiled training step at the BASELINE size (C2: 192x512, config.yml ALBERT) for ncu.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python profiles/run_step.py [--model phoneme_only|multitask] [--batch 192]

Two warm-up steps run unprofiled; cudaProfilerStart/Stop bracket exactly one step (forward, loss, backward, AdamW).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="phoneme_only")
    ap.add_argument("--batch", type=int, default=bench.BATCH)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    import plbert_b200 as P
    from transformers import AlbertConfig, AlbertModel
    dev = torch.device("cuda:0")
    two = args.model == "multitask"
    torch.manual_seed(0)
    enc = AlbertModel(AlbertConfig(vocab_size=bench.NUM_PHONEMES, **bench.MODEL_PARAMS))
    model = (P.MultiTaskModel(enc, bench.NUM_PHONEMES, bench.NUM_TOKENS, 768) if two else
             P.PhonemeOnlyModel(enc, bench.NUM_PHONEMES, 768)).to(dev)
    opt = P.FusedAdamW(model.parameters(), lr=7e-5)
    batch = bench.make_batch(1234, args.batch, two)
    for _ in range(args.warmup):
        model.loss_and_backward(batch)
        opt.step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    loss = model.loss_and_backward(batch)
    opt.step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("loss", loss.item())


if __name__ == "__main__":
    main()
