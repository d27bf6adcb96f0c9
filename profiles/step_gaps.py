"""GPU busy time vs. wall span of one training step (C2 size), from CUPTI kernel records (torch.profiler).

Answers "how much of ms_per_step is the GPU idle between launches": prints the step span, the summed kernel durations,
the idle remainder, and the kernels ranked by total time.  python profiles/step_gaps.py [--model multitask]
"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="phoneme_only")
    ap.add_argument("--steps", type=int, default=3)
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
    batch = bench.make_batch(1234, bench.BATCH, two)
    for _ in range(3):
        model.loss_and_backward(batch)
        opt.step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.steps):
            model.loss_and_backward(batch)
            opt.step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
    busy = sum(e.time_range.end - e.time_range.start for e in evs)
    gaps = collections.Counter()
    for a, b in zip(evs, evs[1:]):
        g = b.time_range.start - a.time_range.end
        if g > 0:
            gaps[(a.name[:40], b.name[:40])] += g
    n = args.steps
    print(f"span/step {(t1 - t0) / n / 1e3:.2f} ms   busy/step {busy / n / 1e3:.2f} ms   idle/step {(t1 - t0 - busy) / n / 1e3:.2f} ms   "
          f"launches/step {len(evs) / n:.0f}")
    tot = collections.Counter()
    cnt = collections.Counter()
    for e in evs:
        tot[e.name[:70]] += e.time_range.end - e.time_range.start
        cnt[e.name[:70]] += 1
    for k, v in tot.most_common(25):
        print(f"{v / n / 1e3:8.3f} ms  x{cnt[k] / n:6.1f}  {k}")
    print("largest gaps (us/step):")
    for k, v in gaps.most_common(12):
        print(f"{v / n:8.1f}  {k[0]} -> {k[1]}")


if __name__ == "__main__":
    main()
