"""Same-box A/B of engine switches on the BASELINE step (192 x 512, two heads): alternates the settings in one process.
    python profiles/ab_step.py fuse_ffn1_bias_sums            # engine attribute toggled between True and False"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main(attr, rounds=3, steps=8):
    import plbert_b200 as P
    from transformers import AlbertConfig, AlbertModel
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    enc = AlbertModel(AlbertConfig(vocab_size=bench.NUM_PHONEMES, **bench.MODEL_PARAMS))
    model = P.MultiTaskModel(enc, bench.NUM_PHONEMES, bench.NUM_TOKENS, 768).to(dev)
    opt = P.FusedAdamW(model.parameters(), lr=7e-5)
    from plbert_b200.batch import synthetic_collate_batch
    batches = [model.stage_batch(synthetic_collate_batch(bench.BATCH, 512, seed=s, with_token_ids=True, num_tokens=bench.NUM_TOKENS))
               for s in (1, 2)]
    for i in range(3):
        model.loss_and_backward(batches[i & 1]); opt.step()
    eng = model._engine
    res = {True: [], False: []}
    for r in range(rounds):
        for val in (True, False):
            setattr(eng, attr, val)
            for i in range(2):
                model.loss_and_backward(batches[i & 1]); opt.step()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                model.loss_and_backward(batches[i & 1]); opt.step()
            e1.record()
            torch.cuda.synchronize()
            res[val].append(e0.elapsed_time(e1) / steps)
    for val in (True, False):
        print(f"{attr} = {val}: " + ", ".join(f"{x:.2f}" for x in res[val]) + " ms/step")


if __name__ == "__main__":
    main(sys.argv[1])
