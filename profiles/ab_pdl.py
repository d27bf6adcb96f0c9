"""Same-process A/B of programmatic dependent launch (plb_set_launch_overlap): blocks of device-resident two-head steps at the
BASELINE size with the switch off / on in alternation (the board drifts with time, so only neighbouring blocks compare), and the
loss of the SAME batch from the SAME weights under both settings."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    import plbert_b200 as P
    from plbert_b200._abi import call
    from transformers import AlbertConfig, AlbertModel
    dev = torch.device("cuda:0")
    model_name = sys.argv[1] if len(sys.argv) > 1 else "multitask"
    two = model_name == "multitask"
    torch.manual_seed(0)
    enc = AlbertModel(AlbertConfig(vocab_size=bench.NUM_PHONEMES, **bench.MODEL_PARAMS))
    model = (P.MultiTaskModel(enc, bench.NUM_PHONEMES, bench.NUM_TOKENS, 768) if two else
             P.PhonemeOnlyModel(enc, bench.NUM_PHONEMES, 768)).to(dev)
    opt = P.FusedAdamW(model.parameters(), lr=7e-5)
    eng = model._ensure_engine(dev)
    heads = ["phoneme_predictor", "token_predictor"] if two else ["phoneme_predictor"]
    batches = [model.stage_batch(bench.make_batch(1234 + i, bench.BATCH, two), stager=model.new_stager()) for i in range(2)]
    # same weights, same batch, both settings: the loss must agree (fp32 atomics order aside)
    losses = {}
    for on in (0, 1, 0, 1):
        call("plb_set_launch_overlap", on)
        eng.loss_and_backward(batches[0], heads)
        losses.setdefault(on, []).append(float(eng.loss_buf[0].item()))
        eng.zero_grad() if hasattr(eng, "zero_grad") else None
    print("loss, overlap off:", losses[0], " on:", losses[1], flush=True)

    def block(n=10):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            eng.loss_and_backward(batches[i % 2], heads)
            eng.adamw_step(7e-5, grad_scale=opt.grad_scale)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for on in (1, 0):
        call("plb_set_launch_overlap", on)
        block(3)
    res = {0: [], 1: []}
    for rnd in range(5):
        for on in (0, 1):
            call("plb_set_launch_overlap", on)
            ms = block()
            res[on].append(ms)
            print(f"round {rnd} overlap {'on ' if on else 'off'} {ms:7.3f} ms/step", flush=True)
    a, b = sum(res[0]) / len(res[0]), sum(res[1]) / len(res[1])
    print(f"mean: off {a:.3f}  on {b:.3f}  ({(a - b) / a * 100:+.2f} % faster with overlap)")


if __name__ == "__main__":
    main()
