"""Where does the end-to-end step lose time against the device-resident one?  Alternates the two loops of bench.py on the
same box (A = resident batches, B = BatchFeeder + loss.item() every step) and prints ms per step of each pass, plus B
variants: without the per-step loss read-back, and with the workers idle (batches pre-drawn)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    import plbert_b200 as P
    from plbert_b200.batch import BatchFeeder
    from transformers import AlbertConfig, AlbertModel
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    enc = AlbertModel(AlbertConfig(vocab_size=bench.NUM_PHONEMES, **bench.MODEL_PARAMS))
    model = P.MultiTaskModel(enc, bench.NUM_PHONEMES, bench.NUM_TOKENS, 768).to(dev)
    opt = P.FusedAdamW(model.parameters(), lr=7e-5)
    cores = os.cpu_count() or 8
    workers = max(2, min(12, cores - 2))
    loader = bench.make_loader(0, True, num_workers=workers)
    stream = loader.endless()
    host = [next(stream) for _ in range(2)]
    eng = model._ensure_engine(dev)
    heads = ["phoneme_predictor", "token_predictor"]
    resident = [model.stage_batch(hb, stager=model.new_stager()) for hb in host]
    steps = 10

    def timed(fn, n=steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n

    def dev_step(i):
        eng.loss_and_backward(resident[i % 2], heads)
        eng.adamw_step(7e-5, grad_scale=opt.grad_scale)

    def dev_step_item(i):
        dev_step(i)
        eng.loss_buf[0].item()

    feed = BatchFeeder(model, stream)

    def api_step(i):
        loss = feed.step()
        opt.step()
        return loss.item()

    def api_step_noitem(i):
        feed.step()
        opt.step()

    pre = [next(stream) for _ in range(steps + 2)]
    feed2 = [None]

    def api_step_predrawn(i):
        loss = feed2[0].step()
        opt.step()
        return loss.item()

    for i in range(3):
        dev_step(i)
    for name, fn in [("A resident", dev_step), ("B e2e (feeder + item)", api_step), ("A resident", dev_step),
                     ("A resident + item()", dev_step_item), ("B e2e without item()", api_step_noitem),
                     ("B e2e (feeder + item)", api_step), ("A resident", dev_step)]:
        ms, wall = timed(fn)
        print(f"{name:28s} {ms:7.2f} ms/step (events)  {wall:7.2f} ms/step (wall)", flush=True)
    feed2[0] = BatchFeeder(model, iter(pre))
    ms, wall = timed(api_step_predrawn)
    print(f"{'B e2e, batches pre-drawn':28s} {ms:7.2f} ms/step (events)  {wall:7.2f} ms/step (wall)")
    # host cost of the per-step pieces
    t0 = time.perf_counter()
    for hb in pre[:5]:
        model.stage_batch(hb, stager=feed._stager)
    torch.cuda.synchronize()
    print(f"stage_batch host time: {(time.perf_counter() - t0) * 1e3 / 5:.2f} ms per batch")
    t0 = time.perf_counter()
    for _ in range(5):
        next(stream)
    print(f"next(stream) (workers ahead): {(time.perf_counter() - t0) * 1e3 / 5:.2f} ms per batch")
    t0 = time.perf_counter()
    eng.loss_and_backward(resident[0], heads)
    eng.adamw_step(7e-5, grad_scale=opt.grad_scale)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"host enqueue of one step: {(t1 - t0) * 1e3:.2f} ms")


if __name__ == "__main__":
    main()
