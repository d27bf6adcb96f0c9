"""BASELINE.json C5: encoder-only forward (last_hidden_state), 64 x 256 tokens, bf16, no grad — the call StyleTTS makes
(R/README.md:91).  This path (`model.encode`) vs HF AlbertModel under bf16 autocast on the same GPU.
    python profiles/c5_encoder_bench.py [--iters 50]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def timeit(fn, iters):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--seq", type=int, default=256)
    args = ap.parse_args()
    import plbert_b200 as P
    from transformers import AlbertConfig, AlbertModel
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    enc = AlbertModel(AlbertConfig(vocab_size=bench.NUM_PHONEMES, **bench.MODEL_PARAMS)).to(dev).eval()
    model = P.PhonemeOnlyModel(enc, bench.NUM_PHONEMES, 768).to(dev).eval()
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(1, 185, (args.batch, args.seq), generator=g).to(dev)
    mask = torch.ones_like(ids, dtype=torch.int32)
    tokens = args.batch * args.seq

    def ours():
        return model.encode(ids, mask)

    def ours_graph():
        return model.encode(ids, mask, use_graph=True)

    def hf():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return enc(ids, attention_mask=mask).last_hidden_state

    a, b = ours().float(), hf().float()
    cos = torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0).item()
    assert torch.equal(ours(), ours_graph())
    ms_o, ms_g, ms_h = timeit(ours, args.iters), timeit(ours_graph, args.iters), timeit(hf, args.iters)
    print(json.dumps({"config": f"C5: encoder-only forward {args.batch}x{args.seq}, bf16, no grad", "cosine_vs_hf": cos,
                      "plbert_b200": {"ms": ms_o, "tokens_per_s": tokens / (ms_o * 1e-3)},
                      "plbert_b200_cuda_graph": {"ms": ms_g, "tokens_per_s": tokens / (ms_g * 1e-3)},
                      "hf_autocast_bf16": {"ms": ms_h, "tokens_per_s": tokens / (ms_h * 1e-3)}}))


if __name__ == "__main__":
    main()
