"""Times the attention kernels at the BASELINE shape (192 sequences x 512 rows x 12 heads x 64) through the C ABI."""
import math
import re
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from plbert_b200._abi import call, load, ptr  # noqa: E402


def main():
    load()
    dev = torch.device("cuda:0")
    B, L, nh = 192, 512, 12
    H, T = nh * 64, B * L
    if os.environ.get("VARLEN"):   # BASELINE configs[3]: lengths U[32, 512], sorted by length like the collate functions, packed rows
        g = torch.Generator().manual_seed(int(os.environ["VARLEN"]))
        lens_l = sorted(torch.randint(32, 513, (B,), generator=g).tolist(), reverse=True)
        T = sum(lens_l)
        lens = torch.tensor(lens_l, device=dev, dtype=torch.int32)
        seq_start = torch.tensor([0] + torch.tensor(lens_l).cumsum(0).tolist()[:-1], device=dev, dtype=torch.int32)
    else:
        lens_l = [L] * B
        lens = torch.full((B,), L, device=dev, dtype=torch.int32)
        seq_start = torch.arange(B, device=dev, dtype=torch.int32) * L
    qkv = torch.randn(T, 3 * H, device=dev).to(torch.bfloat16)
    ctx = torch.empty(T, H, device=dev, dtype=torch.bfloat16)
    dctx = torch.randn(T, H, device=dev).to(torch.bfloat16)
    dqkv = torch.empty(T, 3 * H, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(nh, T, device=dev)
    delta = torch.empty(nh, T, device=dev)
    dbias = None if os.environ.get("NO_DBIAS") else torch.zeros(3 * H, device=dev)  # NO_DBIAS=1: skip the fused bias sums
    S = torch.cuda.current_stream().cuda_stream
    fwd_flops = 4.0 * nh * sum(n * n for n in lens_l) * 64

    def timeit(fn, flops, label, iters=10):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"{label:28s} {ms * 1e3:9.1f} us  {flops / ms * 1e-9:8.1f} TFLOP/s (algorithmic)")

    names = [n for n in ("plb_attn_fwd", "plb_attn_fwd_tc") if hasattr(load(), n)]
    for n in names:
        timeit(lambda: call(n, ptr(qkv), 3 * H, ptr(seq_start), ptr(lens), ptr(lens), B, nh, 64, L, T, ptr(ctx), H, ptr(lse), S),
               fwd_flops, n)
    for n in [n for n in ("plb_attn_bwd", "plb_attn_bwd_tc") if hasattr(load(), n)]:
        extra = (ptr(dbias),) if n.endswith("_tc") else ()
        timeit(lambda: call(n, ptr(qkv), 3 * H, ptr(seq_start), ptr(lens), ptr(lens), B, nh, 64, L, T, ptr(ctx), ptr(dctx), H,
                            ptr(lse), ptr(delta), ptr(dqkv), 3 * H, *extra, S), 2.5 * fwd_flops, n)

    if hasattr(load(), "plb_attn_bwd_fused"):
        nbytes = load().plb_attn_bwd_fused_workspace_bytes(B, nh)
        ws = torch.empty(nbytes // 4, device=dev)
        for use_ctx in (ptr(ctx), 0):   # 0: delta given (the step's configuration: the O-projection DGRAD produced it)
            timeit(lambda: call("plb_attn_bwd_fused", ptr(qkv), 3 * H, ptr(seq_start), ptr(lens), ptr(lens), B, nh, 64, L, T,
                                use_ctx, ptr(dctx), H, ptr(lse), ptr(delta), ptr(dqkv), 3 * H, ptr(dbias), ptr(ws), nbytes, S),
                   2.5 * fwd_flops, "plb_attn_bwd_fused" + ("" if use_ctx else " (delta given)"))

    # per-kernel durations of one backward call (CUPTI)
    from torch.profiler import ProfilerActivity, profile
    n = "plb_attn_bwd_tc"
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            call("plb_attn_fwd_tc", ptr(qkv), 3 * H, ptr(seq_start), ptr(lens), ptr(lens), B, nh, 64, L, T, ptr(ctx), H, ptr(lse), S)
            call(n, ptr(qkv), 3 * H, ptr(seq_start), ptr(lens), ptr(lens), B, nh, 64, L, T, ptr(ctx), ptr(dctx), H,
                 ptr(lse), ptr(delta), ptr(dqkv), 3 * H, ptr(dbias), S)
        torch.cuda.synchronize()
    import collections
    tot = collections.Counter()
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            m = re.search(r"attn_\w+", e.name)
            tot[m.group(0) if m else e.name[:40]] += (e.time_range.end - e.time_range.start) / 5
    for k, v in tot.most_common():
        print(f"  {v:9.1f} us  {k}")


if __name__ == "__main__":
    main()
