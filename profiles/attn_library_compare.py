"""Context for the attention numbers: the same shape (192 x 12 heads x 512 x 64, bf16, no mask needed at full length) through the
attention LIBRARIES of this image — torch SDPA per backend (cuDNN, FlashAttention-2, memory-efficient) and the flash_attn package —
next to the hand-written kernels (C ABI).  Forward and backward timed separately with CUDA events, 10 launches after 3."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from torch.nn.attention import SDPBackend, sdpa_kernel  # noqa: E402

from plbert_b200._abi import call, load, ptr  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    load()
    dev = torch.device("cuda:0")
    B, L, nh, D = 192, 512, 12, 64
    H, T = nh * D, B * L
    fwd_flops = 4.0 * B * nh * L * L * D
    torch.manual_seed(0)
    qkv = torch.randn(T, 3 * H, device=dev).to(torch.bfloat16)
    dctx = torch.randn(T, H, device=dev).to(torch.bfloat16)
    S = torch.cuda.current_stream().cuda_stream
    lens = torch.full((B,), L, device=dev, dtype=torch.int32)
    ss = torch.arange(B, device=dev, dtype=torch.int32) * L
    ctx = torch.empty(T, H, device=dev, dtype=torch.bfloat16)
    dqkv = torch.empty(T, 3 * H, device=dev, dtype=torch.bfloat16)
    lse, delta = torch.empty(nh, T, device=dev), torch.empty(nh, T, device=dev)
    dbias = torch.zeros(3 * H, device=dev)
    nbytes = load().plb_attn_bwd_fused_workspace_bytes(B, nh)
    ws = torch.empty(nbytes // 4, device=dev)
    rows = []
    f = timeit(lambda: call("plb_attn_fwd_tc", ptr(qkv), 3 * H, ptr(ss), ptr(lens), ptr(lens), B, nh, D, L, T, ptr(ctx), H, ptr(lse), S))
    b = timeit(lambda: call("plb_attn_bwd_fused", ptr(qkv), 3 * H, ptr(ss), ptr(lens), ptr(lens), B, nh, D, L, T, ptr(ctx), ptr(dctx), H,
                            ptr(lse), ptr(delta), ptr(dqkv), 3 * H, ptr(dbias), ptr(ws), nbytes, S))
    rows.append(("this repo (tcgen05, C ABI; backward incl. delta + QKV bias sums)", f, b))

    q, k, v = (qkv.view(B, L, 3, nh, D)[:, :, i].permute(0, 2, 1, 3).contiguous().requires_grad_(True) for i in range(3))
    do = dctx.view(B, L, nh, D).permute(0, 2, 1, 3).contiguous()
    for name, be in (("torch SDPA cuDNN", SDPBackend.CUDNN_ATTENTION), ("torch SDPA FlashAttention-2", SDPBackend.FLASH_ATTENTION),
                     ("torch SDPA memory-efficient", SDPBackend.EFFICIENT_ATTENTION)):
        try:
            with sdpa_kernel([be]):
                f = timeit(lambda: F.scaled_dot_product_attention(q.detach(), k.detach(), v.detach()))
                out = F.scaled_dot_product_attention(q, k, v)
                b = timeit(lambda: torch.autograd.grad(out, (q, k, v), do, retain_graph=True))
            rows.append((name, f, b))
        except Exception as e:  # noqa: BLE001
            rows.append((name + f" (unavailable: {type(e).__name__})", float("nan"), float("nan")))
    try:
        from flash_attn import flash_attn_func
        q2, k2, v2 = (qkv.view(B, L, 3, nh, D)[:, :, i].contiguous().requires_grad_(True) for i in range(3))
        do2 = dctx.view(B, L, nh, D)
        f = timeit(lambda: flash_attn_func(q2.detach(), k2.detach(), v2.detach()))
        out = flash_attn_func(q2, k2, v2)
        b = timeit(lambda: torch.autograd.grad(out, (q2, k2, v2), do2, retain_graph=True))
        rows.append(("flash_attn package " + __import__("flash_attn").__version__, f, b))
    except Exception as e:  # noqa: BLE001
        rows.append((f"flash_attn package (unavailable: {type(e).__name__}: {e})"[:100], float("nan"), float("nan")))
    print(f"attention at {B} x {nh} heads x {L} x {D}, bf16; algorithmic FLOPs: forward {fwd_flops / 1e9:.1f} G, backward {2.5 * fwd_flops / 1e9:.1f} G")
    print("| implementation | forward us | TFLOP/s | backward us | TFLOP/s |")
    print("|---|---:|---:|---:|---:|")
    for name, f, b in rows:
        print(f"| {name} | {f:.0f} | {fwd_flops / f * 1e-6:.0f} | {b:.0f} | {2.5 * fwd_flops / b * 1e-6:.0f} |")


if __name__ == "__main__":
    main()
