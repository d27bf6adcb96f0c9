"""Runs the attention forward + the single-kernel backward a few times at the BASELINE shape (for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from plbert_b200._abi import call, load, ptr  # noqa: E402


def main(reps=3):
    load()
    dev = torch.device("cuda:0")
    B, L, nh = int(os.environ.get("B", 192)), 512, 12
    H, T = nh * 64, B * L
    qkv = torch.randn(T, 3 * H, device=dev).to(torch.bfloat16)
    lens = torch.full((B,), L, device=dev, dtype=torch.int32)
    seq_start = torch.arange(B, device=dev, dtype=torch.int32) * L
    ctx = torch.empty(T, H, device=dev, dtype=torch.bfloat16)
    dctx = torch.randn(T, H, device=dev).to(torch.bfloat16)
    dqkv = torch.empty(T, 3 * H, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(nh, T, device=dev)
    delta = torch.empty(nh, T, device=dev)
    dbias = torch.zeros(3 * H, device=dev)
    S = torch.cuda.current_stream().cuda_stream
    nbytes = load().plb_attn_bwd_fused_workspace_bytes(B, nh)
    ws = torch.empty(nbytes // 4, device=dev)
    call("plb_attn_fwd_tc", ptr(qkv), 3 * H, ptr(seq_start), ptr(lens), ptr(lens), B, nh, 64, L, T, ptr(ctx), H, ptr(lse), S)
    call("plb_attn_bwd_tc", ptr(qkv), 3 * H, ptr(seq_start), ptr(lens), ptr(lens), B, nh, 64, L, T, ptr(ctx), ptr(dctx), H,
         ptr(lse), ptr(delta), ptr(dqkv), 3 * H, 0, S)   # fills delta
    for _ in range(reps):
        call("plb_attn_bwd_fused", ptr(qkv), 3 * H, ptr(seq_start), ptr(lens), ptr(lens), B, nh, 64, L, T, 0, ptr(dctx), H,
             ptr(lse), ptr(delta), ptr(dqkv), 3 * H, ptr(dbias), ptr(ws), nbytes, S)
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()
