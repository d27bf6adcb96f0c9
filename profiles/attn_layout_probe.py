import os, sys
sys.path.insert(0, "/root/repo")
import torch
from plbert_b200._abi import call, load, ptr
load()
dev = torch.device("cuda:0")
S = torch.cuda.current_stream().cuda_stream
def run(B, L, nh, label):
    H, T = nh * 64, B * L
    qkv = torch.randn(T, 3 * H, device=dev).to(torch.bfloat16)
    lens = torch.full((B,), L, device=dev, dtype=torch.int32)
    ss = torch.arange(B, device=dev, dtype=torch.int32) * L
    ctx = torch.empty(T, H, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(nh, T, device=dev)
    f = lambda: call("plb_attn_fwd_tc", ptr(qkv), 3 * H, ptr(ss), ptr(lens), ptr(lens), B, nh, 64, L, T, ptr(ctx), H, ptr(lse), S)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    print(f"{label}: {e0.elapsed_time(e1)/10*1e3:.1f} us")
run(192, 512, 12, "B=192 heads=12 (row stride 4608 B)")
run(2304, 512, 1, "B=2304 heads=1 (row stride 384 B)")
run(192*4, 128, 12, "B=768 L=128 heads=12")
run(148, 512, 1, "B=148 heads=1: one CTA per SM")
run(148*2, 512, 1, "B=296 heads=1: two CTAs per SM")
