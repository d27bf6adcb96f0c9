"""Small driver for `ncu --set full`: one launch each of the hot kernels at sizes that fill the chip for a few waves."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from plbert_b200._abi import call, load, ptr  # noqa: E402


def main():
    which = sys.argv[1:] or ["attn", "gemm"]
    load()
    dev = torch.device("cuda:0")
    S = torch.cuda.current_stream().cuda_stream
    bf = torch.bfloat16
    if "attn" in which:
        B, L, nh = 37, 512, 12  # 444 CTAs = 3 waves
        H, T = nh * 64, B * L
        qkv = torch.randn(T, 3 * H, device=dev).to(bf)
        lens = torch.full((B,), L, device=dev, dtype=torch.int32)
        ss = torch.arange(B, device=dev, dtype=torch.int32) * L
        ctx = torch.empty(T, H, device=dev, dtype=bf)
        dctx = torch.randn(T, H, device=dev).to(bf)
        dqkv = torch.empty(T, 3 * H, device=dev, dtype=bf)
        lse, delta = torch.empty(nh, T, device=dev), torch.empty(nh, T, device=dev)
        for _ in range(2):
            call("plb_attn_fwd_tc", ptr(qkv), 3 * H, ptr(ss), ptr(lens), ptr(lens), B, nh, 64, L, T, ptr(ctx), H, ptr(lse), S)
            if hasattr(load(), "plb_attn_bwd_tc"):
                call("plb_attn_bwd_tc", ptr(qkv), 3 * H, ptr(ss), ptr(lens), ptr(lens), B, nh, 64, L, T, ptr(ctx), ptr(dctx), H,
                     ptr(lse), ptr(delta), ptr(dqkv), 3 * H, 0, S)
    if "gemm" in which:
        T = 148 * 128 * 4
        for (op, epi, N, K) in ((0, 0, 2304, 768), (0, 0, 768, 128), (0, 1, 768, 768), (0, 2, 2048, 768), (1, 3, 2048, 768), (2, 0, 768, 0)):
            if op == 2:
                A = torch.randn(T, 2304, device=dev).to(bf)
                Bm = torch.randn(T, 768, device=dev).to(bf)
                C = torch.zeros(2304, 768, device=dev)
                for _ in range(2):
                    call("plb_gemm_bf16", 2, 0, 2304, 768, T, ptr(A), 2304, ptr(Bm), 768, ptr(C), 768, 0, 0, 0, 0, 0, S)
                continue
            A = torch.randn(T, K, device=dev).to(bf)
            W = torch.randn(N, K, device=dev).to(bf) if op == 0 else torch.randn(K, N, device=dev).to(bf)
            C = torch.empty(T, N, device=dev, dtype=bf)
            X = torch.randn(T, N, device=dev).to(bf)
            U = torch.empty(T, N, device=dev, dtype=bf)
            bias = torch.randn(N, device=dev) if op == 0 else None
            for _ in range(2):
                call("plb_gemm_bf16", op, epi, T, N, K, ptr(A), K, ptr(W), W.stride(0), ptr(C), N, ptr(bias),
                     ptr(X) if epi in (1, 3) else 0, N, ptr(U) if epi == 2 else 0, N, S)
    torch.cuda.synchronize()
    print("probe done")


if __name__ == "__main__":
    main()
