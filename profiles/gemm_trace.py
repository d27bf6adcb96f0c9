"""clock64 timeline of ONE epilogue warp of the tcgen05 GEMM (CTA 0, warp 2) for one fused-epilogue variant.

Needs the trace build of the library (never the product one):
    nvcc <flags of plbert_b200/build.py> -DPLB_GEMM_TRACE -c plbert_b200/csrc/gemm_tcgen05.cu ... -> profiles/_trace/libplbert_b200_gemmtrace.so
    PLBERT_B200_LIB=$PWD/profiles/_trace/libplbert_b200_gemmtrace.so python profiles/gemm_trace.py
Tags: 1 tile start, 2 accumulator ready (tmem_full), per 64-column chunk: 3 staging slot / aux operand ready, 4 accumulator
chunk in registers (tcgen05.wait::ld), 5 math done + staged, 6 fence.proxy.async done; 7 tile handed back to the MMA warp.
"""
import ctypes
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from plbert_b200 import _abi  # noqa: E402
from plbert_b200._abi import call, ptr  # noqa: E402

NAMES = {(1, 2): "wait for the accumulator (bias staging + tmem_full)", (2, 3): "wait for the staging slot / aux operand",
         (3, 4): "tcgen05.ld + wait", (4, 5): "epilogue math + st.shared", (5, 6): "fence.proxy.async",
         (6, 3): "arrive + loop to the next chunk", (6, 7): "arrive + tmem_empty hand-back", (7, 1): "next unit's coordinates"}


def main():
    lib = _abi.load()
    if not hasattr(lib, "plb_debug_gemm_trace"):
        raise SystemExit("this is not the trace build: set PLBERT_B200_LIB to profiles/_trace/libplbert_b200_gemmtrace.so")
    dev = torch.device("cuda:0")
    S = torch.cuda.current_stream().cuda_stream
    bf = torch.bfloat16
    T = 98304
    cases = [("fwd qkv (bias)", 0, 0, 2304, 768, True), ("fwd o (bias+residual)", 0, 1, 768, 768, True),
             ("fwd ffn1 (bias+gelu, dual store)", 0, 2, 2048, 768, True), ("dgrad ffn2 (x gelu')", 1, 3, 2048, 768, False),
             ("fwd ffn1 WITHOUT bias (how much of the chunk time is the bias LDS?)", 0, 2, 2048, 768, False)]
    for label, op, epi, N, K, with_bias in cases:
        A = torch.randn(T, K, device=dev).to(bf)
        W = (torch.randn(N, K, device=dev) if op == 0 else torch.randn(K, N, device=dev)).to(bf)
        C = torch.empty(T, N, device=dev, dtype=bf)
        X = torch.randn(T, N, device=dev).to(bf)
        U = torch.empty(T, N, device=dev, dtype=bf)
        bias = torch.randn(N, device=dev) if with_bias else None
        for _ in range(2):
            lib.plb_debug_gemm_trace_clear()
            call("plb_gemm_bf16", op, epi, T, N, K, ptr(A), K, ptr(W), W.stride(0), ptr(C), N, ptr(bias),
                 ptr(X) if epi in (1, 3) else 0, N, ptr(U) if epi == 2 else 0, N, S)
            torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * 8192)()
        lib.plb_debug_gemm_trace(buf, 8192)
        ev = [(v >> 56, v & ((1 << 56) - 1)) for v in buf if v]
        # skip the first two tiles (pipeline fill), aggregate the deltas between consecutive tags
        starts = [i for i, (t, _) in enumerate(ev) if t == 1]
        ev = ev[starts[2]:] if len(starts) > 3 else ev
        agg = defaultdict(lambda: [0, 0])
        for (t0, c0), (t1, c1) in zip(ev, ev[1:]):
            agg[(t0, t1)][0] += 1
            agg[(t0, t1)][1] += c1 - c0
        tiles = sum(1 for t, _ in ev if t == 7)
        total = ev[-1][1] - ev[0][1]
        print(f"\n{label}: {tiles} tiles of CTA 0, {total / max(tiles, 1):.0f} clk per tile (one epilogue warp)")
        for key, (n, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"  {NAMES.get(key, str(key)):55s} {tot / max(tiles, 1):8.0f} clk/tile  ({n / max(tiles, 1):.1f} x {tot / n:6.0f})")


if __name__ == "__main__":
    main()
