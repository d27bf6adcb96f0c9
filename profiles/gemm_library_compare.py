"""Context for the GEMM numbers: the layer's shapes (T = 98 304 rows) through cuBLAS / cuBLASLt (torch) next to the tcgen05 kernel
(C ABI), alone on the GPU, CUDA events, 10 launches after 3, operands rotated over 3 buffer sets (larger than L2)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from plbert_b200._abi import call, load, ptr  # noqa: E402


def timeit(fns, iters=12):
    for f in fns:
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fns[i % len(fns)]()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    load()
    dev = torch.device("cuda:0")
    bf = torch.bfloat16
    T = 98304
    S = torch.cuda.current_stream().cuda_stream
    print("| GEMM (M = 98 304 tokens) | this repo us | TFLOP/s | cuBLAS (torch) us | TFLOP/s |")
    print("|---|---:|---:|---:|---:|")
    for label, N, K in (("QKV forward  x[T,768] W[2304,768]^T + b", 2304, 768), ("O projection forward [T,768] x [768,768]^T + b", 768, 768),
                        ("FFN1 forward (plain epilogue) [T,768] x [2048,768]^T + b", 2048, 768),
                        ("FFN2 forward [T,2048] x [768,2048]^T + b", 768, 2048)):
        sets = [(torch.randn(T, K, device=dev).to(bf), torch.empty(T, N, device=dev, dtype=bf)) for _ in range(3)]
        W = (torch.randn(N, K, device=dev) * 0.02).to(bf)
        b32 = torch.randn(N, device=dev)
        bb = b32.to(bf)
        ours = timeit([(lambda A=A, C=C: call("plb_gemm_bf16", 0, 0, T, N, K, ptr(A), K, ptr(W), K, ptr(C), N, ptr(b32), 0, 0, 0, 0, S))
                       for A, C in sets])
        lib = timeit([(lambda A=A: F.linear(A, W, bb)) for A, _ in sets])
        fl = 2.0 * T * N * K
        print(f"| {label} | {ours:.0f} | {fl / ours * 1e-6:.0f} | {lib:.0f} | {fl / lib * 1e-6:.0f} |")
        del sets
    for label, Mo, No in (("WGRAD QKV  dy[T,2304]^T x[T,768] (fp32 accumulate)", 2304, 768), ("WGRAD FFN1 dy[T,2048]^T x[T,768]", 2048, 768),
                          ("WGRAD O    dy[T,768]^T x[T,768]", 768, 768)):
        sets = [(torch.randn(T, Mo, device=dev).to(bf), torch.randn(T, No, device=dev).to(bf)) for _ in range(3)]
        C = torch.zeros(Mo, No, device=dev)
        ours = timeit([(lambda A=A, B=B: call("plb_gemm_bf16", 2, 0, Mo, No, T, ptr(A), Mo, ptr(B), No, ptr(C), No, 0, 0, 0, 0, 0, S))
                       for A, B in sets])
        lib = timeit([(lambda A=A, B=B: torch.mm(A.t(), B)) for A, B in sets])
        fl = 2.0 * T * Mo * No
        print(f"| {label} | {ours:.0f} | {fl / ours * 1e-6:.0f} | {lib:.0f} (bf16 output, no accumulate) | {fl / lib * 1e-6:.0f} |")
        del sets


if __name__ == "__main__":
    main()
