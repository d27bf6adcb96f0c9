"""Times the HBM-bound row kernels at the BASELINE size (T = 98304 rows) and reports achieved GB/s (algorithmic bytes)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from plbert_b200._abi import call, load, ptr  # noqa: E402


def main():
    load()
    dev = torch.device("cuda:0")
    S = torch.cuda.current_stream().cuda_stream
    bf = torch.bfloat16
    T, H, I = 98304, 768, 2048
    z = torch.randn(T, H, device=dev).to(bf)
    dy = torch.randn(T, H, device=dev).to(bf)
    y = torch.empty_like(z)
    dz = torch.empty_like(z)
    stats = torch.empty(T, 2, device=dev)
    w, b = torch.randn(H, device=dev), torch.randn(H, device=dev)
    dw, db, cs = torch.zeros(H, device=dev), torch.zeros(H, device=dev), torch.zeros(H, device=dev)
    du = torch.randn(T, I, device=dev).to(bf)
    dqkv = torch.randn(T, 3 * H, device=dev).to(bf)
    csu, csq = torch.zeros(I, device=dev), torch.zeros(3 * H, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def stream_rate(fns, nbytes, label, iters=24):
        """Back-to-back launches rotating over independent buffer sets (each larger than L2), no flush in between: the
        regime the kernels see inside a training step."""
        for f in fns:
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fns[i % len(fns)]()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"{label:24s} {ms * 1e3:8.1f} us  {nbytes / ms * 1e-6:8.0f} GB/s   (streaming, {len(fns)} buffer sets)")

    def timeit(fn, nbytes, label, iters=10):
        for _ in range(3):
            fn()
        tot = 0.0
        for _ in range(iters):
            flush.zero_()  # evict L2 between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        ms = tot / iters
        print(f"{label:24s} {ms * 1e3:8.1f} us  {nbytes / ms * 1e-6:8.0f} GB/s")

    call("plb_ln_fwd", ptr(z), T, H, ptr(w), ptr(b), 1e-12, ptr(y), ptr(stats), S)
    timeit(lambda: call("plb_ln_fwd", ptr(z), T, H, ptr(w), ptr(b), 1e-12, ptr(y), ptr(stats), S), 2 * T * H * 2 + T * 8, "ln_fwd")
    timeit(lambda: call("plb_ln_bwd", ptr(dy), ptr(z), ptr(stats), ptr(w), T, H, ptr(dz), ptr(dw), ptr(db), ptr(cs), S),
           3 * T * H * 2 + T * 8, "ln_bwd")
    sets = [(torch.randn(T, H, device=dev).to(bf), torch.randn(T, H, device=dev).to(bf), torch.empty(T, H, device=dev, dtype=bf),
             torch.rand(T, 2, device=dev) + 0.5) for _ in range(3)]
    stream_rate([(lambda a=a, o=o, st=st: call("plb_ln_fwd", ptr(a), T, H, ptr(w), ptr(b), 1e-12, ptr(o), ptr(st), S))
                 for a, _, o, st in sets], 2 * T * H * 2 + T * 8, "ln_fwd")
    stream_rate([(lambda a=a, d=d, o=o, st=st: call("plb_ln_bwd", ptr(d), ptr(a), ptr(st), ptr(w), T, H, ptr(o), ptr(dw), ptr(db),
                                                    ptr(cs), S)) for a, d, o, st in sets], 3 * T * H * 2 + T * 8, "ln_bwd")
    timeit(lambda: call("plb_colsum", ptr(du), I, T, I, ptr(csu), S), T * I * 2, "colsum [T,2048]")
    timeit(lambda: call("plb_colsum", ptr(dqkv), 3 * H, T, 3 * H, ptr(csq), S), T * 3 * H * 2, "colsum [T,2304]")


if __name__ == "__main__":
    main()
