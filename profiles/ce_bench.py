"""Times the masked cross-entropy of the 64 k token head at the BASELINE size (M = 11 421 masked rows of 192x512, V = 64 000)
and reports achieved GB/s of algorithmic traffic (one read + one write of the bf16 logits).

    python profiles/ce_bench.py                      the product library
    python profiles/ce_bench.py build "CACHE=0 MINB=8" "CACHE=8 MINB=5" ...   (CPU container) -D variants of csrc/loss.cu
                                                     -> profiles/_variants/libce_v<n>.so (git-ignored)
    python profiles/ce_bench.py run N                (GPU box) times variants 0..N-1 and checks them against variant 0
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "profiles", "_variants")


def build(specs):
    from plbert_b200.build import BUILD, CSRC, NVCC_FLAGS, _nvcc, build as build_product
    build_product(selftest=False)
    os.makedirs(OUT, exist_ok=True)
    others = [os.path.join(BUILD, f) for f in os.listdir(BUILD) if f.endswith(".o") and f != "loss.o"]
    procs = []
    for n, spec in enumerate(specs):
        obj = os.path.join(OUT, f"loss_v{n}.o")
        defs = [f"-DPLB_CE_{kv}" for kv in spec.split()]
        procs.append((n, obj, subprocess.Popen([_nvcc(), *NVCC_FLAGS, *defs, "-c", os.path.join(CSRC, "loss.cu"), "-o", obj])))
    for n, obj, pr in procs:
        assert pr.wait() == 0, n
        subprocess.run([_nvcc(), "-shared", "-o", os.path.join(OUT, f"libce_v{n}.so"), obj, *others, "-gencode",
                        "arch=compute_100a,code=sm_100a", "-lcuda"], check=True)
        print("built variant", n, specs[n])


def bench(fn_of, names, shapes):
    import torch
    dev = torch.device("cuda:0")
    S = torch.cuda.current_stream().cuda_stream
    for M, V in shapes:
        ld = (V + 7) // 8 * 8
        src = (torch.randn(M, ld, device=dev) * 3).to(torch.bfloat16)
        z = src.clone()
        labels = torch.randint(0, V, (M,), device=dev, dtype=torch.int32)
        w = torch.rand(M, device=dev) / M
        row_loss = torch.empty(M, device=dev)
        loss = torch.zeros(1, device=dev)
        first = None
        for name in names:
            tot, iters = 0.0, 8
            for it in range(iters + 2):
                z.copy_(src)   # 1.46 GB rewritten: nothing of the row survives in L2
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn_of(name)(z.data_ptr(), ld, M, V, labels.data_ptr(), w.data_ptr(), row_loss.data_ptr(), loss.data_ptr(), 0, S)
                e1.record()
                torch.cuda.synchronize()
                if it >= 2:
                    tot += e0.elapsed_time(e1)
            ms = tot / iters
            nbytes = 2 * M * ld * 2
            note = ""
            if first is None:
                first = (z.clone(), row_loss.clone())
            else:
                dz = (z.float() - first[0].float()).norm() / first[0].float().norm()
                dl = (row_loss - first[1]).abs().max() / first[1].abs().max()
                note = f"  vs first: dz rel {dz.item():.2e}, row_loss rel {dl.item():.2e}"
            print(f"{name:>10s} masked_ce M={M} V={V}: {ms * 1e3:8.1f} us  {nbytes / ms * 1e-6:8.0f} GB/s (incl. the loss-sum launch){note}",
                  flush=True)


def main():
    P, I = ctypes.c_void_p, ctypes.c_int
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        return build(sys.argv[2:])
    if len(sys.argv) > 1 and sys.argv[1] == "run":
        libs = {}
        for n in range(int(sys.argv[2])):
            lib = ctypes.CDLL(os.path.join(OUT, f"libce_v{n}.so"))
            lib.plb_masked_ce.argtypes = [P, I, I, I, P, P, P, P, I, P]
            libs[f"v{n}"] = lib

        def fn_of(name):
            def f(*a):
                assert libs[name].plb_masked_ce(*a) == 0
            return f
        import torch  # noqa: F401
        return bench(fn_of, list(libs), [(11421, 64000)])
    from plbert_b200._abi import call, load
    load()
    bench(lambda name: (lambda *a: call("plb_masked_ce", *a)), ["product"],
          [(11421, 64000), (11421, 65536), (4096, 70000), (11421, 30000)])


if __name__ == "__main__":
    main()
