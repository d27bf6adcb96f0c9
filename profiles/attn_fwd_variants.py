"""A/B of forward-attention kernel variants (-DPLB_FWD_X=n builds of csrc/attn_tc.cu linked against the product objects).

`python profiles/attn_fwd_variants.py build 0 1 2 ...` (CPU container) -> profiles/_variants/libv<n>.so (git-ignored);
`python profiles/attn_fwd_variants.py run 0 1 2 ...` (GPU box): times plb_attn_fwd_tc at the BASELINE shape per variant (CUDA events,
interleaved rounds) and checks every variant's output against variant 0.
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "profiles", "_variants")


def build(variants):
    from plbert_b200.build import BUILD, CSRC, NVCC_FLAGS, _nvcc, build as build_product
    build_product(selftest=False)
    os.makedirs(OUT, exist_ok=True)
    others = [os.path.join(BUILD, f) for f in os.listdir(BUILD) if f.endswith(".o") and f != "attn_tc.o"]
    procs = []
    for v in variants:
        obj = os.path.join(OUT, f"attn_tc_v{v}.o")
        procs.append((v, obj, subprocess.Popen([_nvcc(), *NVCC_FLAGS, f"-DPLB_FWD_X={v}", "-c", os.path.join(CSRC, "attn_tc.cu"), "-o", obj])))
    for v, obj, pr in procs:
        assert pr.wait() == 0, v
        subprocess.run([_nvcc(), "-shared", "-o", os.path.join(OUT, f"libv{v}.so"), obj, *others, "-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"], check=True)
        print("built variant", v)


def run(variants):
    import torch
    dev = torch.device("cuda:0")
    B, L, nh = 192, 512, 12
    H, T = nh * 64, B * L
    P, I = ctypes.c_void_p, ctypes.c_int
    torch.manual_seed(0)
    qkv = torch.randn(T, 3 * H, device=dev).to(torch.bfloat16)
    lens = torch.full((B,), L, device=dev, dtype=torch.int32)
    seq_start = torch.arange(B, device=dev, dtype=torch.int32) * L
    S = torch.cuda.current_stream().cuda_stream
    flops = 4.0 * nh * B * L * L * 64
    libs, outs = {}, {}
    for v in variants:
        lib = ctypes.CDLL(os.path.join(OUT, f"libv{v}.so"))
        lib.plb_attn_fwd_tc.argtypes = [P, I, P, P, P, I, I, I, I, I, P, I, P, P]
        libs[v] = lib
        outs[v] = (torch.empty(T, H, device=dev, dtype=torch.bfloat16), torch.empty(nh, T, device=dev))

    def fwd(v):
        ctx, lse = outs[v]
        rc = libs[v].plb_attn_fwd_tc(qkv.data_ptr(), 3 * H, seq_start.data_ptr(), lens.data_ptr(), lens.data_ptr(), B, nh, 64, L, T,
                                     ctx.data_ptr(), H, lse.data_ptr(), S)
        assert rc == 0, v

    times = {v: [] for v in variants}
    for rnd in range(4):
        for v in variants:
            for _ in range(3):
                fwd(v)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fwd(v)
            e1.record()
            torch.cuda.synchronize()
            times[v].append(e0.elapsed_time(e1) / 10 * 1e3)
    ref_ctx, ref_lse = outs[variants[0]]
    for v in variants:
        ctx, lse = outs[v]
        t = sorted(times[v])
        print(f"variant {v:3d}: {t[0]:7.1f} / {t[len(t) // 2]:7.1f} us (best / median of 4 rounds)  {flops / t[0] * 1e-6:7.1f} TFLOP/s   "
              f"max|ctx - v{variants[0]}| {(ctx.float() - ref_ctx.float()).abs().max().item():.3e}  max|lse diff| {(lse - ref_lse).abs().max().item():.3e}")


if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1]]([int(a) for a in sys.argv[2:]] or [0])
