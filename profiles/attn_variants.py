"""A/B of attention kernel variants: -DPLB_FWD_X=n builds of csrc/attn_tc.cu or -DPLB_BWD_X=n builds of csrc/attn_bwd_fused.cu,
linked against the product objects.

`python profiles/attn_variants.py build fwd|bwd 0 1 2 ...` (CPU container) -> profiles/_variants/lib<kind>_v<n>.so (git-ignored);
`python profiles/attn_variants.py run fwd|bwd 0 1 2 ...` (GPU box): times the kernel at the BASELINE shape per variant (CUDA events,
interleaved rounds) and checks every variant's outputs against the first one listed.  VARLEN=seed: lengths U[32, 512], packed rows.
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "profiles", "_variants")
SRC = {"fwd": ("attn_tc", "PLB_FWD_X"), "bwd": ("attn_bwd_fused", "PLB_BWD_X")}


def build(kind, variants):
    from plbert_b200.build import BUILD, CSRC, NVCC_FLAGS, _nvcc, build as build_product
    build_product(selftest=False)
    os.makedirs(OUT, exist_ok=True)
    stem, macro = SRC[kind]
    others = [os.path.join(BUILD, f) for f in os.listdir(BUILD) if f.endswith(".o") and f != stem + ".o"]
    procs = []
    for v in variants:
        obj = os.path.join(OUT, f"{stem}_v{v}.o")
        procs.append((v, obj, subprocess.Popen([_nvcc(), *NVCC_FLAGS, f"-D{macro}={v}", "-c", os.path.join(CSRC, stem + ".cu"), "-o", obj])))
    for v, obj, pr in procs:
        assert pr.wait() == 0, v
        subprocess.run([_nvcc(), "-shared", "-o", os.path.join(OUT, f"lib{kind}_v{v}.so"), obj, *others, "-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"], check=True)
        print("built", kind, "variant", v)


def run(kind, variants):
    import torch
    dev = torch.device("cuda:0")
    B, L, nh = 192, 512, 12
    H = nh * 64
    P, I = ctypes.c_void_p, ctypes.c_int
    torch.manual_seed(0)
    if os.environ.get("VARLEN"):
        g = torch.Generator().manual_seed(int(os.environ["VARLEN"]))
        lens_l = sorted(torch.randint(32, 513, (B,), generator=g).tolist(), reverse=True)
    else:
        lens_l = [L] * B
    T = sum(lens_l)
    lens = torch.tensor(lens_l, device=dev, dtype=torch.int32)
    seq_start = torch.tensor([0] + torch.tensor(lens_l).cumsum(0).tolist()[:-1], device=dev, dtype=torch.int32)
    qkv = torch.randn(T, 3 * H, device=dev).to(torch.bfloat16)
    dctx = torch.randn(T, H, device=dev).to(torch.bfloat16)
    S = torch.cuda.current_stream().cuda_stream
    flops = 4.0 * nh * sum(n * n for n in lens_l) * 64 * (1.0 if kind == "fwd" else 2.5)
    libs, outs = {}, {}
    for v in variants:
        lib = ctypes.CDLL(os.path.join(OUT, f"lib{kind}_v{v}.so"))
        lib.plb_attn_fwd_tc.argtypes = [P, I, P, P, P, I, I, I, I, I, P, I, P, P]
        lib.plb_attn_bwd_fused.argtypes = [P, I, P, P, P, I, I, I, I, I, P, P, I, P, P, P, I, P, P, ctypes.c_longlong, P]
        libs[v] = lib
        ctx, lse = torch.empty(T, H, device=dev, dtype=torch.bfloat16), torch.empty(nh, T, device=dev)
        outs[v] = dict(ctx=ctx, lse=lse)
        if kind == "bwd":
            nbytes = lib.plb_attn_bwd_fused_workspace_bytes(B, nh)
            outs[v].update(dqkv=torch.empty(T, 3 * H, device=dev, dtype=torch.bfloat16), dbias=torch.zeros(3 * H, device=dev),
                           delta=torch.empty(nh, T, device=dev), ws=torch.empty(nbytes // 4, device=dev), nbytes=nbytes)

    def fwd(v):
        o = outs[v]
        rc = libs[v].plb_attn_fwd_tc(qkv.data_ptr(), 3 * H, seq_start.data_ptr(), lens.data_ptr(), lens.data_ptr(), B, nh, 64, L, T,
                                     o["ctx"].data_ptr(), H, o["lse"].data_ptr(), S)
        assert rc == 0, v

    def bwd(v, with_ctx=0):
        o = outs[v]
        rc = libs[v].plb_attn_bwd_fused(qkv.data_ptr(), 3 * H, seq_start.data_ptr(), lens.data_ptr(), lens.data_ptr(), B, nh, 64, L, T,
                                        o["ctx"].data_ptr() if with_ctx else 0, dctx.data_ptr(), H, o["lse"].data_ptr(), o["delta"].data_ptr(),
                                        o["dqkv"].data_ptr(), 3 * H, o["dbias"].data_ptr(), o["ws"].data_ptr(), o["nbytes"], S)
        assert rc == 0, v

    if kind == "bwd":
        for v in variants:
            fwd(v)
            bwd(v, 1)   # computes delta from ctx once; the timed calls take it as given (the step's configuration)
    fn = fwd if kind == "fwd" else bwd
    times = {v: [] for v in variants}
    for rnd in range(4):
        for v in variants:
            for _ in range(3):
                fn(v)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn(v)
            e1.record()
            torch.cuda.synchronize()
            times[v].append(e0.elapsed_time(e1) / 10 * 1e3)
    if kind == "bwd":   # one clean call for the comparison of the accumulated bias sums
        for v in variants:
            outs[v]["dbias"].zero_()
            bwd(v)
        torch.cuda.synchronize()
    ref = outs[variants[0]]
    keys = ("ctx", "lse") if kind == "fwd" else ("dqkv", "dbias")
    for v in variants:
        t = sorted(times[v])
        diffs = "  ".join(f"max|{k} diff| {(outs[v][k].float() - ref[k].float()).abs().max().item():.3e}" for k in keys)
        print(f"{kind} variant {v:3d}: {t[0]:7.1f} / {t[len(t) // 2]:7.1f} us (best / median of 4 rounds)  {flops / t[0] * 1e-6:7.1f} TFLOP/s   {diffs}")


if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1]](sys.argv[2], [int(a) for a in sys.argv[3:]] or [0])
