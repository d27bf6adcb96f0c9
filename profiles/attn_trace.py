"""Timeline of one CTA of the tcgen05 attention kernels (clock64 stamps from lane 0 of every warp of CTA 0).

`python profiles/attn_trace.py build` compiles plbert_b200/csrc/*.cu with -DPLB_ATTN_TRACE into
profiles/_trace/libplbert_b200_trace.so (git-ignored; done in the CPU container, travels to the GPU box);
`python profiles/attn_trace.py run [fwd]` launches the kernel at the BASELINE shape and prints the records of CTA 0.
The product library never contains the instrumentation.
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "profiles", "_trace")
LIB = os.path.join(OUT, "libplbert_b200_trace.so")


def build():
    from plbert_b200.build import NVCC_FLAGS, _nvcc, _sources
    os.makedirs(OUT, exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, "-DPLB_ATTN_TRACE", f"-DPLB_TRACE_BLOCK={os.environ.get('TRACE_BLOCK', '0')}", "-shared", "-o", LIB, *_sources(), "-lcuda"]
    print(" ".join(cmd))
    subprocess.run(cmd, check=True)


def run(which="fwd"):
    import torch
    lib = ctypes.CDLL(LIB)
    dev = torch.device("cuda:0")
    B, L, nh = 192, 512, 12
    H, T = nh * 64, B * L
    P, I = ctypes.c_void_p, ctypes.c_int
    qkv = torch.randn(T, 3 * H, device=dev).to(torch.bfloat16)
    lens = torch.full((B,), L, device=dev, dtype=torch.int32)
    qlens = torch.full((B,), int(os.environ.get("QLEN", L)), device=dev, dtype=torch.int32)  # QLEN=128: one softmax group only
    seq_start = torch.arange(B, device=dev, dtype=torch.int32) * L
    ctx = torch.empty(T, H, device=dev, dtype=torch.bfloat16)
    dctx = torch.randn(T, H, device=dev).to(torch.bfloat16)
    dqkv = torch.empty(T, 3 * H, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(nh, T, device=dev)
    delta = torch.empty(nh, T, device=dev)
    dbias = torch.zeros(3 * H, device=dev)
    S = torch.cuda.current_stream().cuda_stream
    lib.plb_attn_fwd_tc.argtypes = [P, I, P, P, P, I, I, I, I, I, P, I, P, P]
    lib.plb_attn_bwd_tc.argtypes = [P, I, P, P, P, I, I, I, I, I, P, P, I, P, P, P, I, P, P]
    lib.plb_attn_trace_dump.argtypes = [P, I]

    def fwd():
        rc = lib.plb_attn_fwd_tc(qkv.data_ptr(), 3 * H, seq_start.data_ptr(), qlens.data_ptr(), lens.data_ptr(), B, nh, 64, L, T,
                                 ctx.data_ptr(), H, lse.data_ptr(), S)
        assert rc == 0

    def bwd():
        rc = lib.plb_attn_bwd_tc(qkv.data_ptr(), 3 * H, seq_start.data_ptr(), lens.data_ptr(), lens.data_ptr(), B, nh, 64, L, T,
                                 ctx.data_ptr(), dctx.data_ptr(), H, lse.data_ptr(), delta.data_ptr(), dqkv.data_ptr(), 3 * H,
                                 dbias.data_ptr(), S)
        assert rc == 0

    if which == "fused":
        return run_fused(lib, locals())
    cap = 3 * 32 * 512
    buf = (ctypes.c_ulonglong * cap)()
    lib.plb_attn_trace_dump(buf, cap)  # clears the table
    fwd()
    if which != "fwd":
        bwd()
    lib.plb_attn_trace_dump(buf, cap)  # warm-up records discarded
    (fwd if which == "fwd" else bwd)()
    n = lib.plb_attn_trace_dump(buf, cap)
    kid = {"fwd": 0, "dq": 1, "dkv": 2}[which]
    recs = sorted(((v >> 24), (v >> 16) & 31, (v >> 8) & 255, v & 255) for v in buf[:n] if v and ((v >> 21) & 7) == kid)
    t0 = recs[0][0]
    for clk, warp, code, idx in recs:
        print(f"{clk - t0:8d} warp{warp:2d} code {code:3d} idx {idx}")


def run_fused(lib, v):
    """Timeline of one CTA of the single-kernel backward (attn_bwd_fused.cu); prints the records and a per-phase summary."""
    import collections
    P, I = ctypes.c_void_p, ctypes.c_int
    B, nh, L, T, H, S = v["B"], v["nh"], v["L"], v["T"], v["H"], v["S"]
    lib.plb_attn_bwd_fused.argtypes = [P, I, P, P, P, I, I, I, I, I, P, P, I, P, P, P, I, P, P, ctypes.c_longlong, P]
    lib.plb_attn_fb_trace_dump.argtypes = [P, I]
    nbytes = lib.plb_attn_bwd_fused_workspace_bytes(B, nh)
    import torch
    ws = torch.empty(nbytes // 4, device="cuda")
    v["fwd"]()

    def bwd():
        rc = lib.plb_attn_bwd_fused(v["qkv"].data_ptr(), 3 * H, v["seq_start"].data_ptr(), v["lens"].data_ptr(), v["lens"].data_ptr(),
                                    B, nh, 64, L, T, 0, v["dctx"].data_ptr(), H, v["lse"].data_ptr(), v["delta"].data_ptr(),
                                    v["dqkv"].data_ptr(), 3 * H, v["dbias"].data_ptr(), ws.data_ptr(), nbytes, S)
        assert rc == 0

    cap = 16 * 1024
    buf = (ctypes.c_ulonglong * cap)()
    bwd()
    lib.plb_attn_fb_trace_dump(buf, cap)  # warm-up records discarded
    bwd()
    n = lib.plb_attn_fb_trace_dump(buf, cap)
    recs = sorted(((x >> 24), (x >> 16) & 31, (x >> 8) & 255, x & 255) for x in buf[:n] if x)
    t0 = recs[0][0]
    for clk, warp, code, idx in recs:
        print(f"{clk - t0:8d} warp{warp:2d} code {code:3d} idx {idx}")
    # per-phase summary over blocks 16..47 (items 1 and 2 of the CTA: steady state)
    at = {(w, c, i): clk - t0 for clk, w, c, i in recs}
    def span(w, c0, c1, di=0):
        d = [at[(w, c1, i + di)] - at[(w, c0, i)] for i in range(16, 48) if (w, c0, i) in at and (w, c1, i + di) in at]
        return (sum(d) / max(1, len(d)), min(d, default=0), max(d, default=0))
    print("# summary (clk: mean, min, max) over blocks 16..47")
    for label, w, c0, c1, di in [("block period (S/dP commit n -> n+1)", 1, 1, 1, 1), ("softmax: st_full -> loads done", 4, 10, 11, 0),
                                 ("softmax: loads done -> math done", 4, 11, 12, 0), ("softmax: math done -> dS stored, acc_done(n-1) seen", 4, 12, 13, 0),
                                 ("softmax: P^T store -> pds_ready", 4, 13, 14, 0), ("softmax: pds_ready -> next st_full", 4, 14, 10, 1),
                                 ("issuer B: pds_ready seen -> dq_free seen", 2, 2, 3, 0), ("issuer B: issue 24 MMAs + commit", 2, 3, 4, 0),
                                 ("issuer B: commit -> next pds_ready", 2, 4, 2, 1),
                                 ("epilogue: dq_full seen -> dq_free arrive", 12, 20, 21, 0), ("epilogue: dq_free -> dQ partial done", 12, 21, 22, 0),
                                 ("epilogue: dQ done -> next dq_full", 12, 22, 20, 1)]:
        m, lo, hi = span(w, c0, c1, di)
        print(f"#   {label:45s} {m:8.0f} {lo:8d} {hi:8d}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    else:
        run(sys.argv[2] if len(sys.argv) > 2 else "fwd")
