"""Context for the HBM-bound kernels: LayerNorm forward / backward, the 64 k head's cross-entropy (loss + logits gradient) and AdamW
through PyTorch's library kernels next to this repo's (C ABI) at the BASELINE sizes; CUDA events, inputs larger than L2."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from plbert_b200._abi import call, load, ptr  # noqa: E402


def timeit(fns, iters=12, pre=None):
    for f in fns:
        f()
    tot = 0.0
    for i in range(iters):
        if pre is not None:
            pre(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fns[i % len(fns)]()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


def main():
    load()
    dev = torch.device("cuda:0")
    bf = torch.bfloat16
    S = torch.cuda.current_stream().cuda_stream
    T, H = 98304, 768
    rows = []
    w, b = torch.randn(H, device=dev), torch.randn(H, device=dev)
    sets = [(torch.randn(T, H, device=dev).to(bf), torch.randn(T, H, device=dev).to(bf), torch.empty(T, H, device=dev, dtype=bf),
             torch.empty(T, 2, device=dev)) for _ in range(3)]
    dw, db, cs = (torch.zeros(H, device=dev) for _ in range(3))
    ours = timeit([(lambda z=z, y=y, st=st: call("plb_ln_fwd", ptr(z), T, H, ptr(w), ptr(b), 1e-12, ptr(y), ptr(st), S)) for z, _, y, st in sets])
    wb, bb_ = w.to(bf), b.to(bf)
    lib = timeit([(lambda z=z: F.layer_norm(z, (H,), wb, bb_, 1e-12)) for z, _, _, _ in sets])
    rows.append(("LayerNorm forward [98304, 768] bf16", ours, lib, 2 * T * H * 2))
    ours = timeit([(lambda z=z, d=d, o=o, st=st: call("plb_ln_bwd", ptr(d), ptr(z), ptr(st), ptr(w), T, H, ptr(o), ptr(dw), ptr(db), ptr(cs), S))
                   for z, d, o, st in sets])
    libfns = []
    for z, d, _, _ in sets:
        zz = z.clone().requires_grad_(True)
        wr, br = wb.clone().requires_grad_(True), bb_.clone().requires_grad_(True)
        out = F.layer_norm(zz, (H,), wr, br, 1e-12)
        libfns.append(lambda out=out, zz=zz, wr=wr, br=br, d=d: torch.autograd.grad(out, (zz, wr, br), d, retain_graph=True))
    lib = timeit(libfns)
    rows.append(("LayerNorm backward (dz, dgamma, dbeta; ours also the Linear's bias sums)", ours, lib, 3 * T * H * 2))
    del sets, libfns
    M, V = 11421, 64000
    src = (torch.randn(M, V, device=dev) * 3).to(bf)
    z = src.clone()
    labels = torch.randint(0, V, (M,), device=dev, dtype=torch.int32)
    wgt = torch.rand(M, device=dev) / M
    row_loss, loss = torch.empty(M, device=dev), torch.zeros(1, device=dev)
    ours = timeit([lambda: call("plb_masked_ce", ptr(z), V, M, V, ptr(labels), ptr(wgt), ptr(row_loss), ptr(loss), 0, S)],
                  pre=lambda i: z.copy_(src))
    zl = src.clone().requires_grad_(True)
    ll = labels.long()

    def lib_ce():
        l = (F.cross_entropy(zl, ll, reduction="none") * wgt).sum()
        torch.autograd.grad(l, zl)
    lib = timeit([lib_ce], iters=6)
    rows.append(("cross-entropy of the 64 k head, loss + logits gradient [11421, 64000] bf16", ours, lib, 2 * M * V * 2))
    del src, z, zl
    n = 55_600_000
    p = torch.randn(n, device=dev)
    g = torch.randn(n, device=dev)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    sh = torch.empty(n, device=dev, dtype=bf)
    step = [0]

    def ours_adam():
        step[0] += 1
        call("plb_adamw_step", ptr(p), ptr(g), ptr(m), ptr(v), ptr(sh), n, step[0], 7e-5, 0.9, 0.999, 1e-8, 0.01, 1.0, 1, S)
    ours = timeit([ours_adam])
    pl = torch.nn.Parameter(torch.randn(n, device=dev))
    pl.grad = torch.randn(n, device=dev)
    opt = torch.optim.AdamW([pl], lr=7e-5, fused=True)
    lib = timeit([lambda: opt.step()])
    rows.append(("AdamW, 55.6 M fp32 parameters (ours also writes the bf16 shadow and zeroes the gradient)", ours, lib, 34 * n))
    print("| kernel | this repo us | GB/s (algorithmic) | PyTorch library us |")
    print("|---|---:|---:|---:|")
    for name, o, l, nbytes in rows:
        print(f"| {name} | {o:.0f} | {nbytes / o * 1e-3:.0f} | {l:.0f} |")


if __name__ == "__main__":
    main()
