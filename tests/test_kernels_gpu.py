"""Per-kernel parity on the device, through the C ABI (ctypes), against plain PyTorch fp32 references of the same
op on the same bf16-rounded inputs.  Integer / index work is compared bit-exactly; floating point with the
tolerance stated at each assert (bf16 outputs carry 2^-8 relative rounding)."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from plbert_b200 import _abi  # noqa: E402
from plbert_b200._abi import call, ptr  # noqa: E402

BF = torch.bfloat16


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _abi.load()
    torch.manual_seed(0)
    return torch.device("cuda:0")


def S():
    return torch.cuda.current_stream().cuda_stream


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def gelu_new(u):
    return 0.5 * u * (1 + torch.tanh(math.sqrt(2 / math.pi) * (u + 0.044715 * u ** 3)))


# ------------------------------------------------------------------------------------------ a1
def test_length_to_mask_golden_bit_exact(dev, golden_dir):
    z = np.load(os.path.join(golden_dir, "length_to_mask.npz"))
    i = 0
    while f"lengths_{i}" in z.files:
        lens = torch.tensor(z[f"lengths_{i}"], dtype=torch.int32, device=dev)
        B, L = lens.numel(), int(lens.max())
        out = torch.empty((B, L), dtype=torch.int32, device=dev)
        call("plb_length_to_mask", ptr(lens), B, L, ptr(out), S())
        assert np.array_equal(out.cpu().numpy(), z[f"attn_mask_{i}"])
        from plbert_b200.train import length_to_mask
        assert np.array_equal(length_to_mask(lens).cpu().numpy(), z[f"pad_mask_{i}"])
        assert np.array_equal(length_to_mask(lens.cpu()).numpy(), z[f"pad_mask_{i}"])
        i += 1
    assert i >= 4


# ------------------------------------------------------------------------------------------ GEMM
def _gemm(op, epi, M, N, K, A, B, C, bias=None, aux_in=None, aux_out=None, fn="plb_gemm_bf16"):
    call(fn, op, epi, M, N, K, ptr(A), A.stride(0), ptr(B), B.stride(0), ptr(C), C.stride(0), ptr(bias), ptr(aux_in),
         0 if aux_in is None else aux_in.stride(0), ptr(aux_out), 0 if aux_out is None else aux_out.stride(0), S())


# (30000, 768, 192): many short-K tiles per CTA — the epilogue runs back to back, which is what exposes ordering bugs between
# its warps (a bias-staging race was caught by exactly this shape)
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1000, 188, 768), (4096, 2304, 768), (777, 768, 2048), (300, 768, 128),
                                   (30000, 768, 192)])
@pytest.mark.parametrize("epi", [_abi.EPI_NONE, _abi.EPI_RESIDUAL, _abi.EPI_GELU])
def test_gemm_fwd(dev, M, N, K, epi):
    ldn = (N + 7) // 8 * 8
    A = torch.randn(M, K, device=dev).to(BF)
    W = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(BF)
    bias = torch.randn(N, device=dev)
    R = torch.randn(M, ldn, device=dev).to(BF)
    C = torch.zeros(M, ldn, device=dev, dtype=BF)
    U = torch.zeros(M, ldn, device=dev, dtype=BF)
    _gemm(_abi.GEMM_FWD, epi, M, N, K, A, W, C, bias, R if epi == _abi.EPI_RESIDUAL else None, U if epi == _abi.EPI_GELU else None)
    ref = A.float() @ W.float().T + bias
    if epi == _abi.EPI_RESIDUAL:
        ref = ref + R[:, :N].float()
    if epi == _abi.EPI_GELU:
        # the second output is gelu_new'(u) on the bf16-rounded pre-activation (what the backward epilogue multiplies by)
        u = ref.to(BF).float().requires_grad_(True)
        (gp,) = torch.autograd.grad(gelu_new(u).sum(), u)
        assert rel(U[:, :N].float(), gp) < 6e-3
        assert (U[:, :N].float() - gp).abs().max().item() <= 0.02
        ref = gelu_new(ref.to(BF).float())
    assert rel(C[:, :N].float(), ref) < 6e-3
    assert torch.isfinite(C.float()).all()
    # element-wise as well: a race that corrupts a few hundred elements does not move the global norm
    assert (C[:, :N].float() - ref).abs().max().item() <= 0.02 * ref.abs().max().item() + 0.05


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (999, 768, 188), (4096, 768, 2304), (2048, 2048, 768), (500, 128, 768)])
@pytest.mark.parametrize("epi", [_abi.EPI_NONE, _abi.EPI_RESIDUAL, _abi.EPI_GELU_BWD])
def test_gemm_dgrad(dev, M, N, K, epi):
    ldk, ldn = (K + 7) // 8 * 8, (N + 7) // 8 * 8
    A = torch.zeros(M, ldk, device=dev, dtype=BF)
    A[:, :K] = torch.randn(M, K, device=dev).to(BF)
    W = torch.zeros(K, ldn, device=dev, dtype=BF)
    W[:, :N] = (torch.randn(K, N, device=dev) / math.sqrt(K)).to(BF)
    X = torch.randn(M, ldn, device=dev).to(BF)
    C = torch.zeros(M, ldn, device=dev, dtype=BF)
    _gemm(_abi.GEMM_DGRAD, epi, M, N, K, A, W, C, None, X if epi != _abi.EPI_NONE else None)
    ref = A[:, :K].float() @ W[:, :N].float()
    if epi == _abi.EPI_RESIDUAL:
        ref = ref + X[:, :N].float()
    if epi == _abi.EPI_GELU_BWD:
        ref = ref * X[:, :N].float()   # aux_in = gelu'(u) as saved by the forward epilogue: a plain multiply
    assert rel(C[:, :N].float(), ref) < 6e-3
    assert (C[:, :N].float() - ref).abs().max().item() <= 0.02 * ref.abs().max().item() + 0.05


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (999, 768, 188), (4097, 2048, 768), (30000, 2048, 192), (300, 100, 768)])
def test_gemm_gelu_bwd_fused_column_sums(dev, M, N, K):
    """PLB_EPI_GELU_BWD with an fp32 aux_out: the column sums of the STORED d_u (the FFN1 bias gradient) are added to the
    accumulator by the GEMM's store warp; C itself is unchanged.  Ragged M / N, many tiles per CTA, repeated accumulation."""
    ldn = (N + 7) // 8 * 8
    A = torch.randn(M, K + (-K) % 8, device=dev).to(BF)
    A[:, K:] = 0
    W = torch.zeros(K, ldn, device=dev, dtype=BF)
    W[:, :N] = (torch.randn(K, N, device=dev) / math.sqrt(K)).to(BF)
    X = torch.randn(M, ldn, device=dev).to(BF)
    C0 = torch.zeros(M, ldn, device=dev, dtype=BF)
    _gemm(_abi.GEMM_DGRAD, _abi.EPI_GELU_BWD, M, N, K, A, W, C0, None, X)
    C = torch.zeros(M, ldn, device=dev, dtype=BF)
    acc = torch.ones(N, device=dev)
    for _ in range(2):
        _gemm(_abi.GEMM_DGRAD, _abi.EPI_GELU_BWD, M, N, K, A, W, C, None, X, acc)
    assert torch.equal(C, C0)
    want = C[:, :N].float().sum(0)
    assert (acc - 1 - 2 * want).abs().max().item() <= 2e-3 * want.abs().max().item() + 1e-3


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (1000, 768, 768), (4097, 256, 192), (77, 128, 2048)])
def test_gemm_rowdot_epilogue(dev, M, N, K):
    """PLB_EPI_ROWDOT: C = A W and, per 64-column block (= attention head), sum_n bf16(C[m,n]) * aux_in[m,n] accumulated
    into an fp32 [N/64, M] array — attention's delta = rowsum(dO o O) fused into the GEMM that produces dO."""
    A = torch.randn(M, K, device=dev).to(BF)
    W = (torch.randn(K, N, device=dev) / math.sqrt(K)).to(BF)
    O = torch.randn(M, N, device=dev).to(BF)
    C = torch.full((M, N), float("nan"), device=dev, dtype=BF)
    delta = torch.zeros(N // 64, M + 5, device=dev)          # ld_aux_out > M is allowed
    _gemm(_abi.GEMM_DGRAD, _abi.EPI_ROWDOT, M, N, K, A, W, C, None, O, delta)
    assert rel(C.float(), A.float() @ W.float()) < 6e-3
    want = (C.float() * O.float()).view(M, N // 64, 64).sum(-1).T      # over the kernel's own ROUNDED outputs
    assert (delta[:, :M] - want).abs().max() <= 1e-5 * want.abs().max() + 1e-6
    assert not delta[:, M:].any()
    # accumulates: a second launch doubles it
    _gemm(_abi.GEMM_DGRAD, _abi.EPI_ROWDOT, M, N, K, A, W, C, None, O, delta)
    assert (delta[:, :M] - 2 * want).abs().max() <= 2e-5 * want.abs().max() + 1e-6
    with pytest.raises(_abi.PlbError):
        _gemm(_abi.GEMM_FWD, _abi.EPI_ROWDOT, M, N, K, A, W.T.contiguous(), C, None, O, delta)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (188, 768, 1000), (2304, 768, 20000), (768, 2048, 4096), (768, 128, 3000)])
def test_gemm_wgrad_accumulates(dev, M, N, K):
    ldm = (M + 7) // 8 * 8
    A = torch.zeros(K, ldm, device=dev, dtype=BF)
    A[:, :M] = torch.randn(K, M, device=dev).to(BF)
    B = torch.randn(K, N, device=dev).to(BF)
    C = torch.randn(M, N, device=dev)
    ref = C.double() + A[:, :M].double().T @ B.double()
    _gemm(_abi.GEMM_WGRAD, _abi.EPI_NONE, M, N, K, A, B, C)
    assert rel(C, ref) < 1e-5  # fp32 accumulate, only the summation order differs
    _gemm(_abi.GEMM_WGRAD, _abi.EPI_NONE, M, N, K, A, B, C)  # second application sums in place (shared layer)
    assert rel(C, ref + A[:, :M].double().T @ B.double()) < 1e-5


def test_native_gemm_selftest_binary(dev):
    """The C++ self-test (tests/native/gemm_selftest.cu, built by plbert_b200.build) checks every GEMM variant against a
    plain CUDA-core GEMM, element by element, without Python in the loop; run it three times (races are timing dependent)."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native", "gemm_selftest")
    if not os.path.exists(exe):
        pytest.skip("native self-test not built")
    for _ in range(3):
        r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "SELFTEST PASSED" in r.stdout, r.stdout[-2000:]


def test_gemm_linearity_at_full_size(dev):
    """Size-independent property at the BASELINE size (T = 192 x 512): (A1 + A2) W^T == A1 W^T + A2 W^T for
    operands chosen so that every product and sum is exact in bf16/fp32."""
    T, N, K = 98304, 768, 768
    A1 = torch.randint(-2, 3, (T, K), device=dev).to(BF)
    A2 = torch.randint(-2, 3, (T, K), device=dev).to(BF)
    W = torch.randint(-1, 2, (N, K), device=dev).to(BF)
    outs = []
    for A in (A1, A2, (A1.float() + A2.float()).to(BF)):
        C = torch.empty(T, N, device=dev, dtype=BF)
        _gemm(_abi.GEMM_FWD, _abi.EPI_NONE, T, N, K, A, W, C)
        outs.append(C.float())
    # |sum| <= 4*768 = 3072 needs 12 bits > bf16's 8: compare in the exactly-representable subset via rounding bound
    assert (outs[2] - (outs[0] + outs[1])).abs().max() <= 2 ** -7 * (outs[0] + outs[1]).abs().max()
    idx = torch.randint(0, T, (64,), device=dev)
    ref = (A1[idx].float() + A2[idx].float()) @ W.float().T
    assert (outs[2][idx] - ref).abs().max() <= 2 ** -8 * ref.abs().max() + 1e-6


def test_gemm_rejects_bad_arguments(dev):
    A = torch.zeros(16, 16, device=dev, dtype=BF)
    with pytest.raises(_abi.PlbError):
        call("plb_gemm_bf16", 0, 0, 16, 16, 16, ptr(A), 12, ptr(A), 16, ptr(A), 16, 0, 0, 0, 0, 0, S())  # lda % 8
    with pytest.raises(_abi.PlbError):
        call("plb_gemm_bf16", 0, 0, 0, 16, 16, ptr(A), 16, ptr(A), 16, ptr(A), 16, 0, 0, 0, 0, 0, S())  # M == 0
    with pytest.raises(_abi.PlbError):
        call("plb_gemm_bf16", 2, 1, 16, 16, 16, ptr(A), 16, ptr(A), 16, ptr(A), 16, 0, 0, 0, 0, 0, S())  # wgrad + epilogue


# ------------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("T,H", [(1000, 768), (999, 768), (513, 128), (64, 1024), (2, 256), (7, 8)])
def test_layernorm_fwd_bwd(dev, T, H):
    z = (torch.randn(T, H, device=dev) * 2 + 0.3).to(BF)
    w = torch.randn(H, device=dev)
    b = torch.randn(H, device=dev)
    y = torch.empty_like(z)
    stats = torch.empty(T, 2, device=dev)
    call("plb_ln_fwd", ptr(z), T, H, ptr(w), ptr(b), 1e-12, ptr(y), ptr(stats), S())
    zf = z.float().requires_grad_(True)
    wf, bf_ = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(zf, (H,), wf, bf_, 1e-12)
    assert rel(y.float(), ref) < 5e-3
    assert rel(stats[:, 0], zf.mean(1)) < 1e-5 and rel(stats[:, 1], 1 / torch.sqrt(zf.var(1, unbiased=False) + 1e-12)) < 1e-4
    dy = torch.randn(T, H, device=dev).to(BF)
    ref.backward(dy.float())
    dz = torch.empty_like(z)
    dw = torch.ones(H, device=dev)
    db = torch.ones(H, device=dev)
    cs = torch.ones(H, device=dev)
    call("plb_ln_bwd", ptr(dy), ptr(z), ptr(stats), ptr(w), T, H, ptr(dz), ptr(dw), ptr(db), ptr(cs), S())
    assert rel(dz.float(), zf.grad) < 5e-3
    assert rel(dw - 1, wf.grad) < 1e-4 and rel(db - 1, bf_.grad) < 1e-4  # accumulated on top of the 1s
    assert rel(cs - 1, dz.float().sum(0)) < 1e-4  # bias gradient = column sum of the stored (rounded) dz


def test_layernorm_constant_row_is_finite(dev):
    z = torch.full((4, 768), 3.0, device=dev).to(BF)
    w, b = torch.ones(768, device=dev), torch.zeros(768, device=dev)
    y, stats = torch.empty_like(z), torch.empty(4, 2, device=dev)
    call("plb_ln_fwd", ptr(z), 4, 768, ptr(w), ptr(b), 1e-12, ptr(y), ptr(stats), S())
    assert torch.isfinite(y.float()).all() and (y.float() == 0).all()


# ------------------------------------------------------------------------------------------ embeddings
@pytest.mark.parametrize("T,E,V,P", [(2000, 128, 188, 512), (333, 32, 188, 64), (5000, 128, 600, 512)])
def test_embed_ln_fwd_bwd(dev, T, E, V, P):
    ids = torch.randint(0, V, (T,), device=dev, dtype=torch.int32)
    ids[:5] = 0
    pos = torch.randint(0, P, (T,), device=dev, dtype=torch.int32)
    word = torch.randn(V, E, device=dev, requires_grad=True)
    pe = torch.randn(P, E, device=dev, requires_grad=True)
    typ = torch.randn(2, E, device=dev, requires_grad=True)
    w = torch.randn(E, device=dev, requires_grad=True)
    b = torch.randn(E, device=dev, requires_grad=True)
    out = torch.empty(T, E, device=dev, dtype=BF)
    stats = torch.empty(T, 2, device=dev)
    call("plb_embed_ln_fwd", ptr(ids), ptr(pos), T, E, ptr(word), ptr(pe), ptr(typ), ptr(w), ptr(b), 1e-12, ptr(out), ptr(stats), S())
    ref = torch.nn.functional.layer_norm((word[ids.long()] + typ[0]) + pe[pos.long()], (E,), w, b, 1e-12)
    assert rel(out.float(), ref) < 5e-3
    de = torch.randn(T, E, device=dev).to(BF)
    ref.backward(de.float())
    g = [torch.zeros_like(t) for t in (word, pe, typ[0], w, b)]
    call("plb_embed_ln_bwd", ptr(de), ptr(ids), ptr(pos), T, E, V, ptr(word), ptr(pe), ptr(typ), ptr(w), ptr(stats), 0,
         ptr(g[0]), ptr(g[1]), ptr(g[2]), ptr(g[3]), ptr(g[4]), S())
    wg = word.grad.clone()
    wg[0] = 0  # padding_idx row
    assert not g[0][0].any()
    assert rel(g[0], wg) < 1e-4 and rel(g[1], pe.grad) < 1e-4 and rel(g[2], typ.grad[0]) < 1e-4
    assert rel(g[3], w.grad) < 1e-4 and rel(g[4], b.grad) < 1e-4


# ------------------------------------------------------------------------------------------ dropout
@pytest.mark.parametrize("rows,cols,p", [(1000, 768, 0.1), (77, 128, 0.5), (4096, 2048, 0.25), (5, 8, 0.0)])
def test_dropout_kernel_matches_replica_mask(dev, rows, cols, p):
    """plb_dropout == x * mask (+ residual) with the mask of oracle/dropout_rng.py, element for element; keep rate and scale
    as torch.nn.Dropout (keep with probability 1 - p, scale 1 / (1 - p)); the same call is its own backward."""
    from oracle import dropout_rng as R
    seed, stream = 0x1234ABCD5678EF01, 7
    x = torch.randn(rows, cols, device=dev).to(BF)
    res = torch.randn(rows, cols, device=dev).to(BF)
    y = torch.empty_like(x)
    call("plb_dropout", ptr(x), cols, 0, 0, ptr(y), cols, rows, cols, p, seed, stream, S())
    mask = torch.from_numpy(R.hidden_mask(rows, cols, p, seed, stream)).to(dev)
    assert torch.equal(y, (x.float() * mask.float()).to(BF))
    keep = (mask != 0).double().mean().item()
    assert abs(keep - (1 - p)) < 4 * math.sqrt(max(p * (1 - p), 1e-12) / (rows * cols)) + 2e-5
    if p > 0:
        assert torch.allclose(mask[mask != 0], torch.full_like(mask[mask != 0], 1 / (1 - p)))
    # residual form (z1 = x + dropout(attn_out), HF :201-202), in place over x
    y2 = x.clone()
    call("plb_dropout", ptr(y2), cols, ptr(res), cols, ptr(y2), cols, rows, cols, p, seed, stream, S())
    want = torch.addcmul(res.float(), x.float(), mask.float()).to(BF)   # one rounding, like the kernel's fma
    assert (y2.float() - want.float()).abs().max().item() <= 1e-2 * want.float().abs().max().item()
    # another stream -> another mask
    if p > 0:
        call("plb_dropout", ptr(x), cols, 0, 0, ptr(y), cols, rows, cols, p, seed, stream + 1, S())
        assert not torch.equal(y, (x.float() * mask.float()).to(BF))


# ------------------------------------------------------------------------------------------ attention
def _attn_ref(qkv, B, L, nh, lens):
    T, H3 = qkv.shape
    H = H3 // 3
    q, k, v = [t.view(B, L, nh, 64).transpose(1, 2) for t in qkv.split(H, dim=1)]
    s = q @ k.transpose(-1, -2) / 8.0
    key_ok = torch.arange(L, device=qkv.device)[None, :] < lens[:, None]
    s = s.masked_fill(~key_ok[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    return (p @ v).transpose(1, 2).reshape(T, H), torch.logsumexp(s, dim=-1)


def _attn_bwd(name, *args):
    """plb_attn_bwd_tc (two kernels) or plb_attn_bwd_fused (one kernel + caller-owned fp32 workspace); `args` are the common
    arguments up to dbias, in ABI order (ints / device pointers)."""
    if name == "plb_attn_bwd_fused":
        B, nh = args[5], args[6]
        nbytes = _abi.load().plb_attn_bwd_fused_workspace_bytes(B, nh)
        assert nbytes > 0
        guard = 4096
        ws = torch.full((nbytes // 4 + guard,), float("nan"), device="cuda")   # contents are irrelevant on entry
        ws[nbytes // 4:] = 1234.5                                               # ... and nothing past the declared size is touched
        call(name, *args, ptr(ws), nbytes, S())
        assert bool((ws[nbytes // 4:] == 1234.5).all()), "attention backward wrote past its workspace"
        return ws
    call(name, *args, S())


BWD_IMPLS = ["plb_attn_bwd_tc", "plb_attn_bwd_fused"]


@pytest.mark.parametrize("B,L,nh,lens", [(2, 64, 2, [64, 64]), (3, 200, 2, [200, 77, 1]), (2, 512, 12, [512, 300]),
                                         (4, 130, 3, [130, 129, 65, 64]),
                                         # tile / chunk parities of the two-group schedules: 1, 2, 3 query tiles, odd chunk
                                         # counts, fewer key tiles than query tiles, a single valid key
                                         (2, 128, 1, [128, 5]), (2, 256, 2, [256, 130]), (3, 384, 2, [384, 100, 257]),
                                         (4, 512, 2, [1, 128, 129, 511]), (3, 320, 1, [320, 64, 200]), (2, 448, 1, [448, 449 - 64])])
@pytest.mark.parametrize("bwd", BWD_IMPLS)
def test_attention_fwd_bwd(dev, B, L, nh, lens, bwd, fwd="plb_attn_fwd_tc"):
    H, T = nh * 64, B * L
    qkv = torch.randn(T, 3 * H, device=dev).to(BF)
    lens_t = torch.tensor(lens, device=dev, dtype=torch.int32)
    seq_start = torch.arange(B, device=dev, dtype=torch.int32) * L
    q_len = torch.full((B,), L, device=dev, dtype=torch.int32)
    ctx = torch.full((T, H), float("nan"), device=dev, dtype=BF)
    lse = torch.zeros(nh, T, device=dev)
    call(fwd, ptr(qkv), 3 * H, ptr(seq_start), ptr(q_len), ptr(lens_t), B, nh, 64, L, T, ptr(ctx), H, ptr(lse), S())
    qf = qkv.float().requires_grad_(True)
    ref, ref_lse = _attn_ref(qf, B, L, nh, lens_t.long())
    assert torch.isfinite(ctx.float()).all()  # padded QUERY rows are computed too, like the reference
    assert rel(ctx.float(), ref) < 8e-3
    got_lse = lse.view(nh, B, L).permute(1, 0, 2) * math.log(2.0)
    assert (got_lse - ref_lse).abs().max() < 2e-3

    dctx = torch.randn(T, H, device=dev).to(BF)
    valid = (torch.arange(L, device=dev)[None, :] < lens_t[:, None]).reshape(T)
    dctx[~valid] = 0  # the loss never reads padded rows (R/train.py:125), so their upstream gradient is zero
    ref.backward(dctx.float())
    dqkv = torch.full((T, 3 * H), float("nan"), device=dev, dtype=BF)
    delta = torch.zeros(nh, T, device=dev)
    dbias = torch.ones(3 * H, device=dev)
    _attn_bwd(bwd, ptr(qkv), 3 * H, ptr(seq_start), ptr(q_len), ptr(lens_t), B, nh, 64, L, T, ptr(ctx), ptr(dctx), H,
              ptr(lse), ptr(delta), ptr(dqkv), 3 * H, ptr(dbias))
    # fused bias gradient = column sums of the stored (rounded) dqkv, accumulated on top of the initial ones
    want = dqkv.float().sum(0)
    assert (dbias - 1 - want).abs().max() <= 1e-3 * want.abs().max() + 1e-4
    # ctx == NULL: delta is an input (the O-projection DGRAD's ROWDOT epilogue produced it) -> identical gradients
    d_in = (dctx.float() * ctx.float()).view(T, nh, 64).sum(-1).T.contiguous()
    dqkv2 = torch.full((T, 3 * H), float("nan"), device=dev, dtype=BF)
    _attn_bwd(bwd, ptr(qkv), 3 * H, ptr(seq_start), ptr(q_len), ptr(lens_t), B, nh, 64, L, T, 0, ptr(dctx), H,
              ptr(lse), ptr(d_in), ptr(dqkv2), 3 * H, 0)
    assert rel(dqkv2.float(), dqkv.float()) < 2e-3
    assert torch.isfinite(dqkv.float()).all()
    for name, sl in (("dq", slice(0, H)), ("dk", slice(H, 2 * H)), ("dv", slice(2 * H, 3 * H))):
        assert rel(dqkv[:, sl].float(), qf.grad[:, sl]) < 2e-2, name
    # masked keys: exactly zero gradient for K and V
    assert not dqkv[~valid][:, H:].any()


@pytest.mark.parametrize("B,L,nh,lens,p", [(2, 200, 2, [200, 77], 0.2), (3, 512, 3, [512, 300, 129], 0.1), (2, 128, 1, [128, 5], 0.5)])
def test_attention_probability_dropout(dev, B, L, nh, lens, p):
    """plb_attn_fwd_tc_dropout / plb_attn_bwd_fused_dropout == dropout(softmax(S)) V and its autograd with the SAME mask (HF :134,
    :194), the mask rebuilt by the numpy replica of the generator; p = 0 through the dropout entry points == the plain kernels."""
    from oracle import dropout_rng as R
    seed, stream = 0xC0FFEE1234567, 11
    H, T = nh * 64, B * L
    qkv = torch.randn(T, 3 * H, device=dev).to(BF)
    lens_t = torch.tensor(lens, device=dev, dtype=torch.int32)
    seq_start = torch.arange(B, device=dev, dtype=torch.int32) * L
    q_len = torch.full((B,), L, device=dev, dtype=torch.int32)
    ctx = torch.full((T, H), float("nan"), device=dev, dtype=BF)
    lse = torch.zeros(nh, T, device=dev)
    call("plb_attn_fwd_tc_dropout", ptr(qkv), 3 * H, ptr(seq_start), ptr(q_len), ptr(lens_t), B, nh, 64, L, T, ptr(ctx), H, ptr(lse),
         p, seed, stream, S())
    mask = torch.from_numpy(R.attention_mask(B * nh, L, L, p, seed, stream)).to(dev).float().view(B, nh, L, L)
    qf = qkv.float().requires_grad_(True)
    q, k, v = [t.view(B, L, nh, 64).transpose(1, 2) for t in qf.split(H, dim=1)]
    sc = q @ k.transpose(-1, -2) / 8.0
    key_ok = torch.arange(L, device=dev)[None, :] < lens_t[:, None]
    sc = sc.masked_fill(~key_ok[:, None, None, :], float("-inf"))
    ref = ((torch.softmax(sc, dim=-1) * mask) @ v).transpose(1, 2).reshape(T, H)
    assert torch.isfinite(ctx.float()).all()
    assert rel(ctx.float(), ref) < 1e-2
    got_lse = lse.view(nh, B, L).permute(1, 0, 2) * math.log(2.0)
    assert (got_lse - torch.logsumexp(sc, dim=-1)).abs().max() < 2e-3      # the normaliser is the undropped one
    dctx = torch.randn(T, H, device=dev).to(BF)
    valid = key_ok.reshape(T)
    dctx[~valid] = 0
    ref.backward(dctx.float())
    nbytes = _abi.load().plb_attn_bwd_fused_workspace_bytes(B, nh)
    ws = torch.empty(nbytes // 4, device=dev)
    dqkv = torch.full((T, 3 * H), float("nan"), device=dev, dtype=BF)
    delta = torch.zeros(nh, T, device=dev)
    call("plb_attn_bwd_fused_dropout", ptr(qkv), 3 * H, ptr(seq_start), ptr(q_len), ptr(lens_t), B, nh, 64, L, T, ptr(ctx), ptr(dctx),
         H, ptr(lse), ptr(delta), ptr(dqkv), 3 * H, 0, ptr(ws), nbytes, p, seed, stream, S())
    assert torch.isfinite(dqkv.float()).all()
    for name, sl in (("dq", slice(0, H)), ("dk", slice(H, 2 * H)), ("dv", slice(2 * H, 3 * H))):
        assert rel(dqkv[:, sl].float(), qf.grad[:, sl]) < 3e-2, name
    assert not dqkv[~valid][:, H:].any()
    # a different stream must give a different mask (and a clearly different result)
    ctx2 = torch.empty_like(ctx)
    call("plb_attn_fwd_tc_dropout", ptr(qkv), 3 * H, ptr(seq_start), ptr(q_len), ptr(lens_t), B, nh, 64, L, T, ptr(ctx2), H, ptr(lse),
         p, seed, stream + 1, S())
    assert rel(ctx2.float(), ref) > 5e-2
    # p = 0 == the plain kernel, bit for bit
    ctx3, ctx4 = torch.empty_like(ctx), torch.empty_like(ctx)
    call("plb_attn_fwd_tc_dropout", ptr(qkv), 3 * H, ptr(seq_start), ptr(q_len), ptr(lens_t), B, nh, 64, L, T, ptr(ctx3), H, ptr(lse),
         0.0, seed, stream, S())
    call("plb_attn_fwd_tc", ptr(qkv), 3 * H, ptr(seq_start), ptr(q_len), ptr(lens_t), B, nh, 64, L, T, ptr(ctx4), H, ptr(lse), S())
    assert torch.equal(ctx3, ctx4)


@pytest.mark.parametrize("bwd", BWD_IMPLS)
@pytest.mark.parametrize("lens", [[512, 300, 1, 77], [65, 64, 130, 129, 128, 127], [200]])
def test_attention_tc_packed_layout(dev, lens, bwd):
    """Packed (cu_seqlens) layout through the same ABI: q_len == kv_len == the sequence length, sequences back to back.
    Rows past a sequence's end belong to its neighbour and must be neither read as keys nor written."""
    nh = 2
    H, B = nh * 64, len(lens)
    starts = [0]
    for n in lens[:-1]:
        starts.append(starts[-1] + n)
    T = sum(lens)
    qkv = torch.randn(T, 3 * H, device=dev).to(BF)
    lens_t = torch.tensor(lens, device=dev, dtype=torch.int32)
    seq_start = torch.tensor(starts, device=dev, dtype=torch.int32)
    ctx = torch.full((T, H), float("nan"), device=dev, dtype=BF)
    lse = torch.zeros(nh, T, device=dev)
    call("plb_attn_fwd_tc", ptr(qkv), 3 * H, ptr(seq_start), ptr(lens_t), ptr(lens_t), B, nh, 64, max(lens), T, ptr(ctx), H,
         ptr(lse), S())
    qf = qkv.float().requires_grad_(True)
    outs = []
    for b0, n in zip(starts, lens):
        q, k, v = [t.view(n, nh, 64).transpose(0, 1) for t in qf[b0:b0 + n].split(H, dim=1)]
        outs.append((torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1) @ v).transpose(0, 1).reshape(n, H))
    ref = torch.cat(outs)
    assert torch.isfinite(ctx.float()).all() and rel(ctx.float(), ref) < 8e-3
    dctx = torch.randn(T, H, device=dev).to(BF)
    ref.backward(dctx.float())
    dqkv = torch.full((T, 3 * H), float("nan"), device=dev, dtype=BF)
    delta = torch.zeros(nh, T, device=dev)
    _attn_bwd(bwd, ptr(qkv), 3 * H, ptr(seq_start), ptr(lens_t), ptr(lens_t), B, nh, 64, max(lens), T, ptr(ctx),
              ptr(dctx), H, ptr(lse), ptr(delta), ptr(dqkv), 3 * H, 0)
    assert torch.isfinite(dqkv.float()).all()
    for name, sl in (("dq", slice(0, H)), ("dk", slice(H, 2 * H)), ("dv", slice(2 * H, 3 * H))):
        assert rel(dqkv[:, sl].float(), qf.grad[:, sl]) < 2e-2, name


@pytest.mark.parametrize("bwd", BWD_IMPLS)
@pytest.mark.parametrize("seed", [0, 1])
def test_attention_tc_many_items_per_cta_variable_lengths(dev, seed, bwd):
    """The forward kernel is persistent (one CTA per SM walks several (sequence, head) items): more items than SMs with
    lengths that give the second softmax group NO tile in runs of consecutive items (<= 128 rows), one, or two tiles.
    Regression for the run-ahead of a tile-less group (round 2: deadlock -> trap on BASELINE configs[3] batches)."""
    nh = 6
    H = nh * 64
    g = torch.Generator().manual_seed(seed)
    short = torch.randint(1, 129, (40,), generator=g)
    mixed = torch.randint(1, 513, (32,), generator=g)
    lens = torch.cat([short, mixed])[torch.randperm(72, generator=g)].tolist()   # 72 x 6 = 432 items on 148 CTAs
    starts = [0]
    for n in lens[:-1]:
        starts.append(starts[-1] + n)
    T, B = sum(lens), len(lens)
    qkv = (torch.randn(T, 3 * H, generator=g) * 1.5).to(dev).to(BF)
    lens_t = torch.tensor(lens, device=dev, dtype=torch.int32)
    seq_start = torch.tensor(starts, device=dev, dtype=torch.int32)
    ctx = torch.full((T, H), float("nan"), device=dev, dtype=BF)
    lse = torch.zeros(nh, T, device=dev)
    for _ in range(3):   # the barrier phases of a persistent CTA depend on timing: repeat
        call("plb_attn_fwd_tc", ptr(qkv), 3 * H, ptr(seq_start), ptr(lens_t), ptr(lens_t), B, nh, 64, max(lens), T, ptr(ctx), H,
             ptr(lse), S())
        torch.cuda.synchronize()
    qf = qkv.float().requires_grad_(True)
    outs = []
    for b0, n in zip(starts, lens):
        q, k, v = [t.view(n, nh, 64).transpose(0, 1) for t in qf[b0:b0 + n].split(H, dim=1)]
        outs.append((torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1) @ v).transpose(0, 1).reshape(n, H))
    ref = torch.cat(outs)
    assert torch.isfinite(ctx.float()).all()
    worst = max(rel(ctx[b0:b0 + n].float(), ref[b0:b0 + n]) for b0, n in zip(starts, lens))
    assert worst < 1e-2, worst
    dctx = torch.randn(T, H, generator=g).to(dev).to(BF)
    ref.backward(dctx.float())
    dqkv = torch.full((T, 3 * H), float("nan"), device=dev, dtype=BF)
    delta = torch.zeros(nh, T, device=dev)
    dbias = torch.zeros(3 * H, device=dev)
    for rep in range(3):   # (persistent kernel: the barrier phases depend on timing, and the workspace starts as NaN)
        dqkv.fill_(float("nan"))
        dbias.zero_()
        _attn_bwd(bwd, ptr(qkv), 3 * H, ptr(seq_start), ptr(lens_t), ptr(lens_t), B, nh, 64, max(lens), T, ptr(ctx),
                  ptr(dctx), H, ptr(lse), ptr(delta), ptr(dqkv), 3 * H, ptr(dbias))
        torch.cuda.synchronize()
        assert torch.isfinite(dqkv.float()).all()
        worst = {name: max(rel(dqkv[b0:b0 + n, sl].float(), qf.grad[b0:b0 + n, sl]) for b0, n in zip(starts, lens))
                 for name, sl in (("dq", slice(0, H)), ("dk", slice(H, 2 * H)), ("dv", slice(2 * H, 3 * H)))}
        assert max(worst.values()) < 3e-2, (rep, worst)
        want = dqkv.float().sum(0)
        assert (dbias - want).abs().max() <= 1e-3 * want.abs().max() + 1e-3


# ------------------------------------------------------------------------------------------ masked-row loss path
def test_gather_scatter_bit_exact(dev):
    T, H, M = 5000, 768, 700
    src = torch.randn(T, H, device=dev).to(BF)
    rows = torch.randperm(T, device=dev)[:M].to(torch.int32)
    dst = torch.empty(M, H, device=dev, dtype=BF)
    call("plb_gather_rows", ptr(src), H, ptr(rows), M, H, ptr(dst), H, S())
    assert torch.equal(dst, src[rows.long()])
    acc = torch.zeros(T, H, device=dev, dtype=BF)
    call("plb_scatter_add_rows", ptr(dst), H, ptr(rows), M, H, ptr(acc), H, S())
    ref = torch.zeros_like(acc)
    ref[rows.long()] = dst
    assert torch.equal(acc, ref)
    # duplicates accumulate
    rows2 = torch.zeros(4, device=dev, dtype=torch.int32)
    ones = torch.ones(4, H, device=dev, dtype=BF)
    acc.zero_()
    call("plb_scatter_add_rows", ptr(ones), H, ptr(rows2), 4, H, ptr(acc), H, S())
    assert (acc[0] == 4).all() and not acc[1:].any()


@pytest.mark.parametrize("M,V", [(1000, 188), (37, 97), (300, 5000), (300, 5003), (200, 64000), (151, 65536), (40, 70000),
                                 (5, 9001), (300, 40003), (1, 64000), (1300, 16385), (3, 131072), (2, 131075)])
def test_masked_ce(dev, M, V):
    ld = (V + 7) // 8 * 8
    z = torch.zeros(M, ld, device=dev, dtype=BF)
    z[:, :V] = (torch.randn(M, V, device=dev) * 3).to(BF)
    z[:, V:] = 7.0  # padding columns must be ignored and then zeroed
    labels = torch.randint(0, V, (M,), device=dev, dtype=torch.int32)
    w = torch.rand(M, device=dev) / M
    zf = z[:, :V].float().requires_grad_(True)
    ref_rows = torch.nn.functional.cross_entropy(zf, labels.long(), reduction="none") * w
    ref_rows.sum().backward()
    row_loss = torch.empty(M, device=dev)
    loss = torch.full((1,), 5.0, device=dev)
    call("plb_masked_ce", ptr(z), ld, M, V, ptr(labels), ptr(w), ptr(row_loss), ptr(loss), 0, S())
    assert rel(row_loss, ref_rows) < 1e-5
    assert abs(loss.item() - ref_rows.sum().item()) < 1e-5 * abs(ref_rows.sum().item()) + 1e-7
    assert rel(z[:, :V].float(), zf.grad) < 6e-3
    assert not z[:, V:].any()
    z2 = torch.zeros(M, ld, device=dev, dtype=BF)
    call("plb_masked_ce", ptr(z2), ld, M, V, ptr(labels), ptr(w), ptr(row_loss), ptr(loss), 1, S())  # accumulate
    assert abs(loss.item() - (ref_rows.sum().item() + (w * math.log(V)).sum().item())) < 1e-4
    call("plb_masked_ce", ptr(z2), ld, 0, V, ptr(labels), ptr(w), ptr(row_loss), ptr(loss), 0, S())  # empty -> 0
    assert loss.item() == 0.0


@pytest.mark.parametrize("M,V", [(64, 188), (64, 5003), (130, 64000), (9, 9001)])
def test_masked_ce_label_outside_vocabulary(dev, M, V):
    """A label outside [0, V) contributes neither loss nor gradient (every kernel variant: warp-per-row, CTA-per-row, cluster)."""
    ld = (V + 7) // 8 * 8
    z = torch.zeros(M, ld, device=dev, dtype=BF)
    z[:, :V] = (torch.randn(M, V, device=dev) * 2).to(BF)
    labels = torch.randint(0, V, (M,), device=dev, dtype=torch.int32)
    bad = torch.arange(M, device=dev) % 3 == 1
    labels[bad] = torch.where(torch.arange(M, device=dev)[bad] % 2 == 0, torch.tensor(V, device=dev), torch.tensor(-1, device=dev)).int()
    w = torch.rand(M, device=dev) / M
    zf = z[:, :V].float().requires_grad_(True)
    safe = labels.long().clamp(0, V - 1)
    ref_rows = torch.nn.functional.cross_entropy(zf, safe, reduction="none") * w * (~bad)
    ref_rows.sum().backward()
    row_loss = torch.full((M,), 9.0, device=dev)
    loss = torch.zeros(1, device=dev)
    call("plb_masked_ce", ptr(z), ld, M, V, ptr(labels), ptr(w), ptr(row_loss), ptr(loss), 0, S())
    assert rel(row_loss, ref_rows) < 1e-5
    assert (row_loss[bad] == 0).all() and not z[bad].any()
    assert rel(z[:, :V].float(), zf.grad) < 6e-3


@pytest.mark.parametrize("T,N", [(3000, 2304), (100, 188), (98304, 768)])
def test_colsum(dev, T, N):
    ld = (N + 7) // 8 * 8
    x = torch.zeros(T, ld, device=dev, dtype=BF)
    x[:, :N] = torch.randn(T, N, device=dev).to(BF)
    out = torch.ones(N, device=dev)
    call("plb_colsum", ptr(x), ld, T, N, ptr(out), S())
    assert rel(out - 1, x[:, :N].float().sum(0)) < 1e-4


@pytest.mark.parametrize("nbytes,offset", [(1, 0), (15, 1), (16, 0), (4, 4), (1000, 3), (4096, 0), (1 << 20, 8), ((1 << 22) + 5, 7)])
def test_fill_zero(dev, nbytes, offset):
    """plb_fill_zero: exactly [dst, dst + bytes) is zeroed, any alignment (guard bytes on both sides stay untouched)."""
    buf = torch.full((nbytes + offset + 64,), 0xAB, dtype=torch.uint8, device=dev)
    call("plb_fill_zero", buf.data_ptr() + offset, nbytes, S())
    assert not buf[offset:offset + nbytes].any()
    assert (buf[:offset] == 0xAB).all() and (buf[offset + nbytes:] == 0xAB).all()
    call("plb_fill_zero", buf.data_ptr(), 0, S())   # nothing to do


# ------------------------------------------------------------------------------------------ AdamW
def test_adamw_matches_torch(dev):
    n = 4096 * 3 + 64
    p = torch.randn(n, device=dev)
    ref_p = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref_p], lr=7e-5)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    shadow = torch.zeros(n, device=dev, dtype=BF)
    for step in range(1, 4):
        g = torch.randn(n, device=dev) * 0.01
        ref_p.grad = g.clone()
        opt.step()
        gg = (g * 2).contiguous()  # grad_scale 0.5 undoes this (1/world after a SUM all-reduce)
        call("plb_adamw_step", ptr(p), ptr(gg), ptr(m), ptr(v), ptr(shadow), n, step, 7e-5, 0.9, 0.999, 1e-8, 0.01, 0.5, 1, S())
        assert not gg.any()  # gradient arena zeroed for the next accumulation
        upd, upd_ref = p - ref_p.detach(), None
        assert (p - ref_p.detach()).abs().max() < 1e-6  # a few fp32 ulps at |p| ~ 4 (op order differs from torch)
    assert torch.equal(shadow, p.to(BF))
    st = opt.state[ref_p]
    assert rel(m, st["exp_avg"]) < 1e-6 and rel(v, st["exp_avg_sq"]) < 1e-6
