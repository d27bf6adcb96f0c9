"""The B200 training engine: flat parameter arenas, activation workspace and the kernel schedule of one
PL-BERT step (forward, masked-row loss, backward, AdamW) over the C ABI in ``include/plbert_b200.h``.

Mapping to the reference: ``forward`` = HF ``AlbertModel.forward`` (HF modeling_albert.py:387-423,
pooler skipped) ; ``loss_and_backward`` = R/train.py:381-390 ``process_batch`` + ``loss.backward()`` (R/train.py:356);
``adamw_step`` = ``optimizer.step()`` (R/train.py:357).  Everything runs on the current CUDA stream; nothing here
synchronises with the host.

HBM layout
  * fp32 arenas ``params / grads / exp_avg / exp_avg_sq`` share one element layout (``ParamLayout``): heads first
    (their gradients are final first, so their all-reduce overlaps the encoder backward), then the shared ALBERT layer
    with Wq|Wk|Wv and bq|bk|bv adjacent (so the fused [3H, H] QKV projection is a view), then the embedding-side
    tensors, then the never-trained pooler.  ``nn.Parameter.data`` of the wrapped modules are views into ``params``
    (``state_dict`` keeps the reference's names/shapes), ``.grad`` are views into ``grads``.
  * a bf16 shadow of ``params`` with the same layout is what the GEMMs read; AdamW rewrites it in the same pass.
  * activations are bf16 [T, *] row-major in the padded layout (row = b*L + j), one set per layer application.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import _abi
from ._abi import EPI_GELU, EPI_GELU_BWD, EPI_NONE, EPI_RESIDUAL, EPI_ROWDOT, GEMM_DGRAD, GEMM_FWD, GEMM_WGRAD, call, ptr
from .batch import DeviceBatch

_LAYER = "encoder.encoder.albert_layer_groups.0.albert_layers.0."
_EMB = "encoder.embeddings."
ALIGN = 64  # elements: every tensor (group) starts on a 256-byte (fp32) / 128-byte (bf16) boundary


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# NVTX ranges around the phases of a step (forward / heads / encoder backward / AdamW) when PLBERT_B200_NVTX=1, so that
# `ncu --nvtx --nvtx-include "backward/"` or a timeline tool can select them; a no-op otherwise.
_NVTX = os.environ.get("PLBERT_B200_NVTX", "0") not in ("", "0")


class _Range:
    __slots__ = ("name",)

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        return False


@dataclass
class EngineConfig:
    vocab_size: int
    embedding_size: int
    hidden_size: int
    num_attention_heads: int
    intermediate_size: int
    num_hidden_layers: int
    max_position_embeddings: int
    type_vocab_size: int
    layer_norm_eps: float
    pad_token_id: int
    num_phonemes: int
    num_tokens: Optional[int]
    hidden_dropout_prob: float = 0.0             # HF :109 (embeddings) and :201 (attention output projection)
    attention_probs_dropout_prob: float = 0.0    # HF :134 / :194

    @classmethod
    def from_albert(cls, cfg, num_phonemes: int, num_tokens: Optional[int]) -> "EngineConfig":
        problems = []
        if getattr(cfg, "num_hidden_groups", 1) != 1 or getattr(cfg, "inner_group_num", 1) != 1:
            problems.append("num_hidden_groups/inner_group_num must be 1 (one shared layer)")
        if cfg.hidden_act != "gelu_new":
            problems.append(f"hidden_act {cfg.hidden_act!r} (only gelu_new)")
        for pname in ("hidden_dropout_prob", "attention_probs_dropout_prob"):
            if not 0.0 <= float(getattr(cfg, pname)) < 1.0:
                problems.append(f"{pname} must be in [0, 1)")
        if cfg.hidden_size % cfg.num_attention_heads or cfg.hidden_size // cfg.num_attention_heads != 64:
            problems.append("head_dim must be 64")
        if cfg.hidden_size % 8 or cfg.intermediate_size % 8 or cfg.embedding_size % 8:
            problems.append("hidden/intermediate/embedding sizes must be multiples of 8")
        if problems:
            raise _abi.PlbError("AlbertConfig not supported by the B200 path: " + "; ".join(problems))
        return cls(cfg.vocab_size, cfg.embedding_size, cfg.hidden_size, cfg.num_attention_heads, cfg.intermediate_size,
                   cfg.num_hidden_layers, cfg.max_position_embeddings, cfg.type_vocab_size, float(cfg.layer_norm_eps),
                   0 if cfg.pad_token_id is None else int(cfg.pad_token_id), num_phonemes, num_tokens,
                   float(cfg.hidden_dropout_prob), float(cfg.attention_probs_dropout_prob))


class ParamLayout:
    """Element offsets of every state_dict tensor inside the flat arenas."""

    def __init__(self, c: EngineConfig):
        E, H, I = c.embedding_size, c.hidden_size, c.intermediate_size
        groups: List[List[Tuple[str, Tuple[int, ...]]]] = []
        if c.num_tokens is not None:
            groups += [[("token_predictor.weight", (c.num_tokens, H))], [("token_predictor.bias", (c.num_tokens,))]]
        groups += [[("phoneme_predictor.weight", (c.num_phonemes, H))], [("phoneme_predictor.bias", (c.num_phonemes,))]]
        self.n_head_elems_end = None  # filled below: end of the head region
        att = _LAYER + "attention."
        enc_groups = [
            [(att + "query.weight", (H, H)), (att + "key.weight", (H, H)), (att + "value.weight", (H, H))],
            [(att + "query.bias", (H,)), (att + "key.bias", (H,)), (att + "value.bias", (H,))],
            [(att + "dense.weight", (H, H))], [(att + "dense.bias", (H,))],
            [(att + "LayerNorm.weight", (H,))], [(att + "LayerNorm.bias", (H,))],
            [(_LAYER + "ffn.weight", (I, H))], [(_LAYER + "ffn.bias", (I,))],
            [(_LAYER + "ffn_output.weight", (H, I))], [(_LAYER + "ffn_output.bias", (H,))],
            [(_LAYER + "full_layer_layer_norm.weight", (H,))], [(_LAYER + "full_layer_layer_norm.bias", (H,))],
            [("encoder.encoder.embedding_hidden_mapping_in.weight", (H, E))],
            [("encoder.encoder.embedding_hidden_mapping_in.bias", (H,))],
            [(_EMB + "word_embeddings.weight", (c.vocab_size, E))],
            [(_EMB + "position_embeddings.weight", (c.max_position_embeddings, E))],
            [(_EMB + "token_type_embeddings.weight", (c.type_vocab_size, E))],
            [(_EMB + "LayerNorm.weight", (E,))], [(_EMB + "LayerNorm.bias", (E,))],
        ]
        frozen = [[("encoder.pooler.weight", (H, H))], [("encoder.pooler.bias", (H,))]]
        self.offsets: Dict[str, int] = {}
        self.shapes: Dict[str, Tuple[int, ...]] = {}
        off = 0

        def place(gs):
            nonlocal off
            for g in gs:
                off = (off + ALIGN - 1) // ALIGN * ALIGN
                for name, shape in g:
                    n = 1
                    for s in shape:
                        n *= s
                    self.offsets[name], self.shapes[name] = off, shape
                    off += n

        place(groups)
        off = (off + ALIGN - 1) // ALIGN * ALIGN
        self.head_end = off          # [0, head_end): head parameters (first all-reduce bucket)
        place(enc_groups)
        off = (off + ALIGN - 1) // ALIGN * ALIGN
        self.trainable_end = off     # [0, trainable_end): everything AdamW updates
        place(frozen)
        self.total = (off + ALIGN - 1) // ALIGN * ALIGN
        self.frozen = {n for g in frozen for n, _ in g}

    def numel(self, name: str) -> int:
        n = 1
        for s in self.shapes[name]:
            n *= s
        return n


class Engine:
    def __init__(self, cfg: EngineConfig, device: torch.device):
        _abi.load()
        if device.type != "cuda":
            raise _abi.PlbError("plbert_b200 runs on CUDA devices only (no CPU fallback)")
        self.cfg, self.device = cfg, device
        self.layout = ParamLayout(cfg)
        n = self.layout.total
        self.params = torch.zeros(n, dtype=torch.float32, device=device)
        self.grads = torch.zeros(n, dtype=torch.float32, device=device)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=device)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=device)
        self.shadow = torch.zeros(n, dtype=torch.bfloat16, device=device)
        self.shadow_valid = False
        self.grads_clean = True       # grads arena is all zeros
        self.have_grads = False
        self.opt_step = 0
        self.launches = 0             # kernels enqueued by this engine (bench.py's gpu_launches)
        self._cap_T = 0
        self._cap_M = 0
        self.loss_buf = torch.zeros(1, dtype=torch.float32, device=device)
        # dropout (identity in the reference: its AlbertConfig has p = 0).  Masks are never stored: every kernel regenerates
        # them from (seed, stream) with stream = 64 * (training forward number) + site; sites: 0 embeddings, 1 + 2 l attention
        # probabilities of layer application l, 2 + 2 l its output projection.
        self.dropout_seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        self.dropout_forwards = 0     # training forwards so far (advances the mask stream)
        self._drop_base = 0           # stream base of the forward whose activations are live (the backward reuses it)
        self._drop_p = (0.0, 0.0)     # (hidden, attention) probabilities that forward actually applied
        self.dropout_enabled = True   # the drop-in models mirror nn.Module.training here (eval() -> identity, like HF)
        # column sums of d_u (FFN1 bias gradient) inside the GELU' DGRAD's store warp instead of a separate plb_colsum pass: same-box
        # A/B (profiles/ab_step.py) 55.8 vs 55.5 ms per step — the store warp's serial sums slow that GEMM (226 -> 305 us alone) by more
        # than the 62 us pass costs, so the pass stays the default; the fused form is kept and tested (tests/test_kernels_gpu.py)
        self.fuse_ffn1_bias_sums = False
        self.comm_stream: Optional[torch.cuda.Stream] = None
        self.always_reduce = False    # set by ddp.GradientSynchronizer (world > 1): every step ends in the collectives
        # raw addresses of every tensor inside the (never reallocated) arenas: building torch views costs 3-5 us each and
        # the step makes ~700 of them — host time that is exposed whenever the caller synchronises every step
        lay = self.layout
        self.pp = {n: self.params.data_ptr() + 4 * o for n, o in lay.offsets.items()}   # fp32 master
        self.gp = {n: self.grads.data_ptr() + 4 * o for n, o in lay.offsets.items()}    # fp32 gradient
        self.wp = {n: self.shadow.data_ptr() + 2 * o for n, o in lay.offsets.items()}   # bf16 shadow
        self.name_of_ptr = {v: k for k, v in self.pp.items()}                           # fp32 master address -> name
        self.grad_views = {n: self.view(self.grads, n) for n in lay.offsets}            # what p.grad points at
        # ddp.GradientSynchronizer (NVSwitch multicast available): head weight name -> MULTICAST address of its gradient; the
        # head's WGRAD then adds into every rank's arena at once (plb_gemm_wgrad_allreduce) and needs no NCCL all-reduce
        self.grad_multicast: Dict[str, int] = {}

    def rebind_grads(self, arena: torch.Tensor) -> None:
        """Move the gradient arena into ``arena`` (same size / dtype / device; e.g. a symmetric-memory allocation).  Only
        legal while no gradient is live (before the first step or right after an optimizer step)."""
        if arena.shape != self.grads.shape or arena.dtype != self.grads.dtype or arena.device != self.grads.device:
            raise _abi.PlbError("rebind_grads: the new arena must match the old one")
        if self.have_grads:
            raise _abi.PlbError("rebind_grads: gradients are live")
        arena.zero_()
        self.grads = arena
        self.grads_clean = True
        lay = self.layout
        self.gp = {n: self.grads.data_ptr() + 4 * o for n, o in lay.offsets.items()}
        self.grad_views = {n: self.view(self.grads, n) for n in lay.offsets}

    # ------------------------------------------------------------------ parameter views
    def view(self, arena: torch.Tensor, name: str) -> torch.Tensor:
        o = self.layout.offsets[name]
        return arena[o:o + self.layout.numel(name)].view(self.layout.shapes[name])

    def p(self, name: str) -> torch.Tensor:      # fp32 master
        return self.view(self.params, name)

    def g(self, name: str) -> torch.Tensor:      # fp32 gradient
        return self.view(self.grads, name)

    def w(self, name: str) -> torch.Tensor:      # bf16 shadow
        return self.view(self.shadow, name)

    def refresh_shadow(self) -> None:
        call("plb_cast_bf16", ptr(self.params), ptr(self.shadow), self.layout.total, _stream())
        self.launches += 1
        self.shadow_valid = True

    # ------------------------------------------------------------------ workspace
    def _ensure_workspace(self, T: int, M: int, train: bool) -> None:
        c = self.cfg
        H, I, E = c.hidden_size, c.intermediate_size, c.embedding_size
        dev, bf = self.device, torch.bfloat16
        n_sets = c.num_hidden_layers if train else 2
        if T > self._cap_T or len(getattr(self, "acts", [])) < n_sets:
            cap = max(T, self._cap_T)
            mk = lambda cols, dt=bf: torch.empty((cap, cols), dtype=dt, device=dev)
            self.e, self.e_stats, self.x0 = mk(E), mk(2, torch.float32), mk(H)
            self.acts = []
            for _ in range(n_sets):
                self.acts.append(dict(qkv=mk(3 * H), ctx=mk(H), lse=torch.empty((c.num_attention_heads, cap), dtype=torch.float32, device=dev),
                                      z1=mk(H), st1=mk(2, torch.float32), a=mk(H), dg=mk(I), g=mk(I), z2=mk(H),
                                      st2=mk(2, torch.float32), y=mk(H)))
            if train:
                # backward scratch, reused by every layer
                self.d_y, self.d_z, self.d_a, self.d_u, self.d_qkv, self.d_ctx = mk(H), mk(H), mk(H), mk(I), mk(3 * H), mk(H)
                self.d_e = mk(E)
                self.delta = torch.empty((c.num_attention_heads, cap), dtype=torch.float32, device=dev)
            self._cap_T = cap
        if train:
            self._ensure_row_workspace(M)

    def _ensure_row_workspace(self, M: int) -> None:
        """Buffers of the masked-row head path: gathered hidden rows, their gradient, logits, per-row loss."""
        if M <= self._cap_M:
            return
        c = self.cfg
        H, dev, bf = c.hidden_size, self.device, torch.bfloat16
        cap = max(int(M * 1.25) + 64, self._cap_M)
        vmax = max(c.num_phonemes, c.num_tokens or 0)
        self.h_rows = torch.empty((cap, H), dtype=bf, device=dev)
        self.dh_rows = torch.empty((cap, H), dtype=bf, device=dev)
        self.logits = torch.empty(cap * ((vmax + 7) // 8 * 8), dtype=bf, device=dev)
        self.row_loss = torch.empty(cap, dtype=torch.float32, device=dev)
        self._cap_M = cap

    def _attn_workspace(self, B: int) -> torch.Tensor:
        """fp32 scratch of the attention backward (per-CTA partial sums of dQ over the key tiles; L2-resident, 19 MB at most)."""
        need = _abi.load().plb_attn_bwd_fused_workspace_bytes(B, self.cfg.num_attention_heads) // 4
        ws = getattr(self, "_attn_ws", None)
        if ws is None or ws.numel() < need:
            ws = self._attn_ws = torch.empty(need, dtype=torch.float32, device=self.device)
        return ws

    # ------------------------------------------------------------------ kernels
    def _gemm(self, op, epi, M, N, K, A, lda, B, ldb, C, ldc, bias=None, aux_in=None, ld_in=0, aux_out=None, ld_out=0):
        call("plb_gemm_bf16", op, epi, M, N, K, ptr(A), lda, ptr(B), ldb, ptr(C), ldc, ptr(bias), ptr(aux_in), ld_in,
             ptr(aux_out), ld_out, _stream())
        self.launches += 1

    def _zero(self, t: torch.Tensor) -> None:
        """t[...] = 0 with the library's own fill kernel (t contiguous): the step launches no ATen kernel."""
        call("plb_fill_zero", t.data_ptr(), t.numel() * t.element_size(), _stream())
        self.launches += 1

    def _head_wgrad(self, head: str, V: int, H: int, m: int, logits, ldv: int) -> None:
        """dW_head += dz^T h_rows — into this rank's arena, or (data parallel on NVSwitch) into every rank's at once."""
        mc = self.grad_multicast.get(head + ".weight")
        if mc is None:
            self._gemm(GEMM_WGRAD, EPI_NONE, V, H, m, logits, ldv, self.h_rows, H, self.gp[head + ".weight"], H)
        else:
            call("plb_gemm_wgrad_allreduce", V, H, m, ptr(logits), ldv, ptr(self.h_rows), H, mc, H, _stream())
            self.launches += 1

    def forward(self, b: DeviceBatch, train: bool) -> torch.Tensor:
        """Encoder forward; returns last_hidden_state as bf16 [T, H] (a workspace view)."""
        c = self.cfg
        T, H, I, E = b.T, c.hidden_size, c.intermediate_size, c.embedding_size
        nh = c.num_attention_heads
        if b.L > c.max_position_embeddings:
            raise _abi.PlbError(f"sequence length {b.L} exceeds max_position_embeddings {c.max_position_embeddings}")
        self._ensure_workspace(T, b.M, train)
        if not self.shadow_valid:
            self.refresh_shadow()
        s = _stream()
        att = _LAYER + "attention."
        call("plb_embed_ln_fwd", ptr(b.ids), ptr(b.pos), T, E, self.pp[_EMB + "word_embeddings.weight"],
             self.pp[_EMB + "position_embeddings.weight"], self.pp[_EMB + "token_type_embeddings.weight"],
             self.pp[_EMB + "LayerNorm.weight"], self.pp[_EMB + "LayerNorm.bias"], c.layer_norm_eps,
             ptr(self.e), ptr(self.e_stats), s)
        self.launches += 1
        on = train and self.dropout_enabled                     # nn.Dropout is the identity outside training (HF modules in eval())
        p_hid = c.hidden_dropout_prob if on else 0.0
        p_att = c.attention_probs_dropout_prob if on else 0.0
        if train:
            self._drop_p = (p_hid, p_att)
        if p_hid > 0 or p_att > 0:
            self._drop_base = (self.dropout_forwards * 64) & 0xFFFFFFFF
            self.dropout_forwards += 1
        seed, base = self.dropout_seed, self._drop_base
        if p_hid > 0:
            call("plb_dropout", ptr(self.e), E, 0, 0, ptr(self.e), E, T, E, p_hid, seed, base, s)
            self.launches += 1
        mp = "encoder.encoder.embedding_hidden_mapping_in."
        self._gemm(GEMM_FWD, EPI_NONE, T, H, E, self.e, E, self.wp[mp + "weight"], E, self.x0, H, bias=self.pp[mp + "bias"])
        wqkv, bqkv = self.wp[att + "query.weight"], self.pp[att + "query.bias"]   # q|k|v are adjacent: one [3H, H] matrix
        x = self.x0
        for l in range(c.num_hidden_layers):
            A = self.acts[l if train else l % 2]
            self._gemm(GEMM_FWD, EPI_NONE, T, 3 * H, H, x, H, wqkv, H, A["qkv"], 3 * H, bias=bqkv)
            if p_att > 0:
                call("plb_attn_fwd_tc_dropout", ptr(A["qkv"]), 3 * H, ptr(b.seq_start), ptr(b.q_len), ptr(b.kv_len), b.B, nh, H // nh,
                     b.L, T, ptr(A["ctx"]), H, ptr(A["lse"]), p_att, seed, (base + 1 + 2 * l) & 0xFFFFFFFF, s)
            else:
                call("plb_attn_fwd_tc", ptr(A["qkv"]), 3 * H, ptr(b.seq_start), ptr(b.q_len), ptr(b.kv_len), b.B, nh, H // nh,
                     b.L, T, ptr(A["ctx"]), H, ptr(A["lse"]), s)
            self.launches += 1
            if p_hid > 0:   # z1 = x + dropout(ctx Wo^T + bo)  (HF :200-202)
                self._gemm(GEMM_FWD, EPI_NONE, T, H, H, A["ctx"], H, self.wp[att + "dense.weight"], H, A["z1"], H,
                           bias=self.pp[att + "dense.bias"])
                call("plb_dropout", ptr(A["z1"]), H, ptr(x), H, ptr(A["z1"]), H, T, H, p_hid, seed, (base + 2 + 2 * l) & 0xFFFFFFFF, s)
                self.launches += 1
            else:
                self._gemm(GEMM_FWD, EPI_RESIDUAL, T, H, H, A["ctx"], H, self.wp[att + "dense.weight"], H, A["z1"], H,
                           bias=self.pp[att + "dense.bias"], aux_in=x, ld_in=H)
            call("plb_ln_fwd", ptr(A["z1"]), T, H, self.pp[att + "LayerNorm.weight"], self.pp[att + "LayerNorm.bias"],
                 c.layer_norm_eps, ptr(A["a"]), ptr(A["st1"]), s)
            self._gemm(GEMM_FWD, EPI_GELU, T, I, H, A["a"], H, self.wp[_LAYER + "ffn.weight"], H, A["g"], I,
                       bias=self.pp[_LAYER + "ffn.bias"], aux_out=A["dg"], ld_out=I)   # g = gelu(u), dg = gelu'(u)
            self._gemm(GEMM_FWD, EPI_RESIDUAL, T, H, I, A["g"], I, self.wp[_LAYER + "ffn_output.weight"], I, A["z2"], H,
                       bias=self.pp[_LAYER + "ffn_output.bias"], aux_in=A["a"], ld_in=H)
            call("plb_ln_fwd", ptr(A["z2"]), T, H, self.pp[_LAYER + "full_layer_layer_norm.weight"],
                 self.pp[_LAYER + "full_layer_layer_norm.bias"], c.layer_norm_eps, ptr(A["y"]), ptr(A["st2"]), s)
            self.launches += 2
            x = A["y"]
        return x[:T]

    def head_logits(self, hidden: torch.Tensor, head: str) -> torch.Tensor:
        """Full-position logits of one head (R/model.py:15-16), bf16 [T, V].  Compatibility / inference path."""
        T, H = hidden.shape
        V = self.layout.shapes[head + ".weight"][0]
        ldv = (V + 7) // 8 * 8
        out = torch.empty((T, ldv), dtype=torch.bfloat16, device=self.device)
        self._gemm(GEMM_FWD, EPI_NONE, T, V, H, hidden, H, self.w(head + ".weight"), H, out, ldv, bias=self.p(head + ".bias"))
        return out[:, :V]

    # ------------------------------------------------------------------ loss + backward
    ROW_CHUNK = 16384   # rows of the all-valid token head handled per pass (logits chunk: 16384 x 64000 bf16 = 2.1 GB)

    def _head_chunk(self, hidden, head: str, rows, weights, labels, m: int, accumulate_loss: bool, backward: bool) -> None:
        """One chunk of one head on its own row set: gather -> logits -> CE (loss accumulated into loss_buf) and, with
        ``backward``, bias / weight gradients, dgrad and scatter-add of the rows' gradient into ``d_y``."""
        H = self.cfg.hidden_size
        s = _stream()
        V = self.layout.shapes[head + ".weight"][0]
        ldv = (V + 7) // 8 * 8
        logits = self.logits[:m * ldv].view(m, ldv)
        call("plb_gather_rows", ptr(hidden), H, ptr(rows), m, H, ptr(self.h_rows), H, s)
        self._gemm(GEMM_FWD, EPI_NONE, m, V, H, self.h_rows, H, self.wp[head + ".weight"], H, logits, ldv,
                   bias=self.pp[head + ".bias"])
        call("plb_masked_ce", ptr(logits), ldv, m, V, ptr(labels), ptr(weights), ptr(self.row_loss), ptr(self.loss_buf),
             1 if accumulate_loss else 0, s)
        self.launches += 2
        if backward:
            call("plb_colsum", ptr(logits), ldv, m, V, self.gp[head + ".bias"], s)
            self._head_wgrad(head, V, H, m, logits, ldv)
            self._gemm(GEMM_DGRAD, EPI_NONE, m, H, V, logits, ldv, self.wp[head + ".weight"], H, self.dh_rows, H)
            call("plb_scatter_add_rows", ptr(self.dh_rows), H, ptr(rows), m, H, ptr(self.d_y), H, s)
            self.launches += 2

    def _all_valid_heads(self, b: DeviceBatch, heads: List[str], hidden, loss_started: bool, backward: bool) -> bool:
        """Heads whose loss runs over every valid position (``DeviceBatch.all_valid``), ROW_CHUNK rows at a time."""
        for i, head in enumerate(heads):
            if not (i < len(b.all_valid) and b.all_valid[i]):
                continue
            for lo in range(0, b.M_all, self.ROW_CHUNK):
                m = min(self.ROW_CHUNK, b.M_all - lo)
                self._head_chunk(hidden, head, b.rows_all[lo:lo + m], b.weights_all[lo:lo + m], b.labels[i][lo:lo + m], m,
                                 loss_started, backward)
                loss_started = True
        return loss_started

    def loss_and_backward(self, b: DeviceBatch, heads: List[str]) -> torch.Tensor:
        """Forward, masked-row CE over ``heads`` (b.labels[i] belongs to heads[i]), full backward into the fp32
        gradient arena.  Returns the loss as a 0-dim fp32 device tensor (sum over heads)."""
        T, H, M = b.T, self.cfg.hidden_size, b.M
        s = _stream()
        uses_all = [i < len(b.all_valid) and b.all_valid[i] for i in range(len(heads))]
        masked_heads = [(i, h) for i, h in enumerate(heads) if not uses_all[i]]
        n_all = b.M_all if any(uses_all) else 0
        self.begin_backward()
        with _Range("forward"):
            hidden = self.forward(b, train=True)
        self._ensure_row_workspace(max(M, min(n_all, self.ROW_CHUNK)))
        if (M == 0 or not masked_heads) and n_all == 0:
            self._empty_step()
            return self.loss_buf[0]
        self.grads_clean = False
        self.have_grads = True
        dy = self.d_y
        self._zero(dy[:T])
        loss_started = False

        self._before_head_backward()
        # ---- heads on the masked rows only: one gather, the heads' row gradients summed, one scatter -------------------
        if M > 0 and masked_heads:
            call("plb_gather_rows", ptr(hidden), H, ptr(b.rows), M, H, ptr(self.h_rows), H, s)
            self.launches += 1
            for k, (i, head) in enumerate(masked_heads):
                V = self.layout.shapes[head + ".weight"][0]
                ldv = (V + 7) // 8 * 8
                logits = self.logits[:M * ldv].view(M, ldv)
                self._gemm(GEMM_FWD, EPI_NONE, M, V, H, self.h_rows, H, self.wp[head + ".weight"], H, logits, ldv,
                           bias=self.pp[head + ".bias"])
                call("plb_masked_ce", ptr(logits), ldv, M, V, ptr(b.labels[i]), ptr(b.weights), ptr(self.row_loss),
                     ptr(self.loss_buf), 1 if loss_started else 0, s)
                loss_started = True
                call("plb_colsum", ptr(logits), ldv, M, V, self.gp[head + ".bias"], s)
                self.launches += 3
                self._head_wgrad(head, V, H, M, logits, ldv)
                if k == 0:
                    self._gemm(GEMM_DGRAD, EPI_NONE, M, H, V, logits, ldv, self.wp[head + ".weight"], H, self.dh_rows, H)
                else:
                    self._gemm(GEMM_DGRAD, EPI_RESIDUAL, M, H, V, logits, ldv, self.wp[head + ".weight"], H, self.dh_rows, H,
                               aux_in=self.dh_rows, ld_in=H)
            call("plb_scatter_add_rows", ptr(self.dh_rows), H, ptr(b.rows), M, H, ptr(dy), H, s)
            self.launches += 1
        elif masked_heads:
            self._zero(self.loss_buf)   # masked heads contribute 0 (every sample's mask set is empty)
            loss_started = True
        # ---- heads over every valid position (upstream word loss), chunked ---------------------------------------------
        if n_all:
            self._all_valid_heads(b, heads, hidden, loss_started, backward=True)
        self._after_head_backward()
        with _Range("backward"):
            self.encoder_backward(b)
        return self.loss_buf[0]

    def _empty_step(self) -> None:
        """R/train.py:129: no masked rows anywhere -> loss 0 and no gradient reaches any parameter (a single process then
        skips the optimizer step, like torch does for ``grad is None``).  Under data parallelism the peers are (or will be)
        inside the bucket all-reduces and will run AdamW on the averaged gradient, so this rank joins both collectives
        with its all-zero contribution and steps like everybody else (DDP semantics, R/train.py:160,218-221) — returning
        early would hang the other ranks and let the replicas drift apart."""
        self.loss_buf.zero_()   # (a torch fill: this path also runs on the CPU stand-ins of tests/test_ddp_gloo.py)
        if not self.always_reduce:
            self.have_grads = False
            return
        self.have_grads = True
        self._before_head_backward()
        self._after_head_backward()
        self._after_encoder_backward()

    def loss_only(self, b: DeviceBatch, heads: List[str]) -> torch.Tensor:
        """Forward + CE without the backward pass (validation, R/train.py:288-305 under ``no_grad``).
        Uses the two-set activation ping-pong of the inference forward; leaves gradients and optimizer state alone."""
        c = self.cfg
        H, M = c.hidden_size, b.M
        s = _stream()
        uses_all = [i < len(b.all_valid) and b.all_valid[i] for i in range(len(heads))]
        masked_heads = [(i, h) for i, h in enumerate(heads) if not uses_all[i]]
        n_all = b.M_all if any(uses_all) else 0
        hidden = self.forward(b, train=False)
        self.loss_buf.zero_()
        self._ensure_row_workspace(max(M, min(n_all, self.ROW_CHUNK)))
        if M > 0 and masked_heads:
            call("plb_gather_rows", ptr(hidden), H, ptr(b.rows), M, H, ptr(self.h_rows), H, s)
            self.launches += 1
            for i, head in masked_heads:
                V = self.layout.shapes[head + ".weight"][0]
                ldv = (V + 7) // 8 * 8
                logits = self.logits[:M * ldv].view(M, ldv)
                self._gemm(GEMM_FWD, EPI_NONE, M, V, H, self.h_rows, H, self.wp[head + ".weight"], H, logits, ldv,
                           bias=self.pp[head + ".bias"])
                call("plb_masked_ce", ptr(logits), ldv, M, V, ptr(b.labels[i]), ptr(b.weights), ptr(self.row_loss),
                     ptr(self.loss_buf), 1, s)
                self.launches += 1
        if n_all:
            self._all_valid_heads(b, heads, hidden, True, backward=False)
        return self.loss_buf[0].clone()

    def begin_backward(self) -> None:
        """Gradients accumulate in place (12 applications of the shared layer), so the arena must start at zero."""
        if not self.grads_clean:
            self._zero(self.grads[:self.layout.trainable_end])
            self.grads_clean = True

    def encoder_backward(self, b: DeviceBatch) -> None:
        """Backward of ``forward(b, train=True)`` given d(last_hidden_state) in ``self.d_y`` (bf16 [T, H])."""
        c = self.cfg
        T, H, I, E = b.T, c.hidden_size, c.intermediate_size, c.embedding_size
        nh = c.num_attention_heads
        s = _stream()
        att = _LAYER + "attention."
        dy = self.d_y
        self.grads_clean = False
        self.have_grads = True
        # ---- 12 applications of the shared layer, last to first ----------------------------------
        wqkv = self.wp[att + "query.weight"]                                        # q|k|v adjacent: [3H, H]
        g_wqkv, g_bqkv = self.gp[att + "query.weight"], self.gp[att + "query.bias"]
        delta_flat = self.delta.view(-1)   # indexed [head * T + row] by the kernels
        attn_ws = self._attn_workspace(b.B)
        p_hid, p_att = self._drop_p   # what the forward whose activations are live applied
        seed, base = self.dropout_seed, self._drop_base
        for l in reversed(range(c.num_hidden_layers)):
            A = self.acts[l]
            x_in = self.acts[l - 1]["y"] if l > 0 else self.x0
            # y = LN2(z2), z2 = a + g W2^T + b2
            call("plb_ln_bwd", ptr(dy), ptr(A["z2"]), ptr(A["st2"]), self.pp[_LAYER + "full_layer_layer_norm.weight"], T, H,
                 ptr(self.d_z), self.gp[_LAYER + "full_layer_layer_norm.weight"],
                 self.gp[_LAYER + "full_layer_layer_norm.bias"], self.gp[_LAYER + "ffn_output.bias"], s)
            self._gemm(GEMM_WGRAD, EPI_NONE, H, I, T, self.d_z, H, A["g"], I, self.gp[_LAYER + "ffn_output.weight"], I)
            if self.fuse_ffn1_bias_sums:   # d_u = (d_z W2) o gelu'(u); b1.grad += colsum(d_u) in the same kernel (its store warp)
                self._gemm(GEMM_DGRAD, EPI_GELU_BWD, T, I, H, self.d_z, H, self.wp[_LAYER + "ffn_output.weight"], I, self.d_u, I,
                           aux_in=A["dg"], ld_in=I, aux_out=self.gp[_LAYER + "ffn.bias"])
            else:
                self._gemm(GEMM_DGRAD, EPI_GELU_BWD, T, I, H, self.d_z, H, self.wp[_LAYER + "ffn_output.weight"], I, self.d_u, I,
                           aux_in=A["dg"], ld_in=I)
                call("plb_colsum", ptr(self.d_u), I, T, I, self.gp[_LAYER + "ffn.bias"], s)
                self.launches += 1
            self._gemm(GEMM_WGRAD, EPI_NONE, I, H, T, self.d_u, I, A["a"], H, self.gp[_LAYER + "ffn.weight"], H)
            self._gemm(GEMM_DGRAD, EPI_RESIDUAL, T, H, I, self.d_u, I, self.wp[_LAYER + "ffn.weight"], H, self.d_a, H,
                       aux_in=self.d_z, ld_in=H)
            # a = LN1(z1), z1 = x + ctx Wo^T + bo
            d_proj = self.d_z   # gradient of the output projection's result; the residual path takes d_z itself
            if p_hid > 0:
                # z1 = x + dropout(proj): d proj = d_z o m (same mask), and the projection's bias gradient is ITS column sum
                call("plb_ln_bwd", ptr(self.d_a), ptr(A["z1"]), ptr(A["st1"]), self.pp[att + "LayerNorm.weight"], T, H,
                     ptr(self.d_z), self.gp[att + "LayerNorm.weight"], self.gp[att + "LayerNorm.bias"], 0, s)
                d_proj = self.d_a   # (consumed by the LayerNorm backward above)
                call("plb_dropout", ptr(self.d_z), H, 0, 0, ptr(d_proj), H, T, H, p_hid, seed, (base + 2 + 2 * l) & 0xFFFFFFFF, s)
                call("plb_colsum", ptr(d_proj), H, T, H, self.gp[att + "dense.bias"], s)
                self.launches += 2
            else:
                call("plb_ln_bwd", ptr(self.d_a), ptr(A["z1"]), ptr(A["st1"]), self.pp[att + "LayerNorm.weight"], T, H,
                     ptr(self.d_z), self.gp[att + "LayerNorm.weight"], self.gp[att + "LayerNorm.bias"],
                     self.gp[att + "dense.bias"], s)
            self._gemm(GEMM_WGRAD, EPI_NONE, H, H, T, d_proj, H, A["ctx"], H, self.gp[att + "dense.weight"], H)
            # dO = dz Wo, and in the same epilogue attention's delta[h, row] = sum_d dO o O over each head's 64 columns
            self._zero(delta_flat[:nh * T])
            self._gemm(GEMM_DGRAD, EPI_ROWDOT, T, H, H, d_proj, H, self.wp[att + "dense.weight"], H, self.d_ctx, H,
                       aux_in=A["ctx"], ld_in=H, aux_out=self.delta, ld_out=T)
            # one kernel: S^T / dP^T and the exponentials are computed once for dQ, dK and dV (attn_bwd_fused.cu).  ctx = NULL:
            # delta is given; the same kernel accumulates the QKV bias gradient (column sums of dqkv)
            if p_att > 0:
                call("plb_attn_bwd_fused_dropout", ptr(A["qkv"]), 3 * H, ptr(b.seq_start), ptr(b.q_len), ptr(b.kv_len), b.B, nh,
                     H // nh, b.L, T, 0, ptr(self.d_ctx), H, ptr(A["lse"]), ptr(self.delta), ptr(self.d_qkv), 3 * H,
                     ptr(g_bqkv), ptr(attn_ws), attn_ws.numel() * 4, p_att, seed, (base + 1 + 2 * l) & 0xFFFFFFFF, s)
            else:
                call("plb_attn_bwd_fused", ptr(A["qkv"]), 3 * H, ptr(b.seq_start), ptr(b.q_len), ptr(b.kv_len), b.B, nh, H // nh, b.L,
                     T, 0, ptr(self.d_ctx), H, ptr(A["lse"]), ptr(self.delta), ptr(self.d_qkv), 3 * H,
                     ptr(g_bqkv), ptr(attn_ws), attn_ws.numel() * 4, s)
            self._gemm(GEMM_WGRAD, EPI_NONE, 3 * H, H, T, self.d_qkv, 3 * H, x_in, H, g_wqkv, H)
            self._gemm(GEMM_DGRAD, EPI_RESIDUAL, T, H, 3 * H, self.d_qkv, 3 * H, wqkv, H, dy, H, aux_in=self.d_z, ld_in=H)
            self.launches += 3  # 2 ln_bwd, 1 attention backward; the delta memset is torch's

        # ---- embedding side -----------------------------------------------------------------------
        mp = "encoder.encoder.embedding_hidden_mapping_in."
        call("plb_colsum", ptr(dy), H, T, H, self.gp[mp + "bias"], s)
        self._gemm(GEMM_WGRAD, EPI_NONE, H, E, T, dy, H, self.e, E, self.gp[mp + "weight"], E)
        self._gemm(GEMM_DGRAD, EPI_NONE, T, E, H, dy, H, self.wp[mp + "weight"], E, self.d_e, E)
        if p_hid > 0:   # e = dropout(LN(...)): the LayerNorm backward sees d_e o m
            call("plb_dropout", ptr(self.d_e), E, 0, 0, ptr(self.d_e), E, T, E, p_hid, seed, base, s)
            self.launches += 1
        call("plb_embed_ln_bwd", ptr(self.d_e), ptr(b.ids), ptr(b.pos), T, E, c.vocab_size, self.pp[_EMB + "word_embeddings.weight"],
             self.pp[_EMB + "position_embeddings.weight"], self.pp[_EMB + "token_type_embeddings.weight"],
             self.pp[_EMB + "LayerNorm.weight"], ptr(self.e_stats), c.pad_token_id,
             self.gp[_EMB + "word_embeddings.weight"], self.gp[_EMB + "position_embeddings.weight"],
             self.gp[_EMB + "token_type_embeddings.weight"], self.gp[_EMB + "LayerNorm.weight"],
             self.gp[_EMB + "LayerNorm.bias"], s)
        self.launches += 2
        self._after_encoder_backward()

    # ------------------------------------------------------------------ data parallel hooks (ddp.py overrides)
    def _before_head_backward(self) -> None:   # every step, every rank, before the first head gradient is written
        pass

    def _after_head_backward(self) -> None:
        pass

    def _after_encoder_backward(self) -> None:
        pass

    # ------------------------------------------------------------------ optimizer
    def adamw_step(self, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01,
                   grad_scale: float = 1.0) -> None:
        """One fused AdamW update of every trainable parameter (R/train.py:357); also rewrites the bf16 shadow and
        zeroes the gradient arena for the next step's in-place accumulation."""
        if not self.have_grads:
            return  # reference: parameters whose .grad is None are skipped entirely (no decay either)
        self.opt_step += 1
        n = self.layout.trainable_end
        if _NVTX:
            torch.cuda.nvtx.mark("adamw")
        call("plb_adamw_step", ptr(self.params), ptr(self.grads), ptr(self.exp_avg), ptr(self.exp_avg_sq), ptr(self.shadow),
             n, self.opt_step, lr, betas[0], betas[1], eps, weight_decay, grad_scale, 1, _stream())
        self.launches += 1
        self.grads_clean = True
        self.have_grads = False
