"""Drop-in ``MultiTaskModel`` / ``PhonemeOnlyModel`` (R/model.py:5-30) backed by the B200 engine.

Same constructors, same attribute names (``encoder``, ``phoneme_predictor``, ``token_predictor``), same
``state_dict`` keys/shapes/dtypes, same ``forward(phonemes, attention_mask=None)`` contract — so
``AlbertConfig(**model_params)`` → ``AlbertModel`` → ``MultiTaskModel(...)``, ``load_state_dict`` of a reference
checkpoint, ``model.encoder.save_pretrained`` (R/convert_to_hf.py:49) and the StyleTTS loader keep working.

Two ways to train:

* reference-style (unchanged caller code, R/train.py:381-390 + :355-357)::

      pred = model(masked_phonemes, attention_mask=mask)          # full [B, L, V] logits, autograd-enabled
      loss = calculate_phoneme_loss(pred, labels, lengths, masked_indices, criterion)
      loss.backward(); optimizer.step()

  ``forward`` runs the fused encoder (no [B,H,L,L] tensor) and a custom autograd node runs the fused backward.

* fused step (what ``bench.py`` measures; what ``plbert_b200.train.process_batch`` uses)::

      loss = model.loss_and_backward(batch)                       # collate tuple in, device scalar out
      optimizer.step()

  which evaluates the heads and the cross-entropy on the masked rows only and never materialises logits for the
  other positions.

There is no CPU/PyTorch fallback: calling the model with CPU tensors raises.
"""
from __future__ import annotations

import weakref
from types import SimpleNamespace
from typing import List, Optional, Sequence

import torch
from torch import nn

from . import _abi
from ._abi import EPI_NONE, EPI_RESIDUAL, GEMM_DGRAD, GEMM_FWD, GEMM_WGRAD, call, ptr
from .batch import BatchStager, DeviceBatch
from .engine import Engine, EngineConfig, _stream

# id(parameter) -> engine that owns its storage (used by FusedAdamW to find the arenas)
_ENGINE_OF_PARAM: "weakref.WeakValueDictionary[int, Engine]" = weakref.WeakValueDictionary()


def engine_of(param: torch.Tensor) -> Optional[Engine]:
    return _ENGINE_OF_PARAM.get(id(param))


class _FullLogitsFn(torch.autograd.Function):
    """forward: fused encoder + all-position head GEMMs; backward: fused head/encoder backward.
    The parameters are passed only so autograd routes gradients to them."""

    @staticmethod
    def forward(ctx, model, batch: DeviceBatch, heads, need_grad, *params):
        eng: Engine = model._engine
        hidden = eng.forward(batch, train=need_grad)
        outs = []
        for h in heads:
            outs.append(eng.head_logits(hidden, h).float().view(batch.B, batch.L, -1))
        ctx.model, ctx.batch, ctx.heads = model, batch, heads
        ctx.hidden = hidden
        ctx.n_params = len(params)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *dlogits):
        model, b, heads = ctx.model, ctx.batch, ctx.heads
        eng: Engine = model._engine
        H, T = eng.cfg.hidden_size, b.T
        eng.begin_backward()
        eng.grads_clean = False
        first = True
        for h, dl in zip(heads, dlogits):
            if dl is None:
                continue
            V = dl.shape[-1]
            ldv = (V + 7) // 8 * 8
            dz = torch.zeros((T, ldv), dtype=torch.bfloat16, device=dl.device)
            dz[:, :V] = dl.reshape(T, V)
            call("plb_colsum", ptr(dz), ldv, T, V, ptr(eng.g(h + ".bias")), _stream())
            eng.launches += 1
            eng._gemm(GEMM_WGRAD, EPI_NONE, V, H, T, dz, ldv, ctx.hidden, H, eng.g(h + ".weight"), H)
            if first:
                eng._gemm(GEMM_DGRAD, EPI_NONE, T, H, V, dz, ldv, eng.w(h + ".weight"), H, eng.d_y, H)
            else:
                eng._gemm(GEMM_DGRAD, EPI_RESIDUAL, T, H, V, dz, ldv, eng.w(h + ".weight"), H, eng.d_y, H,
                          aux_in=eng.d_y, ld_in=H)
            first = False
        if first:
            return (None, None, None, None) + (None,) * ctx.n_params
        eng._after_head_backward()
        eng.encoder_backward(b)
        grads = []
        for name, p in model._arena_params():
            if not p.requires_grad or name in eng.layout.frozen:
                grads.append(None)
                continue
            view = eng.grad_views[name]
            # p.grad may already BE this arena view (published by the fused path, or kept by zero_grad(set_to_none=False)):
            # autograd would then add the returned view onto itself in place and double the gradient.  Dropping the alias
            # makes AccumulateGrad store (a copy of) the fresh gradient instead.
            if p.grad is not None and p.grad.data_ptr() == view.data_ptr():
                p.grad = None
            grads.append(view)
        return (None, None, None, None) + tuple(grads)


class _EncoderFn(torch.autograd.Function):
    """``encoder(ids, attention_mask=...).last_hidden_state`` with autograd (StyleTTS finetunes PL-BERT through exactly this
    call, R/README.md:91): forward = fused encoder, backward = fused encoder backward given d(last_hidden_state)."""

    @staticmethod
    def forward(ctx, model, batch: DeviceBatch, out_dtype, *params):
        eng: Engine = model._engine
        hidden = eng.forward(batch, train=True)
        ctx.model, ctx.batch, ctx.n_params = model, batch, len(params)
        return hidden.to(out_dtype).view(batch.B, batch.L, -1)   # a copy: the workspace is reused by the next call

    @staticmethod
    def backward(ctx, dhidden):
        model, b = ctx.model, ctx.batch
        eng: Engine = model._engine
        eng.begin_backward()
        eng.d_y[:b.T].copy_(dhidden.reshape(b.T, -1))               # fp32 -> bf16 on the device
        eng._after_head_backward()                                   # no head gradients on this path: bucket 0 is zeros
        eng.encoder_backward(b)
        grads = []
        for name, p in model._arena_params():
            if not p.requires_grad or name in eng.layout.frozen or not name.startswith("encoder."):
                grads.append(None)
                continue
            view = eng.grad_views[name]
            if p.grad is not None and p.grad.data_ptr() == view.data_ptr():
                p.grad = None                                        # see _FullLogitsFn.backward
            grads.append(view)
        return (None, None, None) + tuple(grads)


class _EncoderGraph:
    """One captured encoder forward for a fixed [B, L] shape (padded layout).  Inputs live in static device buffers
    (ids, positions, lengths) that every call overwrites with device-side copies; the graph reads the engine's bf16
    weight shadow in place, so weight updates only need the usual shadow refresh before the replay."""

    def __init__(self, model: "_FusedModel", B: int, L: int):
        eng: Engine = model._engine
        dev = eng.device
        self.model, self.B, self.L = model, B, L
        i32 = dict(dtype=torch.int32, device=dev)
        self.ids = torch.zeros(B * L, **i32)
        self.kv_len = torch.full((B,), L, **i32)
        self.batch = DeviceBatch(B=B, L=L, ids=self.ids, pos=torch.arange(L, **i32).repeat(B),
                                 seq_start=torch.arange(B, **i32) * L, q_len=torch.full((B,), L, **i32), kv_len=self.kv_len,
                                 num_valid_tokens=B * L, n_rows=B * L)
        with torch.no_grad():
            eng.forward(self.batch, train=False)          # warm-up: allocates the workspace, configures the kernels
            torch.cuda.synchronize(dev)
            # the graph holds raw pointers into these buffers: keep them alive even if the engine regrows its workspace
            self._keep = (eng.e, eng.e_stats, eng.x0, list(eng.acts[:2]))
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = eng.forward(self.batch, train=False)

    def __call__(self, phonemes: torch.Tensor, attention_mask: Optional[torch.Tensor]) -> torch.Tensor:
        eng: Engine = self.model._engine
        if not eng.shadow_valid:
            eng.refresh_shadow()
        self.ids.copy_(phonemes.reshape(-1), non_blocking=True)          # int64 -> int32 on the device (or pinned H2D)
        if attention_mask is None:
            self.kv_len.fill_(self.L)
        else:
            self.kv_len.copy_(attention_mask.to(eng.device, non_blocking=True).ne(0).sum(dim=1))
        self.graph.replay()
        return self.out.view(self.B, self.L, -1).clone()


class _FusedModel(nn.Module):
    _heads: List[str] = []

    pack_sequences = True  # loss_and_backward lays variable-length batches out without padding rows (cu_seqlens style)

    def _init_fused(self):
        self._engine: Optional[Engine] = None
        self._stager: Optional[BatchStager] = None
        self._versions: List[int] = []
        self._graphs = {}
        self._adopted = None   # [(parameter, arena data_ptr)] once every parameter is a view into the engine's arena
        self._named = None     # cached (name, parameter) list for the per-step gradient publication

    # ------------------------------------------------------------------ engine / arena management
    def _arena_params(self):
        """(state_dict name, parameter) for every parameter, in ``self.parameters()`` order."""
        return list(self.named_parameters())

    def _num_tokens(self) -> Optional[int]:
        tp = getattr(self, "token_predictor", None)
        return None if tp is None else tp.out_features

    def _ensure_engine(self, device: torch.device) -> Engine:
        if device.type != "cuda":
            raise _abi.PlbError("plbert_b200 computes on CUDA only; move the model and inputs to a B200 (no CPU fallback)")
        eng = self._engine
        # fast path (every step): same device, every parameter still IS its arena view, nobody wrote through torch.
        # This runs between the loss read-back of one step and the first launch of the next, i.e. with the GPU idle.
        if eng is not None and eng.device == device and self._adopted is not None:
            if all(p.data_ptr() == q for p, q in self._adopted):
                versions = [p._version for p, _ in self._adopted]
                if versions != self._versions:
                    eng.shadow_valid = False
                    self._versions = versions
                eng.dropout_enabled = self.training
                return eng
        named = self._arena_params()
        if eng is None or eng.device != device:
            cfg = EngineConfig.from_albert(self.encoder.config, self.phoneme_predictor.out_features, self._num_tokens())
            eng = Engine(cfg, device)
            missing = set(eng.layout.offsets) ^ {n for n, _ in named}
            if missing:
                raise _abi.PlbError(f"parameter set does not match the PL-BERT layout: {sorted(missing)}")
            self._engine = eng
            self._stager = BatchStager(device)
            self._versions = []
            self._graphs = {}
        # (re)adopt: parameters whose storage is not the arena view (first call, .to(), load with assign=True ...)
        stale = False
        adopted = []
        for name, p in named:
            view = eng.p(name)
            if p.data_ptr() != view.data_ptr() or p.dtype != torch.float32:
                if tuple(p.shape) != tuple(view.shape):
                    raise _abi.PlbError(f"{name}: shape {tuple(p.shape)} does not match the layout {tuple(view.shape)}")
                with torch.no_grad():
                    view.copy_(p.detach().to(device=device, dtype=torch.float32))
                p.data = view
                _ENGINE_OF_PARAM[id(p)] = eng
                stale = True
            adopted.append((p, view.data_ptr()))
        self._adopted = adopted
        self._named = named
        versions = [p._version for _, p in named]
        if stale or versions != self._versions:
            eng.shadow_valid = False  # someone wrote the fp32 weights through torch (load_state_dict, init, ...)
            self._versions = versions
        eng.dropout_enabled = self.training
        return eng

    def _publish_grads(self) -> None:
        eng = self._engine
        if self._named is None:
            self._named = self._arena_params()
        for name, p in self._named:
            if name in eng.layout.frozen or not p.requires_grad:
                continue
            p.grad = eng.grad_views[name] if eng.have_grads else None

    # ------------------------------------------------------------------ reference-compatible forward
    def _stage_inputs(self, phonemes: torch.Tensor, attention_mask: Optional[torch.Tensor]) -> DeviceBatch:
        if phonemes.dim() != 2:
            raise ValueError("phonemes must be [B, L]")
        B, L = phonemes.shape
        if attention_mask is None:
            lengths = [L] * B
        else:
            am = attention_mask.to("cpu") != 0
            lengths = am.sum(dim=1).tolist()
            # the reference only ever builds prefix masks (R/train.py:34-44); anything else is not this path
            if not torch.equal(am, torch.arange(L)[None, :] < torch.tensor(lengths)[:, None]):
                raise _abi.PlbError("attention_mask must be a right-padded prefix mask (R/train.py:34-44)")
        return self._stager.stage(phonemes, lengths)

    def _forward_logits(self, phonemes, attention_mask):
        self._ensure_engine(phonemes.device)
        batch = self._stage_inputs(phonemes, attention_mask)
        params = [p for _, p in self._arena_params()]
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _FullLogitsFn.apply(self, batch, list(self._heads), need_grad, *params)

    def encode(self, phonemes: torch.Tensor, attention_mask: Optional[torch.Tensor] = None, *,
               use_graph: bool = False) -> torch.Tensor:
        """Encoder-only forward (``AlbertModel(...).last_hidden_state``, R/README.md:91): bf16 [B, L, H], no grad.

        ``use_graph=True`` replays a CUDA graph of the ~90 kernel launches captured once per (B, L)
        shape: no per-call host staging, no device->host sync (the eager path reads the mask back to validate it), so
        short inference batches are not launch-bound.  The mask must still be a right-padded prefix mask."""
        eng = self._ensure_engine(phonemes.device)
        if use_graph:
            key = (int(phonemes.shape[0]), int(phonemes.shape[1]))
            g = self._graphs.get(key)
            if g is None:
                g = self._graphs[key] = _EncoderGraph(self, *key)
            return g(phonemes, attention_mask)
        batch = self._stage_inputs(phonemes, attention_mask)
        with torch.no_grad():
            hidden = eng.forward(batch, train=False)
        return hidden.view(batch.B, batch.L, -1).clone()  # the workspace is reused by the next call

    # ------------------------------------------------------------------ fused training step
    def new_stager(self) -> BatchStager:
        """A private pair of pinned/device staging slots (``BatchFeeder`` owns one, so a batch it staged ahead survives
        any number of interleaved ``loss`` / ``encode`` / ``loss_and_backward`` calls that stage through the model's own)."""
        return BatchStager(next(self.parameters()).device)

    def stage_batch(self, batch: Sequence, stager: Optional[BatchStager] = None) -> DeviceBatch:
        """Host half of ``loss_and_backward``: flatten the collate tuple into the integer arrays the kernels read, fill one
        pinned buffer and enqueue its H2D copy (asynchronous).  Calling it for batch i+1 right after the kernels of step i
        have been enqueued hides the host work behind the GPU (``plbert_b200.batch.BatchFeeder`` does exactly that)."""
        if len(batch) == 5:
            token_ids, labels, masked, lengths, midx = batch
        else:
            labels, masked, lengths, midx = batch
            token_ids = None
        device = next(self.parameters()).device
        self._ensure_engine(device)
        label_tensors = [labels]
        heads = ["phoneme_predictor"]
        if "token_predictor" in self._heads:
            if token_ids is None:
                raise _abi.PlbError("MultiTaskModel.loss_and_backward needs the 5-tuple batch with token_ids")
            label_tensors.append(token_ids)
            heads.append("token_predictor")
        # drop the padding rows when there are any: same loss and gradients, fewer tokens to process
        packed = self.pack_sequences and min(int(x) for x in lengths) < masked.shape[1]
        # row set of the token head: the masked rows (default, the phoneme-loss template of R/train.py:107-131) or every
        # valid position (upstream's word loss)
        all_valid = [1] if (len(heads) > 1 and getattr(self, "token_rows", "masked") == "all_valid") else []
        db = (stager or self._stager).stage(masked, lengths, midx, label_tensors, packed=packed, all_valid_heads=all_valid)
        db.heads = heads
        return db

    def loss_and_backward(self, batch) -> torch.Tensor:
        """R/train.py:381-390 ``process_batch`` + ``loss.backward()`` in one call.

        ``batch`` is the collate tuple: ``(phoneme_labels, masked_phonemes, input_lengths, masked_indices)`` or, with
        token ids, ``(token_ids, phoneme_labels, masked_phonemes, input_lengths, masked_indices)`` — or a ``DeviceBatch``
        already staged by ``stage_batch``.  Returns the loss (mean over non-empty samples of the per-sample mean CE over
        masked rows; heads summed) as a device scalar and leaves the gradients in ``p.grad``."""
        db = batch if isinstance(batch, DeviceBatch) else self.stage_batch(batch)
        if not db.heads:
            raise _abi.PlbError("loss_and_backward needs a batch staged with labels (stage_batch)")
        eng = self._ensure_engine(db.ids.device)
        self.last_batch = db
        loss = eng.loss_and_backward(db, db.heads)
        self._publish_grads()
        return loss

    def loss(self, batch) -> torch.Tensor:
        """Loss of one batch without gradients — ``process_batch`` under ``torch.no_grad()`` as the reference's
        ``validate`` runs it (R/train.py:288-305).  Same batch forms as ``loss_and_backward``."""
        db = batch if isinstance(batch, DeviceBatch) else self.stage_batch(batch)
        if not db.heads:
            raise _abi.PlbError("loss needs a batch staged with labels (stage_batch)")
        eng = self._ensure_engine(db.ids.device)
        return eng.loss_only(db, db.heads)


class MultiTaskModel(_FusedModel):
    """R/model.py:5-18."""

    _heads = ["phoneme_predictor", "token_predictor"]
    # rows the token-head loss runs over: "masked" (same rows as the phoneme loss) or "all_valid" (every valid position,
    # per-sample mean — upstream yl4579/PL-BERT's word loss).  The reference fork has no token loss at all.
    token_rows = "masked"

    def __init__(self, model, num_phonemes, num_tokens, hidden_size, _fuse=True):
        super().__init__()
        self.encoder = model
        self.phoneme_predictor = nn.Linear(hidden_size, num_phonemes)
        self.token_predictor = nn.Linear(hidden_size, num_tokens)
        self._init_fused()
        if _fuse:
            fuse_albert(model, owner=self)   # ``self.encoder(ids, attention_mask=...)`` (R/model.py:14) runs the kernels too

    def forward(self, phonemes, attention_mask=None):
        phoneme_pred, token_pred = self._forward_logits(phonemes, attention_mask)
        return phoneme_pred, token_pred


class PhonemeOnlyModel(_FusedModel):
    """R/model.py:19-30."""

    _heads = ["phoneme_predictor"]

    def __init__(self, model, num_phonemes, hidden_size, _fuse=True):
        super().__init__()
        self.encoder = model
        self.phoneme_predictor = nn.Linear(hidden_size, num_phonemes)
        self._init_fused()
        if _fuse:
            fuse_albert(model, owner=self)

    def forward(self, phonemes, attention_mask=None):
        (phoneme_pred,) = self._forward_logits(phonemes, attention_mask)
        return phoneme_pred


def fuse_albert(albert, owner: Optional[_FusedModel] = None, output_dtype: torch.dtype = torch.float32,
                use_graph: bool = False):
    """Make ``albert(ids, attention_mask=mask).last_hidden_state`` — the call of R/model.py:14-15 and of the StyleTTS
    snippet (R/README.md:91,107) — run the sm_100a kernels.  ``albert`` is the caller's HF ``AlbertModel`` and stays the
    parameter container (``state_dict`` names, ``save_pretrained``, ``from_pretrained``, ``config`` are untouched); only its
    ``forward`` is replaced, on this instance.  The stock implementation remains reachable as ``albert.hf_forward``.

    ``owner``: the ``MultiTaskModel`` / ``PhonemeOnlyModel`` whose engine holds the weights (the constructors pass
    themselves); a stand-alone encoder (StyleTTS: ``bert = fuse_albert(AlbertModel(cfg))``) gets a private engine.
    ``output_dtype``: dtype of ``last_hidden_state`` (HF returns fp32 in an fp32 model; the kernels compute in bf16).
    ``use_graph``: replay a captured CUDA graph for no-grad calls (fixed [B, L]; the mask must be a right-padded prefix
    mask and is NOT validated on that path — no device->host sync).  With gradients enabled the call is differentiable
    (fused backward into the encoder parameters).  CPU tensors raise ``PlbError``: there is no fallback.
    ``pooler_output`` is ``None`` (PL-BERT never reads it: R/model.py:14-15)."""
    from transformers.modeling_outputs import BaseModelOutputWithPooling
    if getattr(albert, "_plb_fused", False):
        if owner is not None:
            object.__setattr__(albert, "_plb_owner", owner)
        return albert
    if owner is None:
        # private engine: the layout needs a phoneme head; this 8-row one never leaves the wrapper (not in albert's state_dict)
        owner = PhonemeOnlyModel(albert, 8, albert.config.hidden_size, _fuse=False)
    object.__setattr__(albert, "_plb_owner", owner)      # plain attribute: no module registration, no state_dict entry
    object.__setattr__(albert, "hf_forward", albert.forward)
    albert.plb_output_dtype, albert.plb_use_graph = output_dtype, use_graph

    def forward(input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, inputs_embeds=None, **kwargs):
        model: _FusedModel = albert._plb_owner
        if input_ids is None or inputs_embeds is not None or position_ids is not None:
            raise _abi.PlbError("fused AlbertModel.forward takes input_ids (+ attention_mask) only, like the reference's calls")
        if token_type_ids is not None and bool(token_type_ids.ne(0).any()):
            raise _abi.PlbError("fused AlbertModel.forward: token_type_ids must be all zeros (the reference never passes any)")
        model._ensure_engine(input_ids.device)
        params = [p for _, p in model._arena_params()]
        if torch.is_grad_enabled() and any(p.requires_grad for n, p in model._arena_params() if n.startswith("encoder.")):
            batch = model._stage_inputs(input_ids, attention_mask)
            hidden = _EncoderFn.apply(model, batch, albert.plb_output_dtype, *params)
        else:
            hidden = model.encode(input_ids, attention_mask, use_graph=albert.plb_use_graph).to(albert.plb_output_dtype)
        return BaseModelOutputWithPooling(last_hidden_state=hidden, pooler_output=None)

    object.__setattr__(albert, "forward", forward)
    object.__setattr__(albert, "_plb_fused", True)
    return albert


FusedAlbertEncoder = fuse_albert   # the name VERDICT round 1 asked for: FusedAlbertEncoder(albert)(ids, attention_mask=m)


# upstream (yl4579/PL-BERT) checkpoints call the heads mask_predictor / word_predictor
UPSTREAM_ALIASES = {"mask_predictor.": "phoneme_predictor.", "word_predictor.": "token_predictor."}


def normalize_state_dict_keys(state_dict):
    """Strip the DDP ``module.`` prefix (R/train.py:98, R/convert_to_hf.py:40) and map upstream head names."""
    out = {}
    for k, v in state_dict.items():
        k = k.replace("module.", "")
        for old, new in UPSTREAM_ALIASES.items():
            if k.startswith(old):
                k = new + k[len(old):]
        out[k] = v
    return out
