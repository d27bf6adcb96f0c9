"""``FusedAdamW`` — drop-in for ``torch.optim.AdamW(bert.parameters(), lr=...)`` (R/train.py:272).

One kernel launch updates every trainable parameter in the engine's flat arena (28 B/param + bf16 shadow), instead
of torch's ~10 foreach passes.  ``state_dict()`` / ``load_state_dict()`` keep torch's layout —
``{'state': {idx: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [...]}`` with no entries for parameters that
never received a gradient (the pooler) — so reference checkpoints (R/train.py:416-421) round-trip.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _abi
from .model import engine_of


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameter")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None,
                        decoupled_weight_decay=True)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise _abi.PlbError("FusedAdamW supports the reference's single parameter group (R/train.py:272)")
        self.grad_scale = 1.0  # set to 1/world_size by the data-parallel wrapper

    def _engine(self):
        eng = None
        for p in self.param_groups[0]["params"]:
            e = engine_of(p)
            if e is None:
                raise _abi.PlbError("FusedAdamW: parameter storage is not an engine arena yet — run one forward / "
                                    "loss_and_backward of the plbert_b200 model first (no PyTorch fallback)")
            if eng is not None and e is not eng:
                raise _abi.PlbError("FusedAdamW: parameters belong to different engines")
            eng = e
        return eng

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        eng = self._engine()
        group = self.param_groups[0]
        names = eng.name_of_ptr
        any_grad = False
        for p in group["params"]:
            if p.grad is None:
                continue
            name = names.get(p.data_ptr())
            if name is None or name in eng.layout.frozen:
                raise _abi.PlbError("FusedAdamW: gradient on a parameter outside the trainable arena")
            any_grad = True
            gview = eng.grad_views[name]
            if p.grad is not gview and p.grad.data_ptr() != gview.data_ptr():
                gview.copy_(p.grad)  # autograd (or the user, e.g. clipping) produced a separate tensor
                eng.grads_clean = False
            st = self.state[p]
            if not st:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = eng.view(eng.exp_avg, name)
                st["exp_avg_sq"] = eng.view(eng.exp_avg_sq, name)
            st["step"] += 1
        if not any_grad:
            return loss
        eng.have_grads = True
        eng.adamw_step(float(group["lr"]), tuple(group["betas"]), float(group["eps"]), float(group["weight_decay"]),
                       self.grad_scale)
        return loss

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # torch deep-copies the state tensors; move them into the arenas and point the state at the views again
        eng: Optional[object] = None
        try:
            eng = self._engine()
        except _abi.PlbError:
            self._pending_state = True  # adopted lazily by sync_state() once the engine exists
            return
        self.sync_state(eng)

    def sync_state(self, eng=None):
        eng = eng or self._engine()
        names = {eng.p(n).data_ptr(): n for n in eng.layout.offsets}
        steps = set()
        for p in self.param_groups[0]["params"]:
            st = self.state.get(p)
            if not st:
                continue
            name = names[p.data_ptr()]
            for key, arena in (("exp_avg", eng.exp_avg), ("exp_avg_sq", eng.exp_avg_sq)):
                view = eng.view(arena, name)
                if st[key].data_ptr() != view.data_ptr():
                    view.copy_(st[key].to(view.device, torch.float32))
                    st[key] = view
            st["step"] = torch.as_tensor(float(st["step"]), dtype=torch.float32)
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise _abi.PlbError(f"FusedAdamW: per-parameter step counts differ ({sorted(steps)}); the fused kernel "
                                "keeps one step for the single parameter group")
        if steps:
            eng.opt_step = steps.pop()
        self._pending_state = False
