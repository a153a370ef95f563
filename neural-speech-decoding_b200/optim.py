"""Fused optimizer step for the decoder's training loop ("Adam lr 1e-3 in torch").

``FusedAdam`` applies ``torch.optim.Adam``'s update (no amsgrad) to every parameter in ONE kernel launch
(``na_adam_multi``) -- the decoder has 16 tensors and 31,764 parameters, so torch's foreach implementation is a dozen
launch-latency-bound kernels per step.  Same constructor arguments and ``step`` / ``zero_grad`` / ``state_dict`` surface as
far as the trainer uses them; the moments live in two flat fp32 buffers.
"""
from __future__ import annotations

from typing import Iterable

import torch

from . import _lib, ops


class FusedAdam:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedAdam: no parameters")
        for p in self.params:
            ops._require_cuda(p)
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("FusedAdam: parameters must be contiguous fp32 CUDA tensors")
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.device = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.step_count = 0
        self._table = None
        self._table_key = None
        self.param_groups = [{"params": self.params, "lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay}]

    def _pointer_table(self) -> torch.Tensor:
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in self.params)
        if key != self._table_key:
            rows, off = [], 0
            for p in self.params:
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    raise RuntimeError("FusedAdam: gradients must be contiguous fp32 tensors")
                rows.append([p.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr() + 4 * off, self.exp_avg_sq.data_ptr() + 4 * off, p.numel()])
                off += p.numel()
            self._table = torch.tensor(rows, dtype=torch.int64).to(self.device)
            self._table_key = key
        return self._table

    @torch.no_grad()
    def step(self, grad_scale: torch.Tensor | None = None) -> None:
        if any(p.grad is None for p in self.params):
            raise RuntimeError("FusedAdam.step: every parameter needs a gradient")
        self.step_count += 1
        lr = float(self.param_groups[0]["lr"])
        b1, b2 = self.betas
        table = self._pointer_table()
        with torch.cuda.device(self.device):
            _lib.call("na_adam_multi", table.data_ptr(), len(self.params), max(p.numel() for p in self.params), lr, b1, b2, self.eps,
                      self.weight_decay, 1.0 - b1 ** self.step_count, 1.0 - b2 ** self.step_count, ops._ptr(grad_scale), ops._stream())
        # the update happens behind autograd's back (no version bump): drop the version-keyed packed-weight caches
        for owner in {id(o): o for o in getattr(self, "_owners", [])}.values():
            owner.invalidate_packed_weights()

    def attach(self, model) -> "FusedAdam":
        """Register a module whose packed-weight caches must be dropped after every step (``EEG_LSTM``)."""
        self._owners = getattr(self, "_owners", []) + [model]
        return self

    def zero_grad(self, set_to_none: bool = False) -> None:
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay}

    def load_state_dict(self, sd) -> None:
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
