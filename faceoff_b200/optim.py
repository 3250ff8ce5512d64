"""Fused optimizer step.  The reference trains with ``optim.Adam(model.parameters(), lr=args.lr)``
(train_faceoff_perceptual.py:190, lr 3e-4, torch defaults otherwise); ``FusedAdam`` has the same constructor and
``step()`` semantics (no amsgrad / maximize / capturable) and the same per-parameter state names (``step``, ``exp_avg``,
``exp_avg_sq``), so ``optimizer.state_dict()`` round-trips with ``torch.optim.Adam``.  All parameters of all groups that
share hyper-parameters are updated by ONE kernel launch (``fo_adam_step``).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from . import ops
from ._lib import FaceoffB200Error


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        if lr < 0.0 or eps < 0.0 or weight_decay < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0):
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables: Dict[tuple, tuple] = {}   # pointer signature -> (table, chunks, n_chunks) on the device

    def _table_for(self, plist: List[torch.Tensor], states: List[dict]):
        rows = [(p.data_ptr(), p.grad.data_ptr(), s["exp_avg"].data_ptr(), s["exp_avg_sq"].data_ptr(), p.numel())
                for p, s in zip(plist, states)]
        key = tuple(rows)
        hit = self._tables.get(key)
        if hit is not None:
            return hit
        ce = ops.adam_chunk_elems()
        chunks = [(i, c) for i, r in enumerate(rows) for c in range((r[4] + ce - 1) // ce)]
        dev = plist[0].device
        # pinned staging + async copies: no host synchronisation when gradients move to new addresses
        h_table = torch.tensor(rows, dtype=torch.int64).pin_memory()
        h_cmap = torch.tensor(chunks, dtype=torch.int32).pin_memory()
        table = h_table.to(dev, non_blocking=True)
        cmap = h_cmap.to(dev, non_blocking=True)
        if len(self._tables) > 8:   # gradients re-allocated at new addresses every step: keep the cache small
            self._tables.clear()
        self._tables[key] = (table, cmap, len(chunks), h_table, h_cmap)
        return self._tables[key]

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            plist, states = [], []
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.dtype == torch.float32
                        and p.grad.is_contiguous() and not p.grad.is_sparse):
                    raise FaceoffB200Error("FusedAdam: contiguous fp32 CUDA parameters and gradients required")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                plist.append(p)
                states.append(st)
            if not plist:
                continue
            steps = {int(s["step"].item()) if torch.is_tensor(s["step"]) else int(s["step"]) for s in states}
            if len(steps) != 1:
                raise FaceoffB200Error("FusedAdam: parameters of one group must share their step count")
            step = steps.pop() + 1
            table, cmap, n_chunks = self._table_for(plist, states)[:3]
            b1, b2 = group["betas"]
            ops.adam_step(table, cmap, n_chunks, group["lr"], b1, b2, group["eps"], group["weight_decay"], step,
                          grad_scale)
            # the kernel writes through raw pointers: tell autograd (and the packed-weight cache in ops.packed_weights,
            # which is keyed on the tensor version) that every parameter changed in place
            torch._C._increment_version(plist)
            for s in states:
                s["step"] = torch.tensor(float(step))
        return loss
