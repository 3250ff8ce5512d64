"""Gate-consistent fp64 oracle (test infrastructure).

ReLU and max-pool make the training step piecewise linear; which piece a run is on is decided by the signs of the
pre-activations and by the pooling winners.  An element whose pre-activation lies within the forward rounding error of
zero may legitimately fall on the other side in another (equally correct) implementation -- and then a whole row of a
weight gradient differs by O(1/sqrt(N)), although every kernel is right.  (The reference's own fp32 run shows the same
effect against fp64: see the |ref_fp32 - fp64| column printed by the tests.)

``GateRecorder`` records, during a run of the CUDA path, every ReLU output (``relu`` outputs of ops.conv, ops.relu) and
every max-pool input.  Installed as the oracle's RELU_HOOK / MAXPOOL_HOOK it makes the fp64 oracle take the SAME branch:
each oracle ReLU / pool is matched to the recorded tensor BY VALUE (same shape, max-normalised difference < 1e-3) and uses
its gates / winners.  Every override is audited: it is only accepted where the oracle's own pre-activation (or top-2 gap of
the window) is below ``tol`` x the tensor's max, i.e. where the branch really is undecidable at the forward accuracy;
anything else raises.  With the branch fixed the step is linear in the seed gradient and every gradient must agree tightly.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from faceoff_b200 import ops
from oracle import faceoff_oracle as O


def _merged_nc(t: torch.Tensor) -> torch.Tensor:
    """channels-last activation (hi|lo pairs in the verification mode) -> fp64 N C (D) H W on the CPU."""
    if ops.PRECISE:
        h = t.shape[-1] // 2
        v = t[..., :h].double() + t[..., h:].double()
    else:
        v = t.double()
    perm = (0, v.dim() - 1) + tuple(range(1, v.dim() - 1))
    return v.permute(*perm).contiguous().cpu()


class GateRecorder:
    def __init__(self, tol: float = 1e-4):
        self.tol = tol
        self.acts, self.pool_in = [], []
        self.relu_overrides = self.pool_overrides = 0
        self.relu_elems = self.pool_windows = 0

    # ---- recording (wraps the product ops) ----
    def __enter__(self):
        self._conv, self._relu, self._pool = ops.conv, ops.relu, ops.maxpool2
        rec = self

        def conv(*a, **k):
            out = rec._conv(*a, **k)
            if out[1] is not None:
                rec.acts.append(_merged_nc(out[1]))
            return out

        def relu(x):
            y = rec._relu(x)
            rec.acts.append(_merged_nc(y))
            return y

        def pool(x):
            rec.pool_in.append(_merged_nc(x))
            return rec._pool(x)

        ops.conv, ops.relu, ops.maxpool2 = conv, relu, pool
        return self

    def __exit__(self, *exc):
        ops.conv, ops.relu, ops.maxpool2 = self._conv, self._relu, self._pool

    # ---- replay (hooks of the oracle) ----
    def _match(self, pool, ref):
        best, best_e = None, 1e9
        for a in pool:
            if a.dim() == 5 and ref.dim() == 5:
                cand = a.permute(0, 1, 2, 3, 4)      # both N C D H W
            else:
                cand = a
            if cand.shape != ref.shape:
                continue
            e = ((cand - ref).abs().max() / (ref.abs().max() + 1e-300)).item()
            if e < best_e:
                best, best_e = cand, e
        return best, best_e

    def relu_hook(self, x):
        xd = x.detach().double()
        cand, e = self._match(self.acts, xd.clamp_min(0))
        if cand is None or e > 1e-3:
            return None          # a ReLU the CUDA path does not materialise (or another dtype run): plain relu
        gate = cand > 0
        flips = gate != (xd > 0)
        self.relu_elems += xd.numel()
        if flips.any():
            worst = (xd.abs()[flips].max() / xd.abs().max()).item()
            assert worst <= self.tol, f"ReLU gate differs where |pre-activation| = {worst:.2e} x max (not a near-zero element)"
            self.relu_overrides += int(flips.sum())
        return x * gate.to(x.dtype)

    def maxpool_hook(self, x):
        xd = x.detach().double()
        cand, e = self._match(self.pool_in, xd)
        if cand is None or e > 1e-3:
            return None
        _, idx = F.max_pool2d(cand, 2, 2, return_indices=True)
        y_ref, idx_ref = F.max_pool2d(xd, 2, 2, return_indices=True)
        flips = idx != idx_ref
        self.pool_windows += idx.numel()
        n, c = xd.shape[:2]
        y = x.flatten(2).gather(2, idx.flatten(2)).view(n, c, *idx.shape[2:])
        if flips.any():
            gap = (y_ref - y.detach().double())[flips]          # winner minus the element the CUDA path picked
            worst = (gap.abs().max() / xd.abs().max()).item()
            assert worst <= self.tol, f"max-pool winner differs where the gap is {worst:.2e} x max"
            self.pool_overrides += int(flips.sum())
        return y

    def install(self):
        O.RELU_HOOK, O.MAXPOOL_HOOK = self.relu_hook, self.maxpool_hook

    @staticmethod
    def uninstall():
        O.RELU_HOOK = O.MAXPOOL_HOOK = None

    def summary(self):
        return (f"{self.relu_overrides} of {self.relu_elems} ReLU gates and {self.pool_overrides} of {self.pool_windows} "
                f"pooling winners taken from the CUDA run (all audited: undecidable within {self.tol:g} x max)")
