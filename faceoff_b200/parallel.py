"""Fused data-parallel layer: ONE bucketed all-reduce per step.

The reference's data parallelism issues, per step: 4 blocking all-reduces inside forward for the two
quantisers' EMA statistics (models/vqvae_conv3d_latent.py:63-64), DDP's bucketed gradient all-reduce, and 6 buffer
broadcasts.  Here every parameter gradient is *born* inside one flat fp32 bucket (the weight-gradient kernels write
straight into views of it) next to the quantisers' cluster-size counts and embedding sums:

    flat = [ counts_t | embed_sum_t | counts_b | embed_sum_b | grads in backward order ... ]

Backward produces the bucket front to back, so it is all-reduced (SUM, NCCL over NVLink/NVSwitch) in a few chunks on a
side stream while the remaining backward still runs.  Afterwards gradients are scaled by 1/world (DDP average) and the
EMA update (:66-75) is applied with the summed statistics -- legal because ``quantize`` uses the pre-update codebook
(:57 precedes :75) and nothing reads ``embed`` again within the step.  Codebooks stay bit-identical across ranks, so
DDP's buffer broadcast is dropped.
"""
from __future__ import annotations

import os

from typing import Dict, List, Optional, Sequence

import torch
from torch import distributed as dist
from torch import nn

from . import distributed as dist_fn


class FlatBucket:
    """Flat fp32 buffer holding gradient tensors (averaged over ranks) and statistic tensors (summed)."""

    def __init__(self, grads: Sequence[torch.Tensor], stats: Sequence[torch.Tensor], device=None):
        self.ext_grads, self.ext_stats = list(grads), list(stats)
        device = device if device is not None else self.ext_grads[0].device
        n_stats = sum(t.numel() for t in self.ext_stats)
        n_grads = sum(t.numel() for t in self.ext_grads)
        self.flat = torch.zeros(n_stats + n_grads, dtype=torch.float32, device=device)
        self.n_stats = n_stats
        self.stat_views, self.grad_views = [], []
        off = 0
        for t in self.ext_stats:
            self.stat_views.append(self.flat[off:off + t.numel()].view(t.shape))
            off += t.numel()
        for t in self.ext_grads:
            self.grad_views.append(self.flat[off:off + t.numel()].view(t.shape))
            off += t.numel()

    def pack(self):
        for v, t in zip(self.stat_views + self.grad_views, self.ext_stats + self.ext_grads):
            v.copy_(t)

    def all_reduce(self, lo: int = 0, hi: Optional[int] = None, async_op: bool = False):
        hi = self.flat.numel() if hi is None else hi
        if dist_fn.get_world_size() == 1 or hi <= lo:
            return None
        return dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=async_op)

    def scale_grads(self):
        ws = dist_fn.get_world_size()
        if ws > 1:
            self.flat[self.n_stats:].mul_(1.0 / ws)

    def unpack(self):
        self.scale_grads()
        for v, t in zip(self.stat_views + self.grad_views, self.ext_stats + self.ext_grads):
            t.copy_(v)


class _DeferredStatSink:
    """Quantize.stat_sink replacement: statistics live in the bucket, EMA is applied after the collective."""

    def __init__(self, owner: "FusedDataParallel"):
        self.owner = owner

    def buffers(self, q):
        return self.owner._stat_buffers(q)

    def submit(self, q, counts, embed_sum):
        self.owner._pending_ema.append((q, counts, embed_sum))


class FusedDataParallel(nn.Module):
    """Wraps the drop-in VQVAE like ``nn.parallel.DistributedDataParallel`` wraps the reference's
    (train_faceoff_perceptual.py:164-169): same forward signature, ``.module`` attribute, gradients in ``.grad``
    after ``loss.backward()``.  Do NOT wrap the frozen LPIPS (the reference's DDP(vqlpips) raises)."""

    def __init__(self, module: nn.Module, n_chunks: int = 4):
        super().__init__()
        self.module = module
        self.world = dist_fn.get_world_size()
        self.n_chunks = max(1, int(os.environ.get("FO_DP_CHUNKS", n_chunks)))   # env: experiments only
        self._pending_ema: List = []
        self._bucket: Optional[torch.Tensor] = None
        self._grad_views: Dict[str, torch.Tensor] = {}
        self._stat_views: Dict[int, tuple] = {}
        self._order: List[str] = []
        self._comm_stream = None
        self._works: List = []
        self._ready_upto = 0
        self._launched_upto = 0
        from .vqvae import Quantize

        self._quantizers = [m for m in module.modules() if isinstance(m, Quantize)]
        if self.world > 1:
            sink = _DeferredStatSink(self)
            for q in self._quantizers:
                q.stat_sink = sink
        module._dp = self

    # ---- bucket layout ------------------------------------------------------------------------
    def _build(self, backward_order: Sequence[str], params: Dict[str, torch.Tensor]):
        device = next(iter(params.values())).device
        n_stats = sum(q.n_embed + q.dim * q.n_embed for q in self._quantizers)
        n_grads = sum(params[n].numel() for n in backward_order)
        self._bucket = torch.zeros(n_stats + n_grads, dtype=torch.float32, device=device)
        off = 0
        for q in self._quantizers:
            c = self._bucket[off:off + q.n_embed]
            off += q.n_embed
            s = self._bucket[off:off + q.dim * q.n_embed].view(q.dim, q.n_embed)
            off += q.dim * q.n_embed
            self._stat_views[id(q)] = (c, s)
        self._n_stats = off
        self._offsets = {}
        for n in backward_order:
            k = params[n].numel()
            self._grad_views[n] = self._bucket[off:off + k].view(params[n].shape)
            self._offsets[n] = (off, off + k)
            off += k
        self._order = list(backward_order)
        total = self._bucket.numel()
        self._chunk_bounds = [total * (i + 1) // self.n_chunks for i in range(self.n_chunks)]
        if device.type == "cuda":
            self._comm_stream = torch.cuda.Stream(device=device)

    def _stat_buffers(self, q):
        c, s = self._stat_views[id(q)]
        c.zero_()
        s.zero_()
        return c, s

    # ---- tape hooks (called from vqvae._GraphFn) ---------------------------------------------
    def begin_forward(self, params: Dict[str, torch.Tensor], forward_order: Sequence[str]):
        """Called at the start of the module's forward: lay the bucket out (once) so the quantisers can write
        their statistics into it."""
        if self._bucket is None or self._bucket.device != next(iter(params.values())).device:
            self._build(list(reversed(list(forward_order))), params)

    def begin_step(self, tape):
        """Called when backward starts: hand the gradient views to the tape."""
        self._done = set()
        self._ready_upto = self._n_stats  # statistics are complete once forward is done
        self._launched_upto = 0
        self._works = []
        tape.dp = self
        for n, v in self._grad_views.items():
            p = tape.params[n]
            tape.grads[n] = v
            if p.grad is not None and p.grad.data_ptr() == v.data_ptr():
                continue  # user is accumulating over several backward passes: keep adding
            tape._fresh.add(n)

    def grad_ready(self, name: str):
        self._done.add(name)
        # advance the contiguous ready frontier
        while True:
            nxt = next((n for n in self._order if self._offsets[n][0] == self._ready_upto), None)
            if nxt is None or nxt not in self._done:
                break
            self._ready_upto = self._offsets[nxt][1]
        self._launch_ready_chunks()

    def _launch_ready_chunks(self, force: bool = False):
        if self.world == 1:
            return
        for b in self._chunk_bounds:
            if b <= self._launched_upto:
                continue
            if b <= self._ready_upto or force:
                hi = b if not force else self._bucket.numel()
                self._launch(self._launched_upto, hi)
                self._launched_upto = hi
                if force:
                    break

    def _launch(self, lo: int, hi: int):
        seg = self._bucket[lo:hi]
        if self._comm_stream is not None:
            self._comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm_stream):
                self._works.append(dist.all_reduce(seg, op=dist.ReduceOp.SUM, async_op=True))
        else:
            self._works.append(dist.all_reduce(seg, op=dist.ReduceOp.SUM, async_op=True))

    def end_step(self):
        """Called when backward finished: flush, wait (stream-ordered), average grads, apply EMA."""
        if self.world > 1:
            if self._launched_upto < self._bucket.numel():
                self._launch(self._launched_upto, self._bucket.numel())
                self._launched_upto = self._bucket.numel()
            for w in self._works:
                w.wait()
            if self._comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self._comm_stream)
            self._bucket[self._n_stats:].mul_(1.0 / self.world)
        for q, counts, embed_sum in self._pending_ema:
            q.apply_ema(counts, embed_sum)
        self._pending_ema.clear()

    def grads_for_autograd(self, tape, names: Sequence[str]):
        """Gradients live in the bucket: bind them to ``.grad`` directly (no copy) and return None to autograd."""
        out = []
        for n in names:
            p, v = tape.params[n], self._grad_views.get(n)
            if v is None or not p.requires_grad:
                out.append(tape.grads.get(n))
                continue
            if p.grad is None:
                p.grad = v
                out.append(None)
            elif p.grad.data_ptr() == v.data_ptr():
                out.append(None)
            else:
                out.append(v)  # foreign .grad tensor: let autograd accumulate into it
        return tuple(out)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def forward_with_ids(self, *args, **kwargs):
        return self.module.forward_with_ids(*args, **kwargs)
