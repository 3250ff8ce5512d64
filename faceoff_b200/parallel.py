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

The deferral is per forward: only a train-mode forward of the fused graph that records a backward routes its statistics
into the bucket (``begin_forward``).  Everything else -- ``torch.no_grad()``, the eager sub-method path
(``encode_quantized``), a stand-alone ``Quantize`` -- keeps the reference behaviour (all_reduce x2 + EMA inside forward).
A deferred forward whose backward never ran is flushed (statistics all-reduced, EMA applied) before the next one starts,
so no EMA update is ever lost or applied twice.

Micro-batches: inside ``with ddp.no_sync():`` backward neither communicates nor updates the codebooks;
gradients AND statistics keep accumulating in the bucket, and the first backward outside the context all-reduces the
accumulated bucket once and applies ONE EMA update for the whole step.
"""
from __future__ import annotations

import contextlib
import os

from typing import Dict, List, Optional, Sequence

import torch
from torch import distributed as dist
from torch import nn

from . import distributed as dist_fn


class FusedDataParallel(nn.Module):
    """Wraps the drop-in VQVAE like ``nn.parallel.DistributedDataParallel`` wraps the reference's
    (train_faceoff_perceptual.py:164-169): same forward signature, ``.module`` attribute, gradients in ``.grad``
    after ``loss.backward()``.  Do NOT wrap the frozen LPIPS (the reference's DDP(vqlpips) raises)."""

    def __init__(self, module: nn.Module, n_chunks: int = 4):
        super().__init__()
        self.module = module
        self.world = dist_fn.get_world_size()
        self.n_chunks = max(1, int(os.environ.get("FO_DP_CHUNKS", n_chunks)))   # env: experiments only
        self._pending_ema: List = []       # quantisers whose statistics sit in the bucket, EMA not applied yet
        self._accumulating = False         # statistics of earlier micro-batches (no_sync) are in the bucket
        self._sync = True                  # False inside no_sync()
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
        module._dp = self

    @contextlib.contextmanager
    def no_sync(self):
        """Like DistributedDataParallel.no_sync(): backward passes inside accumulate locally (gradients and the
        quantisers' EMA statistics); the first backward after the context does the one collective + one EMA update."""
        old, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = old

    # ---- bucket layout ------------------------------------------------------------------------
    def _build(self, backward_order: Sequence[str], params: Dict[str, torch.Tensor]):
        device = next(iter(params.values())).device
        n_stats = sum(q.n_embed + q.dim * q.n_embed for q in self._quantizers)
        n_grads = sum(params[n].numel() for n in backward_order)
        self._bucket = torch.zeros(n_stats + n_grads, dtype=torch.float32, device=device)
        self._grad_views, self._stat_views = {}, {}
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
        self._start_of = {self._offsets[n][0]: n for n in self._order}
        total = self._bucket.numel()
        self._chunk_bounds = [total * (i + 1) // self.n_chunks for i in range(self.n_chunks)]
        self._comm_stream = torch.cuda.Stream(device=device) if device.type == "cuda" else None

    # ---- forward-side hooks (called from vqvae.VQVAE._runner) -----------------------------------
    def begin_forward(self, params: Dict[str, torch.Tensor], forward_order: Sequence[str]):
        """Start of a deferred (train-mode, backward-recording) forward: lay the bucket out (once), settle a previous
        forward whose backward never ran, and zero the statistics unless micro-batches are being accumulated."""
        if self._bucket is None or self._bucket.device != next(iter(params.values())).device:
            self._build(list(reversed(list(forward_order))), params)
        if self._pending_ema and not self._accumulating:
            self._flush_pending()
        if not self._accumulating:
            self._bucket[:self._n_stats].zero_()

    def stat_buffers(self, q):
        """(counts, embed_sum) views of the bucket for quantiser ``q``; the gather kernel ADDS into them."""
        return self._stat_views[id(q)]

    def submit_stats(self, q):
        if all(e is not q for e in self._pending_ema):
            self._pending_ema.append(q)

    def _apply_pending(self):
        for q in self._pending_ema:
            counts, embed_sum = self._stat_views[id(q)]
            q.apply_ema(counts, embed_sum)
        self._pending_ema.clear()
        self._accumulating = False

    def _flush_pending(self):
        """A deferred forward was never followed by backward: do what the reference does inside forward (:63-75)."""
        if self.world > 1:
            dist.all_reduce(self._bucket[:self._n_stats], op=dist.ReduceOp.SUM)
        self._apply_pending()

    # ---- tape hooks (called from vqvae._GraphFn.backward) ---------------------------------------
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
            nxt = self._start_of.get(self._ready_upto)
            if nxt is None or nxt not in self._done:
                break
            self._ready_upto = self._offsets[nxt][1]
        if self._sync:
            self._launch_ready_chunks()

    def _launch_ready_chunks(self):
        if self.world == 1:
            return
        for b in self._chunk_bounds:
            if b <= self._launched_upto:
                continue
            if b <= self._ready_upto:
                self._launch(self._launched_upto, b)
                self._launched_upto = b

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
        if not self._sync:
            self._accumulating = True      # keep gradients and statistics in the bucket for the next micro-batch
            return
        if self.world > 1:
            if self._launched_upto < self._bucket.numel():
                self._launch(self._launched_upto, self._bucket.numel())
                self._launched_upto = self._bucket.numel()
            for w in self._works:
                w.wait()
            if self._comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self._comm_stream)
            self._bucket[self._n_stats:].mul_(1.0 / self.world)
        self._apply_pending()

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

    def bucket_checksums(self):
        """(statistics+gradient bucket, codebook buffers) as two fp64 sums -- identical on every rank after a
        synchronised step (used by bench.py's dp_check and the 2-rank tests)."""
        bsum = self._bucket.double().sum() if self._bucket is not None else torch.zeros((), dtype=torch.float64)
        csum = sum(b.double().sum() for q in self._quantizers for b in (q.embed, q.cluster_size, q.embed_avg))
        return bsum, csum

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def forward_with_ids(self, *args, **kwargs):
        return self.module.forward_with_ids(*args, **kwargs)
