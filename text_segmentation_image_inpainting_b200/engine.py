"""Training-step engine for the partial-conv U-Nets: flat fp32 parameter / gradient arenas, one fused SGD
launch, optional data-parallel gradient all-reduce (NCCL over NVLink) and whole-step CUDA-graph capture.

The reference has no train script: its recipe is prose -- SGD + Nesterov momentum, weight decay,
cyclic LR (checkpoints/ReadME.md:4).  One step here = forward + loss + backward (+ all-reduce) + SGD update,
the unit BASELINE.json's images/sec is quoted on.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch

from . import _lib, ops
from .masks import HoleMask

CL = torch.channels_last


def _flat_view(flat: torch.Tensor, off: int, like: torch.Tensor) -> torch.Tensor:
    """A view into `flat` with the same logical shape AND physical element order as `like`."""
    n = like.numel()
    seg = flat[off:off + n]
    if like.dim() == 4 and like.is_contiguous(memory_format=CL) and not like.is_contiguous():
        co, ci, kh, kw = like.shape
        return seg.view(co, kh, kw, ci).permute(0, 3, 1, 2)
    return seg.view(like.shape)


class FlatParams:
    """All trainable parameters of a module re-pointed into one fp32 arena (and their grads into another)."""

    def __init__(self, module: torch.nn.Module):
        self.params: List[torch.nn.Parameter] = [p for p in module.parameters() if p.requires_grad]
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        # pad every tensor to a multiple of 4 elements so all views stay 16-byte aligned
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.numel = off
        self.flat_p = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(off, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                v = _flat_view(self.flat_p, o, p.data)
                v.copy_(p.data)
                p.data = v
                p.grad = _flat_view(self.flat_g, o, p.data)
        self.true_numel = total
        # Gradient sinks (ops.GradSink): the weight-gradient kernels (4-D convolution weights) and the BatchNorm backward
        # (1-D scale / shift) write straight into the arena -- no gradient tensor, no autograd accumulation kernel.
        self.sinks = []
        self.sink_of = {}
        for i, p in enumerate(self.params):
            ok4 = p.dim() == 4 and p.grad.is_contiguous(memory_format=CL) and (p.grad.data_ptr() % 16) == 0
            ok1 = p.dim() == 1
            if ok4 or ok1:
                p._pcb_grad_sink = ops.GradSink(p.grad)
                p._pcb_grad_sink.prezeroed = True        # TrainStep zeroes the whole gradient arena at the start of every step
                self.sinks.append(p._pcb_grad_sink)
                self.sink_of[i] = p._pcb_grad_sink


class TrainStep:
    def __init__(self, net: torch.nn.Module, compute_dtype=torch.bfloat16, lr=2e-4, momentum=0.9, weight_decay=1e-4,
                 nesterov=True, process_group=None, use_graph=True, bucket_mb=32, overlap_allreduce=None):
        self.net = net.train()
        self.dtype = compute_dtype
        self.lr, self.momentum, self.wd, self.nesterov = lr, momentum, weight_decay, nesterov
        self.flat = FlatParams(net)
        self._wcaches = [m._wcache for m in net.modules() if isinstance(getattr(m, "_wcache", None), dict)]
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        # data parallel: ranks exchange the SUM of their gradient arenas; the 1/world of the mean is folded into the optimiser
        # kernel (no scaling pass over the 131 MB arena)
        self.grad_scale = 1.0 / self.world
        self.use_graph = use_graph
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.static_x = self.static_m = self.static_loss = None
        self.first = True
        self.bucket_elems = bucket_mb * (1 << 20) // 4
        self.launches_per_step = 0
        self.graph_update = None
        self._captured_operands = None
        is_cuda = self.flat.flat_p.is_cuda
        if overlap_allreduce is None:
            overlap_allreduce = os.environ.get("PCB_DDP_OVERLAP", "1") != "0"
        self.overlap = bool(overlap_allreduce) and self.world > 1 and is_cuda
        self.overlap_active = False
        self._comm_stream = torch.cuda.Stream(device=self.flat.flat_p.device) if (self.world > 1 and is_cuda) else None
        self._hooks_installed = False
        self._make_buckets()
        if self.pg is not None:
            self._sync_replicas()

    def _sync_replicas(self):
        """Data parallel start-up: every rank adopts rank 0's parameters and module buffers (BatchNorm running statistics,
        `num_batches_tracked`), like torch DDP does, so replicas built from different seeds or checkpoints cannot silently
        diverge.  BatchNorm batch statistics stay rank-local afterwards (the reference has no SyncBN); running statistics
        therefore drift per rank during training and rank 0's are the ones to checkpoint."""
        dist = torch.distributed
        src = dist.get_global_rank(self.pg, 0)
        dist.broadcast(self.flat.flat_p, src=src, group=self.pg)
        for b in self.net.buffers():
            dist.broadcast(b, src=src, group=self.pg)
        ops.bump_weight_epoch()

    # -- gradient exchange ---------------------------------------------------------------------------
    def _make_buckets(self):
        """Buckets of ~bucket_elems consecutive arena elements, cut at parameter boundaries.  Backward produces gradients
        roughly from the END of the arena (decoder) to its start (encoder), so buckets complete tail first."""
        fp = self.flat
        self.buckets = []               # [start, end, [param indices]]
        start, members = 0, []
        for i, off in enumerate(fp.offsets):
            end = fp.offsets[i + 1] if i + 1 < len(fp.offsets) else fp.numel
            members.append(i)
            if end - start >= self.bucket_elems or i + 1 == len(fp.offsets):
                self.buckets.append([start, end, members])
                start, members = end, []
        self._bucket_of = {}
        for b, (_, _, mem) in enumerate(self.buckets):
            for i in mem:
                self._bucket_of[i] = b

    def _install_hooks(self):
        """A parameter reports ONCE per step (its sink's first write, or its autograd accumulation).  A weight used twice in one
        forward would therefore be exchanged before its second contribution lands: the reference's networks share no weights;
        construct with overlap_allreduce=False for models that do."""
        if self._hooks_installed:
            return
        self._hooks_installed = True
        for i, p in enumerate(self.flat.params):
            sink = self.flat.sink_of.get(i)
            if sink is not None:
                sink.on_written = (lambda idx: (lambda: self._param_ready(idx)))(i)
            # parameters whose gradient still arrives through autograd (or a sink that was refused and fell back to it)
            p.register_post_accumulate_grad_hook((lambda idx: (lambda _p: self._param_ready(idx)))(i))

    def _param_ready(self, idx):
        if not self._overlap_armed:
            return
        b = self._bucket_of[idx]
        if idx in self._pending[b]:
            self._pending[b].discard(idx)
            if not self._pending[b]:
                self._launch_bucket(b)

    def _launch_bucket(self, b):
        """All-reduce (sum) of bucket b on the communication stream, ordered after everything issued so far on the compute
        stream (BatchNorm / bias gradients, autograd accumulations) and on the weight-gradient side stream."""
        if self._launched[b]:
            return
        self._launched[b] = True
        s, e, _ = self.buckets[b]
        comm = self._comm_stream
        comm.wait_stream(torch.cuda.current_stream())
        for st in ops.side_streams():
            comm.wait_stream(st)
        with torch.cuda.stream(comm):
            torch.distributed.all_reduce(self.flat.flat_g[s:e], group=self.pg)

    def _arm_overlap(self, on: bool):
        self._overlap_armed = bool(on)
        if on:
            self._install_hooks()
            self._pending = [set(mem) for _, _, mem in self.buckets]
            self._launched = [False] * len(self.buckets)

    _overlap_armed = False

    def _finish_overlap(self):
        """Buckets whose parameters did not all report (an unused parameter) are exchanged now; then the compute stream waits
        for the communication stream."""
        for b in range(len(self.buckets) - 1, -1, -1):
            if not self._launched[b]:
                self._launch_bucket(b)
        self._overlap_armed = False
        torch.cuda.current_stream().wait_stream(self._comm_stream)

    def _allreduce(self):
        """Un-overlapped exchange (fallback, and the CPU/gloo path): bucketed SUM all-reduce of the whole arena."""
        if self.world == 1:
            return
        g = self.flat.flat_g
        for s in range(0, g.numel(), self.bucket_elems):
            torch.distributed.all_reduce(g[s:s + self.bucket_elems], group=self.pg)

    # -- one eager step ----------------------------------------------------------------------------
    def _prepare(self, x: torch.Tensor, mask: torch.Tensor):
        """reference-style inputs (fp32 NCHW image, fp32/uint8 {0,1} NCHW mask) -> (x*mask in compute dtype NHWC, HoleMask)"""
        n, c, h, w = x.shape
        # the 3-channel image travels as an 8-channel-padded NHWC buffer: 16-byte pixels feed the row-packed
        # tensor-core stem and the tail's second gather source directly
        buf = torch.empty((n, (c + 7) // 8 * 8, h, w), dtype=self.dtype, device=x.device, memory_format=CL).zero_()
        xin = buf[:, :c]
        xin.copy_(x * mask.to(x.dtype))                                                # Dataloader.py:131
        # masks of the reference's data path are one plane repeated over RGB (Dataloader.py:128-129); outside a graph capture
        # that promise is checked (one device sync), inside a capture it was checked by the eager warm-up steps
        if not torch.cuda.is_current_stream_capturing() and mask.shape[1] > 1:
            if not bool((mask == mask[:, :1]).all()):
                raise ValueError("TrainStep expects a channel-uniform hole mask (one plane repeated over the input channels)")
        return xin, HoleMask.from_dense(mask, channel_uniform=True)

    def _fwd_bwd(self, x, mask, overlap=False):
        self.flat.flat_g.zero_()
        for sk in self.flat.sinks:
            sk.used = False
        ops.begin_step_arena(self.flat.flat_g.device)      # one zero fill for every reduction target of the step
        ops.bump_weight_epoch()
        # scheduling switches that are only safe inside a loop that joins once per step (scoped to this call):
        # operand buffers refreshed in place, mask passes running ahead on their own stream
        ops.set_inplace_weight_refresh(True)
        ops.set_mask_chain_stream(True)
        self._arm_overlap(overlap)
        try:
            ops.prefetch_weights(self._wcaches)        # operand re-layout of all layers runs ahead on its own stream
            loss = self._forward_loss(x, mask)
            loss.backward()
        finally:
            ops.set_mask_chain_stream(False)
            ops.set_inplace_weight_refresh(False)
            ops.end_step_arena(self.flat.flat_g.device)
            ops.join_side_streams()
            if overlap:
                self._finish_overlap()
        return loss.detach()

    def _forward_loss(self, x, mask):
        """Forward + scalar loss of one step (SURVEY 8d benchmark loss: out.abs().mean())."""
        xin, hm = self._prepare(x, mask)
        return ops.l1_mean(self.net((xin, hm)))

    def _update(self, first_step: bool):
        ops.sgd_step(self.flat.flat_p, self.flat.flat_g, self.flat.flat_m, self.lr, self.momentum, self.wd, self.nesterov,
                     first_step, grad_scale=self.grad_scale)

    def _step(self, x, mask, first_step: bool, overlap=None):
        overlap = self.overlap if overlap is None else overlap
        loss = self._fwd_bwd(x, mask, overlap=overlap)
        if self.world > 1 and not overlap:
            self._allreduce()
        self._update(first_step)
        return loss

    # -- public -------------------------------------------------------------------------------------
    def warmup_and_capture(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None, eager_warmup=2):
        """Run eager warm-up steps (also counts this library's launches per step), then capture the step."""
        for _ in range(eager_warmup):
            before = _lib.launch_count()
            self._step(x, mask, self.first)
            self.first = False
            self.launches_per_step = _lib.launch_count() - before
        torch.cuda.synchronize()
        self.overlap_active = self.overlap
        if not self.use_graph:
            return
        self.static_x = x.clone()
        self.static_m = mask.clone() if mask is not None else None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._step(self.static_x, self.static_m, False)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = None
        if self.world == 1 or self.overlap:
            # one graph for the whole step.  Data parallel: the bucketed NCCL all-reduces are captured with it, on the
            # communication stream, each ordered after the kernels that complete its bucket -- they overlap the rest of backward
            try:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self.static_loss = self._step(self.static_x, self.static_m, False)
            except Exception as exc:  # noqa: BLE001 -- a collective that cannot be captured: fall back to the split scheme
                if self.world == 1:
                    raise
                print(f"[engine] capturing the overlapped all-reduce failed ({type(exc).__name__}: {exc}); "
                      "falling back to graph(forward+backward) | eager all-reduce | graph(SGD)", flush=True)
                graph = None
                self.overlap = self.overlap_active = False
                self._overlap_armed = False
                torch.cuda.synchronize()
        if graph is None:
            # data parallel fallback: the NCCL all-reduce stays OUTSIDE the captured graphs (graph A = forward+backward,
            # eager bucketed all-reduce of the gradient arena, graph B = fused SGD): no collective is captured
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.static_loss = self._fwd_bwd(self.static_x, self.static_m)
            self.graph_update = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_update):
                self._update(False)
        self.graph = graph
        # The captured kernels hold raw pointers to the per-layer operand buffers (the `_wcache` entries): keep them alive for as
        # long as the graph exists.  An eager forward after capture (validation) misses the cache -- the optimiser bumps the
        # weight epoch -- and, with the in-place refresh off, replaces `cache["val"]` with new buffers; without this list the
        # old ones would be freed under the graph.  (Each replay refreshes the captured buffers itself.)
        self._captured_operands = [c.get("val") for c in self._wcaches]
        torch.cuda.synchronize()

    def step(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x, mask already on the device.  Returns the (device) loss of this step."""
        if self.graph is not None:
            if x.data_ptr() != self.static_x.data_ptr():
                self.static_x.copy_(x, non_blocking=True)
            if mask is not None and mask.data_ptr() != self.static_m.data_ptr():
                self.static_m.copy_(mask, non_blocking=True)
            self.graph.replay()
            if self.graph_update is not None:
                self._allreduce()
                self.graph_update.replay()
            return self.static_loss
        loss = self._step(x, mask, self.first)
        self.first = False
        return loss


    def close(self):
        """Destroy the captured graphs (the step falls back to eager mode).  REQUIRED before
        `torch.distributed.destroy_process_group()` when the all-reduce was captured (`overlap_active`): NCCL keeps a reference
        per graph that holds captured collectives and its communicator teardown waits until those graphs are gone -- with the
        graphs alive the process hangs at exit (measured: the 2-GPU bench printed its line and never returned)."""
        if self.graph is not None or self.graph_update is not None:
            torch.cuda.synchronize()
            self.graph = None
            self.graph_update = None
            self._captured_operands = None
            import gc
            gc.collect()
            torch.cuda.synchronize()


class SegTrainStep(TrainStep):
    """The same step for the dense segmentation networks (models/text_segmentation.py: `net(x)`, no masks): BASELINE.json
    configs[1] (TextSegament, batch 8) and configs[3] (XceptionTextSegment, batch 16, bf16).  `mask` is ignored."""

    def _forward_loss(self, x, mask=None):
        n, c, h, w = x.shape
        buf = torch.empty((n, (c + 7) // 8 * 8, h, w), dtype=self.dtype, device=x.device, memory_format=CL).zero_()
        xin = buf[:, :c]
        xin.copy_(x)
        return ops.l1_mean(self.net(xin))
