"""Data-parallel training with ONE flat fp32 gradient buffer (SURVEY.md sections 2c, 8e / 8f row 4).

`FlatDataParallel(model)` stands where the reference puts `DistributedDataParallel(model, device_ids=[rank],
find_unused_parameters=True)` (train_dist.py:147): one process per GPU, replicas kept identical by averaging gradients
over NCCL.  What differs is the plumbing.  Every parameter is a view of one flat buffer and every `.grad` a view of one
flat gradient buffer (`FlatParams`, shared with `optim.FlatAdam`), so the step's collective is an all-reduce of that
buffer - 9,695,954 floats = 38.8 MB for LineRefineNet - instead of DDP's bucket copies and per-bucket reductions: ONE
NCCL all-reduce when the backward pass ends (0.1 ms on 2 B200s, 0.19 ms on 8, against a 68 ms step).  `overlap=True`
reduces it in two slices instead, the parameters autograd finishes FIRST (everything behind the context encoder: decoder,
heads, projections, 27.6 MB) as soon as their last gradient is there, i.e. under the encoder's backward, and the encoder's
own slice at the end.  Measured, that LOSES here (72.8 vs 68.1 ms per step on 2 GPUs): the train kernels are persistent
and sized for all 148 SMs, so the NCCL kernel and the GEMMs wait for each other's SMs, and there is next to nothing to
hide.  Buffers (BatchNorm running statistics) are broadcast from rank 0 before each forward, like DDP's
broadcast_buffers=True.

No gradient leaves the device and nothing here is on the inference path; `torch.distributed` is the transport.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn

from .flat import FlatParams


class FlatDataParallel(nn.Module):
    def __init__(self, module: nn.Module, process_group=None, overlap: bool = False, broadcast_buffers: bool = True):
        super().__init__()
        if not dist.is_initialized():
            raise RuntimeError("FlatDataParallel needs an initialised torch.distributed process group")
        self.module = module
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.overlap = overlap
        self.broadcast_buffers = broadcast_buffers
        params = [p for p in module.parameters() if p.requires_grad]
        self.flat = FlatParams.of(params) or FlatParams(params)
        dist.broadcast(self.flat.flat, src=self._src(), group=process_group)       # replicas start identical (DDP does the same)
        self._buffers = [b for b in module.buffers() if b.is_floating_point()]
        self._int_buffers = [b for b in module.buffers() if not b.is_floating_point()]
        self._sync_buffers()
        # Bucket 0 = the slice whose gradients autograd completes first.  Parameters are laid out in
        # module.parameters() order (LineRefineNet: context_encoder first) and the backward pass reaches the
        # encoder last, so the early slice is everything from the first non-encoder parameter on.
        first = getattr(module, "context_encoder", None)
        n_first = len([p for p in first.parameters() if p.requires_grad]) if isinstance(first, nn.Module) else 0
        self._split = self.flat.offsets[n_first] if 0 < n_first < len(params) else 0
        self._late = set(range(n_first)) if self._split else set()
        self._pending_early = 0
        self._work = []
        self._armed = False
        self.allreduce_calls = 0
        for i, p in enumerate(params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))

    def _src(self):
        return dist.get_global_rank(self.group, 0) if self.group is not None else 0

    def _sync_buffers(self):
        if not self.broadcast_buffers or not (self._buffers or self._int_buffers):
            return
        for bufs in (self._buffers, self._int_buffers):
            if bufs:
                flat = torch.cat([b.detach().reshape(-1) for b in bufs])
                dist.broadcast(flat, src=self._src(), group=self.group)
                o = 0
                with torch.no_grad():
                    for b in bufs:
                        b.copy_(flat[o:o + b.numel()].view(b.shape))
                        o += b.numel()

    # -- gradient reduction -------------------------------------------------------------------------------
    def _all_reduce(self, t):
        self.allreduce_calls += 1
        if dist.get_backend(self.group) == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True)   # gloo has no AVG
        self._scale.append(t)
        return w

    def _make_hook(self, index):
        def hook(param):
            g = self.flat.grad_views[index]
            if param.grad is not g:               # autograd bound a fresh tensor (first backward after set_to_none)
                g.copy_(param.grad)
                param.grad = g
            if not self._armed:                   # first gradient of this backward pass
                self._armed = True
                self._scale = []
                self._pending_early = len(self.flat.params) - len(self._late)
                torch.autograd.Variable._execution_engine.queue_callback(self._finish)
            if self.overlap and self._split and index not in self._late:
                self._pending_early -= 1
                if self._pending_early == 0:      # decoder-side slice complete: reduce it under the encoder's backward
                    self._work.append(self._all_reduce(self.flat.grad[self._split:]))
        return hook

    def _finish(self):
        early_done = bool(self._work)
        self._work.append(self._all_reduce(self.flat.grad[:self._split] if early_done else self.flat.grad))
        for w in self._work:
            w.wait()                              # NCCL: the current stream waits for the collective; no host sync
        for t in self._scale:
            t.div_(self.world)
        self._work, self._armed = [], False

    def forward(self, *args, **kwargs):
        if self.module.training:
            self._sync_buffers()
        return self.module(*args, **kwargs)
