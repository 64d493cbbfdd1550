"""Flat parameter / gradient storage shared by optim.FlatAdam and ddp.FlatDataParallel (one flat
gradient buffer -> one all-reduce, one Adam launch)."""
from __future__ import annotations

import torch


class FlatParams:
    """Flat fp32 storage for a parameter list: `flat` holds the values (each `p.data` becomes a view), `grad` the
    gradients (each `p.grad` is a view).  Parameters start 64-element (256-byte) aligned."""

    def __init__(self, params):
        params = [p for p in params]
        if not params:
            raise ValueError("FlatParams got an empty parameter list")
        dev, dt = params[0].device, params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in params):
            raise TypeError("FlatParams expects parameters of one device and dtype")
        pad = lambda n: (n + 63) // 64 * 64
        self.params = params
        self.offsets = []
        off = 0
        for p in params:
            self.offsets.append(off)
            off += pad(p.numel())
        self.total = off
        self.flat = torch.zeros(off, dtype=dt, device=dev)
        self.grad = torch.zeros(off, dtype=dt, device=dev)
        self.grad_views = []
        with torch.no_grad():
            for p, o in zip(params, self.offsets):
                n = p.numel()
                self.flat[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self.flat[o:o + n].view(p.shape)          # the module keeps its Parameter objects
                self.grad_views.append(self.grad[o:o + n].view(p.shape))
                p._lrn_flat = self
        self.attach()

    def attach(self, keep_foreign: bool = True):
        """Make every parameter's .grad the view of the flat gradient buffer.  A gradient some other code bound to
        .grad meanwhile (set_to_none, bucket views) is copied in when `keep_foreign`, and dropped otherwise."""
        for p, g in zip(self.params, self.grad_views):
            if p.grad is not g:
                if p.grad is not None and keep_foreign:
                    g.copy_(p.grad)
                p.grad = g

    @staticmethod
    def of(params):
        """The FlatParams that already owns exactly this parameter list, or None."""
        params = [p for p in params]
        owner = getattr(params[0], "_lrn_flat", None) if params else None
        if owner is not None and len(owner.params) == len(params) and all(a is b for a, b in zip(owner.params, params)):
            return owner
        return None
