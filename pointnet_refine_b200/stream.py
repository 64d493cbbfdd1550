"""Host -> device -> host streaming of the encoder over a large batch of segments (the shape of the reference's
whole-scene sweep, inference_whole_scene.py:299-333, batched): segment chunks are copied on a side stream into
double-buffered device staging while the previous chunk is being encoded, and the pooled features are copied
back asynchronously.  Segments are independent, so chunking does not change any result."""
from __future__ import annotations

import torch

from .ops import DEFAULT_CHUNK_ROWS


class HostEncoderPipeline:
    def __init__(self, encoder, segments_per_chunk: int = 512):
        self.enc = encoder
        self.chunk = int(segments_per_chunk)
        self._stage = None
        self._copy = None

    def _buffers(self, n_seg, n_pts, dev):
        key = (n_seg, n_pts, dev)
        if self._stage is None or self._stage[0] != key:
            self._stage = (key, [torch.empty(n_seg, n_pts, 4, dtype=torch.float32, device=dev) for _ in range(2)])
            self._copy = torch.cuda.Stream(device=dev)
        return self._stage[1]

    @torch.no_grad()
    def global_feat(self, host_context: torch.Tensor, host_out: torch.Tensor | None = None) -> torch.Tensor:
        """host_context: pinned (B, N, 4) fp32 CPU tensor.  Returns the (B, 2048) pooled features in `host_out`
        (pinned CPU tensor, allocated if None).  Synchronises the device before returning."""
        if host_context.is_cuda or host_context.dtype != torch.float32:
            raise TypeError("host_context must be a float32 CPU tensor (pinned for asynchronous copies)")
        dev = next(self.enc.parameters()).device
        B, N, _ = host_context.shape
        if host_out is None:
            host_out = torch.empty(B, 2048, dtype=torch.float32).pin_memory()
        chunk = min(self.chunk, B)
        per_wave = DEFAULT_CHUNK_ROWS // N        # segments of one full encoder wave (74 at N = 4096)
        if per_wave >= 1 and chunk >= per_wave:
            chunk = chunk // per_wave * per_wave   # whole waves per host chunk: no partially filled launch in the middle
        stage = self._buffers(chunk, N, dev)
        compute = torch.cuda.current_stream(dev)
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        starts = list(range(0, B, chunk))

        def issue_copy(i):
            s, slot = starts[i], i & 1
            n = min(chunk, B - s)
            with torch.cuda.stream(self._copy):
                if i >= 2:
                    self._copy.wait_event(consumed[slot])        # the chunk that used this slot has been encoded
                stage[slot][:n].copy_(host_context[s:s + n], non_blocking=True)
                copied[slot].record(self._copy)

        issue_copy(0)
        for i, s in enumerate(starts):
            slot, n = i & 1, min(chunk, B - s)
            if i + 1 < len(starts):
                issue_copy(i + 1)
            compute.wait_event(copied[slot])
            gf = self.enc.run_native(stage[slot][:n], pool=True)["global_feat"]
            consumed[slot].record(compute)
            host_out[s:s + n].copy_(gf, non_blocking=True)
        torch.cuda.synchronize(dev)
        return host_out
