"""CUDA-graph replay of the whole eval forward for a fixed (B, N): the reference's whole-scene loop runs the
model once per lane line with B = 1 (inference_whole_scene.py:130-139), where ~100 small launches, not the GPU,
set the latency.  Capture happens once; every call copies the inputs into static buffers and replays."""
from __future__ import annotations

import torch


class GraphedLineRefineNet:
    def __init__(self, model, B: int, N: int, M: int = 32):
        if model.training:
            raise RuntimeError("CUDA-graph replay is for eval-mode inference")
        dev = next(model.parameters()).device
        self.model = model
        self.ctx = torch.zeros(B, N, 4, dtype=torch.float32, device=dev)
        self.line = torch.zeros(B, M, 3, dtype=torch.float32, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.no_grad(), torch.cuda.stream(side):      # warm-up: folds weights, sizes workspaces, sets kernel attributes
            for _ in range(3):
                model(self.ctx, self.line)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = model(self.ctx, self.line)

    @torch.no_grad()
    def __call__(self, context: torch.Tensor, noisy_line: torch.Tensor) -> torch.Tensor:
        self.ctx.copy_(context, non_blocking=True)
        self.line.copy_(noisy_line, non_blocking=True)
        self.graph.replay()
        return self.out
