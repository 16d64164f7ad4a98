"""Host-side input pipeline for the head: double-buffered asynchronous H2D copies of each
step's (pinned) inputs on a copy stream and a one-step-deferred read-back of the loss, so the
GPU never waits for the host.  Every step still copies its own inputs host->device and has its
loss read device->host -- the copies just overlap the previous step's kernels, the way the
reference overlaps its tf.data input pipeline with sess.run (data.py:195-281, train.py:228).

    runner = HostPipelinedStep(lambda X, y: asoftmax_head(X, y, C, 4, lam, weights=W)[0::2], B, D, dev)
    for X_host, y_host in batches:                 # pinned host tensors
        prev_loss = runner.submit(X_host, y_host)  # loss of the PREVIOUS step (None at first)
    last_loss = runner.flush()
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


class HostPipelinedStep:
    def __init__(self, step_fn: Callable, batch: int, dim: int, device, labels_dtype=torch.int32,
                 loss_stream: bool = False, read_dx: bool = False):
        """loss_stream=True reads the loss back on a dedicated stream behind the step-end event
        instead of on the compute stream, so the 4-byte DMA no longer sits between two steps
        (opt-in until it has been measured on hardware)."""
        self.step_fn = step_fn
        self.dev = torch.device(device)
        self.X = [torch.empty(batch, dim, device=self.dev, dtype=torch.float32) for _ in range(2)]
        self.y = [torch.empty(batch, device=self.dev, dtype=labels_dtype) for _ in range(2)]
        self.loss_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        # read_dx: the step's dX (what a host-side consumer of the head gets back) is read D2H too
        self.dx_host = [torch.zeros(batch, dim, dtype=torch.float32).pin_memory() for _ in range(2)] if read_dx else None
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.d2h_stream = torch.cuda.Stream(self.dev) if loss_stream else None
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.loss_ready = [torch.cuda.Event() for _ in range(2)]
        self.n = 0
        self.out = [None, None]

    def submit(self, X_host: torch.Tensor, y_host: torch.Tensor) -> Optional[float]:
        i = self.n & 1
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.copy_stream):
            if self.n >= 2:
                self.copy_stream.wait_event(self.consumed[i])      # step n-2 is done with this buffer
            self.X[i].copy_(X_host, non_blocking=True)
            self.y[i].copy_(y_host, non_blocking=True)
            self.copied[i].record(self.copy_stream)
        cur.wait_event(self.copied[i])
        out = self.step_fn(self.X[i], self.y[i])
        self.out[i] = out
        self.consumed[i].record(cur)
        if self.d2h_stream is None:
            if self.dx_host is not None:
                self.dx_host[i].copy_(out[1], non_blocking=True)
            self.loss_host[i].copy_(out[0].reshape(1), non_blocking=True)
            self.loss_ready[i].record(cur)
        else:
            # `out` stays referenced in self.out[i] until step n+2, i.e. after loss_ready[i] has
            # been synchronised at step n+1, so the caching allocator cannot recycle it early
            with torch.cuda.stream(self.d2h_stream):
                self.d2h_stream.wait_event(self.consumed[i])
                if self.dx_host is not None:
                    self.dx_host[i].copy_(out[1], non_blocking=True)
                self.loss_host[i].copy_(out[0].reshape(1), non_blocking=True)
                self.loss_ready[i].record(self.d2h_stream)
        prev = None
        if self.n >= 1:
            self.loss_ready[i ^ 1].synchronize()                   # previous step's loss (D2H done)
            prev = float(self.loss_host[i ^ 1][0])
        self.n += 1
        return prev

    def flush(self) -> Optional[float]:
        if self.n == 0:
            return None
        i = (self.n - 1) & 1
        self.loss_ready[i].synchronize()
        return float(self.loss_host[i][0])
