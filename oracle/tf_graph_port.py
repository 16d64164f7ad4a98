"""Op-by-op CPU port of the graph a TensorFlow A-softmax head would execute.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/asoftmax_ref.py header; parity unpinned).
The reference's TensorFlow CPU path cannot run here or on the GPU box (no TensorFlow,
Python 2 code, missing module data_parallel.py:19), so this stand-in executes the same op
graph -- matmul, column/row norms, gather / scatter of the target column, psi(theta)
margin + lambda blend, mean sparse softmax cross-entropy, autograd backward -- in torch
CPU fp32, using every host thread.  It follows the conventions the reference pins:
W [D,C] no bias (nets/sphere.py:84-90), mean CE (nets/sphere.py:109-111), int32 labels
(data.py:259), gradient of the global-batch mean (data_parallel.py:37,179).
bench.py times it as `cpu_baseline` (kind "port") and as `--impl reference`.
"""
from __future__ import annotations

import math

import torch


def _cheb(t: torch.Tensor, m: int) -> torch.Tensor:
    if m == 1:
        return t
    if m == 2:
        return 2 * t * t - 1
    if m == 3:
        return 4 * t ** 3 - 3 * t
    if m == 4:
        return 8 * t ** 4 - 8 * t * t + 1
    raise ValueError("m must be in {1,2,3,4}")


def asoftmax_graph(X: torch.Tensor, W: torch.Tensor, y: torch.Tensor, m: int = 4,
                   lam: float = 0.0):
    """Forward graph; returns (loss, logits). Differentiable w.r.t. X and W."""
    B = X.shape[0]
    rows = torch.arange(B)
    y = y.long()
    n = torch.linalg.vector_norm(X, dim=1)                    # tf.norm(features, axis=1)
    c = torch.linalg.vector_norm(W, dim=0, keepdim=True)      # tf.norm(weights, axis=0)
    S = X @ (W / c)                                           # tf.matmul
    s_y = S[rows, y]                                          # tf.gather_nd
    t = torch.clamp(s_y / n, -1.0, 1.0)
    with torch.no_grad():                                     # k is piecewise constant
        k = torch.clamp(torch.floor(m * torch.acos(t) / math.pi), max=m - 1)
        sgn = 1.0 - 2.0 * torch.remainder(k, 2)
    psi = sgn * _cheb(t, m) - 2.0 * k
    f_y = (lam * s_y + n * psi) / (1.0 + lam)
    f = S.index_put((rows, y), f_y)                           # tf.scatter_nd blend
    loss = torch.nn.functional.cross_entropy(f, y)            # mean sparse softmax CE
    return loss, f


def step(X: torch.Tensor, W: torch.Tensor, y: torch.Tensor, m: int = 4, lam: float = 0.0):
    """One fwd+bwd pass: returns (loss, dX, dW) like tf.gradients(total_loss, params)."""
    Xr = X.detach().clone().requires_grad_(True)
    Wr = W.detach().clone().requires_grad_(True)
    loss, _ = asoftmax_graph(Xr, Wr, y, m, lam)
    loss.backward()
    return loss.detach(), Xr.grad, Wr.grad
