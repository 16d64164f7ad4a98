"""Center loss on the B200 (loss.py:29-45), class-sharded like the A-softmax head.

    loss, grad = center_loss(features, labels, centers, alpha=0.99, weight=1.0)

mirrors `center_loss(features, labels, num_classes, alpha, weight) -> (loss, centers_update_op)`:
`centers` ([C_local, D] fp32, the non-trainable 'centers' variable, zero-initialised) is updated
in place (the reference's centers_update_op), `loss` is mean(square(features - centers[labels]))
with the pre-update centers, and `grad` is d(weight * loss)/d(features) (what the 'losses'
collection entry contributes through tf.gradients).  Runs in libasoftmax_b200.so; no fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def center_loss(features: torch.Tensor, labels: torch.Tensor, centers: torch.Tensor, alpha: float = 0.99,
                weight: float = 1.0, class_offset: int = 0, grad_accum: torch.Tensor | None = None):
    """Returns (loss, grad).  With class sharding pass the shard's `centers` slice and its
    `class_offset`; the returned loss is then that shard's partial (sum over shards = loss) and
    only rows whose label the shard owns receive a gradient.  `grad_accum` ([B, D]) is added to
    in place when given (e.g. the head's dX partial before its reduce-scatter)."""
    if not (features.is_cuda and labels.is_cuda and centers.is_cuda):
        raise RuntimeError("center_loss needs CUDA tensors (no CPU fallback)")
    if features.dtype != torch.float32 or centers.dtype != torch.float32 or not centers.is_contiguous():
        raise TypeError("features / centers must be float32, centers contiguous (updated in place)")
    if labels.dtype not in (torch.int32, torch.int64):
        raise TypeError("labels must be int32 or int64")
    X = features.contiguous()
    y = labels.contiguous()
    B, D = X.shape
    if grad_accum is not None and not (grad_accum.is_cuda and grad_accum.dtype == torch.float32
                                       and grad_accum.is_contiguous() and tuple(grad_accum.shape) == (B, D)):
        raise TypeError("grad_accum must be a contiguous CUDA float32 tensor of the features' shape [B, D]")
    lib = _lib.load()
    grad = grad_accum if grad_accum is not None else torch.zeros_like(X)
    loss = torch.empty(1, device=X.device, dtype=torch.float32)
    scratch = torch.zeros(int(lib.asm_center_scratch_bytes(B)) // 4, device=X.device, dtype=torch.float32)
    stream = C.c_void_p(torch.cuda.current_stream(X.device).cuda_stream)
    with torch.cuda.device(X.device):
        rc = lib.asm_center_loss(X.data_ptr(), B, D, y.data_ptr(), y.element_size(), centers.data_ptr(),
                                 centers.shape[0], class_offset, alpha, weight, loss.data_ptr(),
                                 grad.data_ptr(), scratch.data_ptr(), stream)
    if rc != _lib.ASM_OK:
        raise _lib.AsmError(rc, "asm_center_loss failed")
    return loss[0], grad
