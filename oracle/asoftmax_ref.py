"""CPU oracle for the A-softmax (SphereFace angular-margin) classification head.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (tf_face_toolbox_b200/) may
import this module; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs use it, and only as the checker or the timed CPU baseline.

PARITY UNPINNED.  The reference snapshot (/root/reference) does not contain the A-softmax
code (loss.py:18,29,47 define only focal / center / triplet losses; README.md:14,19 only
claim the feature), ships no tests or golden vectors, and cannot be executed here
(TensorFlow r1.8 contrib + Python 2, not installed, no network; data_parallel.py:19 imports
a module missing from the snapshot).  This file is therefore a float64 NumPy restatement of
the *published* SphereFace A-softmax definition (Liu et al., CVPR 2017), written against
the conventions the reference does pin:

  * classifier weight W is [D, C] (in, out), no bias         nets/sphere.py:84-90
  * loss = mean over the batch of sparse softmax CE           nets/sphere.py:109-111
  * labels are int32 [B]                                      data.py:259,271
  * labels are passed into forward (margin head)              data_parallel.py:220
  * tower grads x 1/num_gpus then nccl.all_sum  => gradient of the GLOBAL-batch mean
                                                              data_parallel.py:37,179,248
  * the lambda-annealing clock is global_step                 train.py:157, data_parallel.py:252-253

Implementation-defined choices fixed here: no epsilon in the norms;
t clipped to [-1,1] before psi; gradients flow through both norms; k / sign terms are
piecewise constants; lambda is a host scalar; returned logits are the margin-modified f.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


# --------------------------------------------------------------------------------------
# lambda annealing (SphereFace schedule; clock = global_step, train.py:157)
# --------------------------------------------------------------------------------------
def lambda_schedule(iteration: int, base: float = 1000.0, gamma: float = 0.12,
                    power: float = 1.0, lambda_min: float = 5.0) -> float:
    """lambda(it) = max(lambda_min, base * (1 + gamma*it)^(-power)); `it` counts from 1."""
    return max(lambda_min, base * (1.0 + gamma * iteration) ** (-power))


# --------------------------------------------------------------------------------------
# psi(theta) = (-1)^k cos(m theta) - 2k,  theta in [k pi/m, (k+1) pi/m]
# --------------------------------------------------------------------------------------
def chebyshev(t: np.ndarray, m: int):
    """T_m(t) and T_m'(t) for m in 1..4 (cos(m theta) as a polynomial in t = cos theta)."""
    if m == 1:
        return t, np.ones_like(t)
    if m == 2:
        return 2 * t * t - 1, 4 * t
    if m == 3:
        return 4 * t ** 3 - 3 * t, 12 * t * t - 3
    if m == 4:
        return 8 * t ** 4 - 8 * t * t + 1, 32 * t ** 3 - 16 * t
    raise ValueError("m must be in {1,2,3,4}")


def psi_kform(t: np.ndarray, m: int):
    """Generic-m psi via k = min(floor(m*acos(t)/pi), m-1). Returns (psi, dpsi/dt, k)."""
    t = np.clip(np.asarray(t, dtype=np.float64), -1.0, 1.0)
    k = np.minimum(np.floor(m * np.arccos(t) / math.pi), m - 1)
    sgn = np.where(k % 2 == 0, 1.0, -1.0)
    T, dT = chebyshev(t, m)
    return sgn * T - 2.0 * k, sgn * dT, k.astype(np.int64)


def psi4_signform(t: np.ndarray):
    """m=4 by sign tests (the Caffe formulation): s0=sign(t), s3=s0*sign(2t^2-1),
    s4=2*s0+s3-3, psi=s3*T4(t)+s4.  sign(0)=0 keeps psi continuous at the branch points."""
    t = np.clip(np.asarray(t, dtype=np.float64), -1.0, 1.0)
    s0 = np.sign(t)
    s3 = s0 * np.sign(2 * t * t - 1)
    s4 = 2 * s0 + s3 - 3
    T, dT = chebyshev(t, 4)
    return s3 * T + s4, s3 * dT


# --------------------------------------------------------------------------------------
# forward / backward
# --------------------------------------------------------------------------------------
@dataclass
class HeadResult:
    loss: float
    logits: np.ndarray      # f  [B, C]  margin-modified logits (what the softmax sees)
    dX: np.ndarray          # [B, D]
    dW: np.ndarray          # [D, C]
    n: np.ndarray           # [B] embedding norms
    c: np.ndarray           # [C] class-weight column norms
    t: np.ndarray           # [B] cos(theta) on the target column
    psi: np.ndarray         # [B]
    row_max: np.ndarray     # [B]
    row_logz: np.ndarray    # [B] log sum exp(f - row_max)


def asoftmax_head(X, W, y, m: int = 4, lam: float = 0.0, dtype=np.float64) -> HeadResult:
    """Forward + closed-form backward of the A-softmax head.

    X [B,D] embeddings, W [D,C] class weights (reference layout nets/sphere.py:86),
    y [B] integer labels, m margin, lam = lambda >= 0 (host scalar).
    Loss is the mean over the batch (nets/sphere.py:109); dX, dW are d(loss)/d(X,W).
    """
    X = np.asarray(X, dtype=dtype)
    W = np.asarray(W, dtype=dtype)
    y = np.asarray(y).astype(np.int64)
    B, D = X.shape
    D2, C = W.shape
    assert D == D2 and y.shape == (B,)
    if y.min() < 0 or y.max() >= C:
        raise ValueError("label out of range")
    rows = np.arange(B)

    n = np.sqrt((X * X).sum(axis=1))                 # [B]
    c = np.sqrt((W * W).sum(axis=0))                 # [C]
    What = W / c                                     # [D,C]
    S = X @ What                                     # s_ij = n_i cos(theta_ij)
    s_y = S[rows, y]
    t = np.clip(s_y / n, -1.0, 1.0)
    psi, dpsi, _ = psi_kform(t, m)
    psi = psi.astype(dtype)
    dpsi = dpsi.astype(dtype)
    f = S.copy()
    f[rows, y] = (lam * s_y + n * psi) / (1.0 + lam)

    row_max = f.max(axis=1)
    E = np.exp(f - row_max[:, None])
    Z = E.sum(axis=1)
    loss = float(np.mean(np.log(Z) + row_max - f[rows, y]))

    g = E / Z[:, None]
    g[rows, y] -= 1.0
    g /= B
    g_y = g[rows, y].copy()
    Gp = g                                           # G' (in place)
    Gp[rows, y] = g_y * (lam + dpsi) / (1.0 + lam)
    r = g_y * (psi - t * dpsi) / ((1.0 + lam) * n)
    dX = Gp @ What.T + r[:, None] * X
    dWhat = X.T @ Gp
    q = (Gp * S).sum(axis=0)                         # = what_j . dWhat_j
    dW = (dWhat - What * q) / c
    return HeadResult(loss, f, dX, dW, n, c, t, psi, row_max, np.log(Z))


@dataclass
class StreamedResult:
    loss: float
    dX: np.ndarray          # [B, D]
    dW: dict                # (lo, hi) -> dW[:, lo:hi] for every requested class range
    row_max: np.ndarray
    row_logz: np.ndarray


def asoftmax_head_streamed(X, W, y, m: int = 4, lam: float = 0.0, chunk: int = 16384,
                           dw_ranges=(), dtype=np.float64) -> StreamedResult:
    """The same head evaluated in class chunks, never holding a [B, C] matrix: for the
    BASELINE shapes whose logits do not fit in host memory (config 4: 1024 x 1,000,000) or are
    slow to check in one piece (config 5: 2048 x 85,742).  Pass 1 keeps an online (max, sum-exp)
    per row; pass 2 rebuilds each chunk's G' from it and accumulates dX; dW is produced only for
    the class ranges asked for.  Same formulas as asoftmax_head (checked equal in
    tests/test_oracle.py); `dtype` is the matmul precision (float32 halves the time, the
    statistics stay float64)."""
    X64 = np.asarray(X, dtype=np.float64)
    Xc = X64.astype(dtype)
    y = np.asarray(y).astype(np.int64)
    B, D = Xc.shape
    C = W.shape[1]
    if y.min() < 0 or y.max() >= C:
        raise ValueError("label out of range")
    n = np.sqrt((X64 * X64).sum(axis=1))
    bounds = [(lo, min(C, lo + chunk)) for lo in range(0, C, chunk)]

    def chunk_logits(lo, hi):
        Wc = np.asarray(W[:, lo:hi], dtype=np.float64)
        c = np.sqrt((Wc * Wc).sum(axis=0))
        What = (Wc / c).astype(dtype)
        return (Xc @ What).astype(np.float64), What, c

    # target column first: t, psi, f_y (the target's own chunk is rebuilt in pass 2)
    Wy = np.asarray(W[:, y], dtype=np.float64)                 # [D, B]
    cy = np.sqrt((Wy * Wy).sum(axis=0))
    s_y = np.einsum("bd,db->b", X64, Wy / cy)
    t = np.clip(s_y / n, -1.0, 1.0)
    psi, dpsi, _ = psi_kform(t, m)
    f_y = (lam * s_y + n * psi) / (1.0 + lam)

    M = np.full(B, -np.inf)
    Z = np.zeros(B)
    for lo, hi in bounds:                                      # pass 1: online log-sum-exp
        f, _, _ = chunk_logits(lo, hi)
        own = np.nonzero((y >= lo) & (y < hi))[0]
        f[own, y[own] - lo] = f_y[own]
        mc = f.max(axis=1)
        Mn = np.maximum(M, mc)
        Z = Z * np.exp(M - Mn) + np.exp(f - Mn[:, None]).sum(axis=1)
        M = Mn
    lse = M + np.log(Z)
    loss = float(np.mean(lse - f_y))

    dX = np.zeros((B, D))
    dW = {}
    for lo, hi in bounds:                                      # pass 2: gradients
        S, What, c = chunk_logits(lo, hi)
        own = np.nonzero((y >= lo) & (y < hi))[0]
        f = S.copy()
        f[own, y[own] - lo] = f_y[own]
        Gp = np.exp(f - lse[:, None]) / B
        g_y = Gp[own, y[own] - lo] - 1.0 / B
        Gp[own, y[own] - lo] = g_y * (lam + dpsi[own]) / (1.0 + lam)
        dX += (Gp.astype(dtype) @ What.T).astype(np.float64)
        dX[own] += (g_y * (psi[own] - t[own] * dpsi[own]) / ((1.0 + lam) * n[own]))[:, None] * X64[own]
        for (rlo, rhi) in dw_ranges:
            a, b = max(lo, rlo), min(hi, rhi)
            if a >= b:
                continue
            sl = slice(a - lo, b - lo)
            dWhat = (Xc.T @ Gp[:, sl].astype(dtype)).astype(np.float64)
            q = (Gp[:, sl] * S[:, sl]).sum(axis=0)
            dW.setdefault((rlo, rhi), np.zeros((D, rhi - rlo)))[:, a - rlo:b - rlo] = \
                (dWhat - What[:, sl].astype(np.float64) * q) / c[sl]
    return StreamedResult(loss, dX, dW, M, np.log(Z))


def asoftmax_loss_only(X, W, y, m=4, lam=0.0, dtype=np.float64) -> float:
    """Loss without gradients (for finite differences)."""
    X = np.asarray(X, dtype=dtype)
    W = np.asarray(W, dtype=dtype)
    y = np.asarray(y).astype(np.int64)
    rows = np.arange(X.shape[0])
    n = np.sqrt((X * X).sum(axis=1))
    c = np.sqrt((W * W).sum(axis=0))
    S = X @ (W / c)
    s_y = S[rows, y]
    t = np.clip(s_y / n, -1.0, 1.0)
    psi, _, _ = psi_kform(t, m)
    f = S
    f[rows, y] = (lam * s_y + n * psi) / (1.0 + lam)
    mx = f.max(axis=1)
    return float(np.mean(np.log(np.exp(f - mx[:, None]).sum(axis=1)) + mx - f[rows, y]))


# --------------------------------------------------------------------------------------
# class-sharded evaluation: what G ranks compute and exchange
# --------------------------------------------------------------------------------------
def shard_bounds(C: int, G: int):
    """GPU g owns classes [g*ceil(C/G), min(C,(g+1)*ceil(C/G)))."""
    per = -(-C // G)
    return [(min(C, g * per), min(C, (g + 1) * per)) for g in range(G)]


def sharded_partial_stats(X, W_shard, y, class_offset, m=4, lam=0.0):
    """Per-shard forward: returns (local max [B], local sumexp [B], target logit or 0 [B],
    owned mask [B]) -- the three floats per row each rank contributes to the exchange."""
    X = np.asarray(X, dtype=np.float64)
    W_shard = np.asarray(W_shard, dtype=np.float64)
    y = np.asarray(y).astype(np.int64)
    B = X.shape[0]
    Cl = W_shard.shape[1]
    n = np.sqrt((X * X).sum(axis=1))
    c = np.sqrt((W_shard * W_shard).sum(axis=0))
    f = X @ (W_shard / c)
    yl = y - class_offset
    owned = (yl >= 0) & (yl < Cl)
    ro = np.nonzero(owned)[0]
    s_y = f[ro, yl[ro]]
    t = np.clip(s_y / n[ro], -1.0, 1.0)
    psi, _, _ = psi_kform(t, m)
    f[ro, yl[ro]] = (lam * s_y + n[ro] * psi) / (1.0 + lam)
    mloc = f.max(axis=1) if Cl > 0 else np.full(B, -np.inf)
    zloc = np.exp(f - mloc[:, None]).sum(axis=1) if Cl > 0 else np.zeros(B)
    fy = np.zeros(B)
    fy[ro] = f[ro, yl[ro]]
    return mloc, zloc, fy, owned


def sharded_combine(stats):
    """Combine per-rank (max, sumexp, target-or-0): M=max_g m_g, Z=sum_g z_g e^{m_g-M},
    loss = mean(log Z + M - f_y)."""
    ms = np.stack([s[0] for s in stats])
    zs = np.stack([s[1] for s in stats])
    fy = np.sum(np.stack([s[2] for s in stats]), axis=0)
    M = ms.max(axis=0)
    Z = (zs * np.exp(ms - M)).sum(axis=0)
    return M, np.log(Z), float(np.mean(np.log(Z) + M - fy))


# --------------------------------------------------------------------------------------
# classifier optimizer step (data_parallel.py:186-196) on the L2-regularised gradient
# --------------------------------------------------------------------------------------
def optimizer_step(W, dW, state0, state1, kind="momentum", lr=0.1, momentum=0.9, beta1=0.5,
                   beta2=0.999, epsilon=1e-8, weight_decay=5e-4, step=1):
    """Returns (W_new, state0_new, state1_new).  TF semantics: MomentumOptimizer
    (accum = momentum*accum + g; var -= lr*accum) and AdamOptimizer
    (lr_t = lr*sqrt(1-b2^t)/(1-b1^t); var -= lr_t*m/(sqrt(v)+eps)); g = dW + wd*W is the
    gradient of cross_entropy + reg_loss (nets/net_base.py:103-107, nets/sphere.py:88)."""
    W = np.asarray(W, dtype=np.float64)
    g = np.asarray(dW, dtype=np.float64) + weight_decay * W
    if kind == "momentum":
        s0 = momentum * np.asarray(state0, dtype=np.float64) + g
        return W - lr * s0, s0, None
    m = beta1 * np.asarray(state0, dtype=np.float64) + (1 - beta1) * g
    v = beta2 * np.asarray(state1, dtype=np.float64) + (1 - beta2) * g * g
    lr_t = lr * math.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
    return W - lr_t * m / (np.sqrt(v) + epsilon), m, v


# --------------------------------------------------------------------------------------
# center loss (loss.py:29-45)
# --------------------------------------------------------------------------------------
def center_loss(features, labels, centers, alpha=0.99, weight=1.0):
    """Restates loss.py:29-45: centers_batch = gather(centers, labels); diffs = (1-alpha) *
    (centers_batch - features); centers = scatter_sub(centers, labels, diffs) (duplicates
    accumulate); loss = mean(square(features - centers_batch)) with the PRE-update centers.
    Returns (loss, centers_new, d(weight*loss)/d(features))."""
    X = np.asarray(features, dtype=np.float64)
    y = np.asarray(labels).astype(np.int64)
    Cn = np.asarray(centers, dtype=np.float64)
    cb = Cn[y]
    diffs = (1.0 - alpha) * (cb - X)
    new = Cn.copy()
    np.subtract.at(new, y, diffs)
    loss = float(np.mean((X - cb) ** 2))
    grad = weight * 2.0 * (X - cb) / X.size
    return loss, new, grad
