"""Property tests of the oracle (hypothesis): the size-independent identities the GPU parity
tests lean on at full size, checked here over random shapes, margins, lambdas and shard cuts."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import asoftmax_ref as ref

SET = dict(max_examples=25, deadline=None)


def problem(seed, B, D, C):
    g = np.random.default_rng(seed)
    W = g.normal(0, 0.05, (D, C))
    y = g.integers(0, C, B)
    What = W / np.sqrt((W * W).sum(0))
    a = np.array([1.5, 0.3, -0.3, -1.5])[np.arange(B) % 4]      # all psi branches
    X = a[:, None] * What[:, y].T * np.sqrt(D) + g.normal(0, 1, (B, D))
    return X, W, y


@settings(**SET)
@given(seed=st.integers(0, 10_000), B=st.integers(1, 9), D=st.integers(2, 12), C=st.integers(2, 17),
       m=st.integers(1, 4), lam=st.sampled_from([0.0, 0.7, 5.0, 892.86]), cuts=st.integers(1, 5))
def test_any_class_sharding_reproduces_the_unsharded_loss(seed, B, D, C, m, lam, cuts):
    X, W, y = problem(seed, B, D, C)
    full = ref.asoftmax_head(X, W, y, m, lam)
    g = np.random.default_rng(seed + 1)
    bounds = sorted(set([0, C] + list(g.integers(1, C, min(cuts, C - 1))) if C > 1 else [0, C]))
    stats = []
    for lo, hi in zip(bounds, bounds[1:]):
        mloc, zloc, fy, owned = ref.sharded_partial_stats(X, W[:, lo:hi], y, lo, m, lam)
        stats.append((mloc, zloc, fy))
    M, logZ, loss = ref.sharded_combine(stats)
    assert abs(loss - full.loss) <= 1e-10 * max(1.0, abs(full.loss))
    np.testing.assert_allclose(M + logZ, full.row_max + full.row_logz, rtol=1e-12, atol=1e-12)


@settings(**SET)
@given(seed=st.integers(0, 10_000), B=st.integers(1, 8), D=st.integers(2, 10), C=st.integers(2, 12),
       m=st.integers(1, 4), lam=st.sampled_from([0.0, 1.0, 5.0]))
def test_column_scaling_leaves_loss_and_dx_and_rescales_dw(seed, B, D, C, m, lam):
    X, W, y = problem(seed, B, D, C)
    alpha = np.random.default_rng(seed + 2).uniform(0.2, 5.0, C)
    r0 = ref.asoftmax_head(X, W, y, m, lam)
    r1 = ref.asoftmax_head(X, W * alpha, y, m, lam)
    assert abs(r0.loss - r1.loss) <= 1e-10 * max(1.0, abs(r0.loss))
    np.testing.assert_allclose(r1.dX, r0.dX, rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(r1.dW * alpha, r0.dW, rtol=1e-8, atol=1e-12)
    # normalisation Jacobian: every dW column is orthogonal to its weight column
    assert np.abs((r0.dW * W).sum(0)).max() <= 1e-10 * max(1.0, np.abs(r0.dW).max() * np.abs(W).max() * D)


@settings(**SET)
@given(seed=st.integers(0, 10_000), B=st.integers(2, 8), D=st.integers(2, 10), C=st.integers(2, 12),
       m=st.integers(1, 4), lam=st.sampled_from([0.0, 5.0]))
def test_loss_is_the_mean_over_rows_and_gradients_add_up(seed, B, D, C, m, lam):
    """Per-tower means over B/G rows, scaled 1/G and summed (data_parallel.py:37, 179, 248) equal
    the global-batch mean: checked by splitting the batch in two."""
    X, W, y = problem(seed, B, D, C)
    h = B // 2
    full = ref.asoftmax_head(X, W, y, m, lam)
    a = ref.asoftmax_head(X[:h], W, y[:h], m, lam)
    b = ref.asoftmax_head(X[h:], W, y[h:], m, lam)
    wa, wb = h / B, (B - h) / B
    assert abs(wa * a.loss + wb * b.loss - full.loss) <= 1e-10 * max(1.0, abs(full.loss))
    np.testing.assert_allclose(wa * a.dW + wb * b.dW, full.dW, rtol=1e-8, atol=1e-12)
    np.testing.assert_allclose(np.concatenate([wa * a.dX, wb * b.dX]), full.dX, rtol=1e-8, atol=1e-12)
