"""Second, independently derived restatement of the A-softmax head: the ORIGINAL SphereFace
formulation (Liu et al., CVPR 2017; the Caffe `MarginInnerProduct` layer of the authors' public
release, type QUADRUPLE, m = 4) in NumPy float64.

TEST INFRASTRUCTURE ONLY (same rule as oracle/asoftmax_ref.py).  PARITY STAYS UNPINNED: the
layer's source is not in /root/reference (nor anywhere in this container: no network), so this
file restates the published algorithm from its description -- the sign-function form of psi built
from cos_theta_quadratic / cos_theta_quartic, weights normalised to unit L2 norm before the
product, the lambda blend applied to the whole top blob, the hand-written bottom gradient with
its two coefficients -- and is used to check oracle/asoftmax_ref.py against a formulation that
shares no code and no derivation with it (tests/test_oracle_caffe_form.py):

  forward   top[i, j] = x_i . what_j                                   (j != y_i)
            top[i, y] = (|x_i| (s3 (8 c^4 - 8 c^2 + 1) + s4) + lambda x_i . what_y) / (1 + lambda)
            with c = cos(theta_iy), s0 = sign(c), s3 = s0 sign(2 c^2 - 1), s4 = 2 s0 + s3 - 3
  lambda    iter += 1 at the start of every forward;  lambda = max(lambda_min, base (1 + gamma iter)^-power)
  backward  bottom_diff_i = sum_{j != y} top_diff_ij what_j
                            + top_diff_iy / (1 + lambda) * (coeff_w what_y + coeff_x x_i) + lambda-part
            coeff_w = s3 (32 c^3 - 16 c)                                (= psi')
            coeff_x = (s3 (-24 c^4 + 8 c^2 + 1) + s4) / |x_i|           (= (psi - c psi') / |x|)

Known, documented differences between that layer and what the TensorFlow reference computes by
automatic differentiation (and therefore what oracle/asoftmax_ref.py and the CUDA path compute):
  1. The layer renormalises the weight blob in place and its weight gradient is the plain
     inner-product gradient top_diff^T x: it does NOT differentiate through the weight norm and
     ignores the margin on the target column.  tf.gradients does both.  Only the forward and the
     gradient w.r.t. the embeddings are comparable, and only those are compared.
  2. The released layer rescales (coeff_w, coeff_x) to unit length before using them (a
     gradient-normalisation heuristic, not a derivative).  `normalize_coeffs=True` reproduces
     that; the default False is the exact derivative, which is what autodiff yields.
"""
from __future__ import annotations

import numpy as np


class MarginInnerProductQuadruple:
    """Stateful like the layer: owns iter_ and the lambda schedule parameters."""

    def __init__(self, base=1000.0, gamma=0.12, power=1.0, lambda_min=5.0):
        self.base, self.gamma, self.power, self.lambda_min = base, gamma, power, lambda_min
        self.iter = 0
        self.lam = None

    def _advance(self, lam_override=None):
        self.iter += 1
        lam = self.base * (1.0 + self.gamma * self.iter) ** (-self.power)
        self.lam = max(lam, self.lambda_min) if lam_override is None else float(lam_override)
        return self.lam

    def forward(self, bottom, weight_kn, label, lam_override=None):
        """bottom [M, K] embeddings, weight_kn [K, N] in the REFERENCE's [in, out] layout (the layer
        itself stores [N, K]; transposed here once), label [M].  Returns top [M, N]."""
        lam = self._advance(lam_override)
        x = np.asarray(bottom, dtype=np.float64)
        w = np.asarray(weight_kn, dtype=np.float64).T.copy()           # [N, K] rows = classes
        w /= np.sqrt((w * w).sum(axis=1, keepdims=True))               # normalise every class row
        y = np.asarray(label).astype(np.int64)
        M = x.shape[0]
        rows = np.arange(M)
        x_norm = np.sqrt((x * x).sum(axis=1))
        ip = x @ w.T                                                    # x'w for every class
        cos_theta = np.clip(ip[rows, y] / x_norm, -1.0, 1.0)
        sign_0 = np.sign(cos_theta)
        cos_quadratic = cos_theta * cos_theta
        cos_cubic = cos_quadratic * cos_theta
        cos_quartic = cos_quadratic * cos_quadratic
        sign_3 = sign_0 * np.sign(2.0 * cos_quadratic - 1.0)
        sign_4 = 2.0 * sign_0 + sign_3 - 3.0
        top = ip.copy()
        top[rows, y] = x_norm * (sign_3 * (8.0 * cos_quartic - 8.0 * cos_quadratic + 1.0) + sign_4)
        top = (top + lam * ip) / (1.0 + lam)                            # + lambda x'w, then / (1 + lambda)
        self.cache = dict(x=x, w=w, y=y, x_norm=x_norm, cos=cos_theta, cos2=cos_quadratic, cos3=cos_cubic,
                          cos4=cos_quartic, s3=sign_3, s4=sign_4, lam=lam)
        return top

    def backward_bottom(self, top_diff, normalize_coeffs=False):
        """Gradient w.r.t. the embeddings, as the layer's Backward computes it."""
        c = self.cache
        g = np.asarray(top_diff, dtype=np.float64)
        M = g.shape[0]
        rows = np.arange(M)
        lam = c["lam"]
        g_y = g[rows, c["y"]]
        g_rest = g.copy()
        g_rest[rows, c["y"]] = 0.0
        # every non-target entry is (1 + lambda) x'w / (1 + lambda) = x'w: plain inner-product gradient
        bottom_diff = g_rest @ c["w"]
        coeff_w = c["s3"] * (32.0 * c["cos3"] - 16.0 * c["cos"])
        coeff_x = (c["s3"] * (-24.0 * c["cos4"] + 8.0 * c["cos2"] + 1.0) + c["s4"]) / c["x_norm"]
        if normalize_coeffs:
            nrm = np.sqrt(coeff_w * coeff_w + coeff_x * coeff_x)
            coeff_w, coeff_x = coeff_w / nrm, coeff_x / nrm
        w_y = c["w"][c["y"]]
        bottom_diff += (g_y / (1.0 + lam) * coeff_w)[:, None] * w_y
        bottom_diff += (g_y / (1.0 + lam) * coeff_x)[:, None] * c["x"]
        bottom_diff += (g_y * lam / (1.0 + lam))[:, None] * w_y          # the lambda x'w part of the target
        return bottom_diff


def softmax_loss_and_diff(top, label):
    """SoftmaxWithLoss (normalisation VALID = mean over the batch) and its top_diff."""
    t = np.asarray(top, dtype=np.float64)
    y = np.asarray(label).astype(np.int64)
    M = t.shape[0]
    rows = np.arange(M)
    mx = t.max(axis=1, keepdims=True)
    e = np.exp(t - mx)
    p = e / e.sum(axis=1, keepdims=True)
    loss = float(np.mean(-np.log(p[rows, y])))
    d = p.copy()
    d[rows, y] -= 1.0
    return loss, d / M
