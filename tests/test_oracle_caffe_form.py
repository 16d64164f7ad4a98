"""oracle/asoftmax_ref.py against an independently derived restatement: the original SphereFace
Caffe MarginInnerProduct formulation (oracle/caffe_margin_ref.py).  The two share no code and no
derivation (k-form / closed-form gradient vs. sign-form / the layer's coefficient formulas); they
must agree on the forward, the loss, the lambda schedule and the gradient w.r.t. the embeddings.
Parity with the reference itself stays unpinned (nothing in /root/reference computes this path)."""
import numpy as np
import pytest

from oracle import asoftmax_ref as ref
from oracle import caffe_margin_ref as caffe
from tf_face_toolbox_b200.synthetic import make_inputs


@pytest.mark.parametrize("lam", [0.0, 5.0, 892.857])
@pytest.mark.parametrize("B,D,C,seed", [(64, 32, 301, 3), (37, 128, 1000, 11), (256, 64, 50, 5)])
def test_forward_loss_and_embedding_gradient_agree(B, D, C, seed, lam):
    inp = make_inputs(B, D, C, seed=seed, w_std=0.05)
    X, W, y = inp.X.double().numpy(), inp.W.double().numpy(), inp.y.numpy()
    a = ref.asoftmax_head(X, W, y, 4, lam)
    layer = caffe.MarginInnerProductQuadruple()
    top = layer.forward(X, W, y, lam_override=lam)
    np.testing.assert_allclose(top, a.logits, rtol=1e-11, atol=1e-11)
    loss, top_diff = caffe.softmax_loss_and_diff(top, y)
    assert loss == pytest.approx(a.loss, rel=1e-12)
    dX = layer.backward_bottom(top_diff)
    np.testing.assert_allclose(dX, a.dX, rtol=1e-9, atol=1e-12 * np.abs(a.dX).max() + 1e-15)


def test_all_four_psi_branches_are_exercised():
    inp = make_inputs(256, 64, 50, seed=5, w_std=0.05)
    layer = caffe.MarginInnerProductQuadruple()
    layer.forward(inp.X.numpy(), inp.W.numpy(), inp.y.numpy(), lam_override=5.0)
    # s4 = 2 s0 + s3 - 3 equals -2k on the four branches (0, -2, -4, -6)
    branches = set(np.round(-layer.cache["s4"] / 2.0).astype(int))
    assert branches == {0, 1, 2, 3}, branches


def test_lambda_schedule_is_the_layers():
    layer = caffe.MarginInnerProductQuadruple()
    inp = make_inputs(8, 16, 20, seed=1, w_std=0.05)
    for it in range(1, 6):
        layer.forward(inp.X.numpy(), inp.W.numpy(), inp.y.numpy())
        assert layer.iter == it
        assert layer.lam == pytest.approx(ref.lambda_schedule(it))
    layer.iter = 10 ** 6
    layer.forward(inp.X.numpy(), inp.W.numpy(), inp.y.numpy())
    assert layer.lam == 5.0


def test_the_layers_gradient_normalisation_is_not_a_derivative():
    """Documented difference 2: with normalize_coeffs=True the bottom gradient changes length (and
    direction) on the target terms -- it is not what autodiff of the forward yields."""
    inp = make_inputs(32, 32, 100, seed=9, w_std=0.05)
    X, W, y = inp.X.double().numpy(), inp.W.double().numpy(), inp.y.numpy()
    a = ref.asoftmax_head(X, W, y, 4, 5.0)
    layer = caffe.MarginInnerProductQuadruple()
    top = layer.forward(X, W, y, lam_override=5.0)
    _, top_diff = caffe.softmax_loss_and_diff(top, y)
    exact = layer.backward_bottom(top_diff, normalize_coeffs=False)
    heur = layer.backward_bottom(top_diff, normalize_coeffs=True)
    np.testing.assert_allclose(exact, a.dX, rtol=1e-9, atol=1e-15)
    assert np.abs(heur - a.dX).max() > 1e-3 * np.abs(a.dX).max()
