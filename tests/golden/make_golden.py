"""Generates tests/golden/asoftmax_small.npz.

The reference holds no golden vectors for this path (parity unpinned), and its
TensorFlow code cannot be imported here, so these fixtures are produced by an INDEPENDENT
evaluation -- torch float64 autograd over the op graph in oracle/tf_graph_port.py -- not by
the closed-form oracle they are used to check.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import tf_graph_port as port                      # noqa: E402
from tf_face_toolbox_b200.synthetic import make_inputs        # noqa: E402

CASES = [  # name, B, D, C, m, lam, seed
    ("m4_lam5", 16, 32, 50, 4, 5.0, 11),
    ("m4_lam0", 16, 32, 50, 4, 0.0, 12),
    ("m4_lam892", 8, 64, 33, 4, 1000 / 1.12, 13),
    ("m3_lam5", 8, 16, 21, 3, 5.0, 14),
    ("m2_lam1", 8, 16, 21, 2, 1.0, 15),
    ("m1_lam7", 8, 16, 21, 1, 7.0, 16),
]

out = {}
for name, B, D, C, m, lam, seed in CASES:
    inp = make_inputs(B, D, C, seed=seed, w_std=0.05)
    X = inp.X.double().requires_grad_(True)
    W = inp.W.double().requires_grad_(True)
    loss, f = port.asoftmax_graph(X, W, inp.y, m, lam)
    loss.backward()
    out[name + "_shape"] = np.array([B, D, C, m])
    out[name + "_lam"] = np.float64(lam)
    out[name + "_seed"] = np.int64(seed)
    out[name + "_loss"] = np.float64(loss.item())
    out[name + "_logits"] = f.detach().numpy()
    out[name + "_dX"] = X.grad.numpy()
    out[name + "_dW"] = W.grad.numpy()
np.savez_compressed(os.path.join(os.path.dirname(__file__), "asoftmax_small.npz"), **out)
print("wrote", len(CASES), "cases")
