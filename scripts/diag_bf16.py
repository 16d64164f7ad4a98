# which of the two calls of test_bf16_embeddings_at_the_boundary has the wrong dW, and where
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf_face_toolbox_b200.head import asoftmax_head
from tf_face_toolbox_b200.synthetic import make_inputs
from oracle import asoftmax_ref as ref
dev = torch.device("cuda:0")
inp = make_inputs(300, 192, 2002, seed=61)
X16 = inp.X.to(dev).to(torch.bfloat16); X32 = X16.float()
y, W = inp.y.to(dev), inp.W.to(dev)
r = ref.asoftmax_head(X32.cpu().numpy(), inp.W.numpy(), inp.y.numpy(), 4, 5.0)
def rep(tag, dW):
    d = dW.cpu().numpy() - r.dW
    e = np.abs(d).max(axis=0) / np.abs(r.dW).max()
    bad = np.nonzero(e > 1e-2)[0]
    rows = np.nonzero(np.abs(d).max(axis=1) / np.abs(r.dW).max() > 1e-2)[0]
    print(tag, "max rel err", float(e.max()), "bad cols", len(bad), bad[:12], bad[-4:] if len(bad) else "", "bad rows", len(rows), rows[:12])
for tag, X in (("fp32 #1", X32), ("fp32 #2", X32), ("bf16 #1", X16), ("bf16 #2", X16), ("fp32 #3", X32)):
    _, _, dX, dW = asoftmax_head(X, y, 2002, 4, 5.0, weights=W, mode="bf16")
    torch.cuda.synchronize()
    rep(tag, dW)
