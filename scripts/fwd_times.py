# forward half alone (prep + forward GEMM + combine) at config 3, CUDA events over many calls
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf_face_toolbox_b200.head import asoftmax_head
from tf_face_toolbox_b200.synthetic import make_inputs
dev = torch.device("cuda:0")
inp = make_inputs(512, 512, 85742)
X, y, W = inp.X.to(dev), inp.y.to(dev), inp.W.to(dev)
def run(grads, n=40):
    for _ in range(5): asoftmax_head(X, y, 85742, 4, 5.0, weights=W, mode="bf16", compute_grads=grads, check_labels=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = asoftmax_head(X, y, 85742, 4, 5.0, weights=W, mode="bf16", compute_grads=grads, check_labels=False)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000, float(out[0])
f, lf = run(False); s, ls = run(True)
print("env", {k: v for k, v in os.environ.items() if k.startswith("ASM_")}, "forward %.1f us  step %.1f us  loss %.6f %.6f" % (f, s, lf, ls))
