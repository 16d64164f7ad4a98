"""Quick parity probe of the single-GPU head against the oracle (used for same-box A/B of env knobs):
    python scripts/try_head.py [B D C mode]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import asoftmax_ref as ref
from tf_face_toolbox_b200 import asoftmax_head
from tf_face_toolbox_b200.synthetic import make_inputs

B, D, C = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (512, 512, 10572)
mode = sys.argv[4] if len(sys.argv) > 4 else "bf16"
inp = make_inputs(B, D, C, seed=21)
dev = torch.device("cuda:0")
r = ref.asoftmax_head(inp.X.numpy(), inp.W.numpy(), inp.y.numpy(), 4, 5.0)
loss, _, dX, dW = asoftmax_head(inp.X.to(dev), inp.y.to(dev), C, 4, 5.0, weights=inp.W.to(dev), mode=mode)
torch.cuda.synchronize()
cos = lambda a, b: float((a.ravel() @ b.ravel()) / np.sqrt((a.ravel() @ a.ravel()) * (b.ravel() @ b.ravel())))
dWn = dW.double().cpu().numpy()
print(f"{mode} B={B} D={D} C={C} env={ {k: v for k, v in os.environ.items() if k.startswith('ASM_')} } "
      f"loss_rel={abs(float(loss) - r.loss) / r.loss:.2e} cos_dX={cos(dX.double().cpu().numpy(), r.dX):.6f} "
      f"cos_dW={cos(dWn, r.dW):.6f} max|dW-ref|/max|ref|={np.abs(dWn - r.dW).max() / np.abs(r.dW).max():.2e}")
