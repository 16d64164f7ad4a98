"""Bring-up probe (not a pytest): reports per-contraction parity of the tcgen05 path for a
set of MN-major shared-memory descriptor parameters.  Each candidate runs in a subprocess
because the parameters are read (ASM_UMMA_MN_*) when the handle is created.
    python tests/probe_umma.py            # sweep
    python tests/probe_umma.py one        # current env only
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import numpy as np
    import torch
    from oracle import asoftmax_ref as ref
    from tf_face_toolbox_b200 import asoftmax_head
    from tf_face_toolbox_b200.synthetic import make_inputs
    dev = torch.device("cuda:0")
    inp = make_inputs(256, 512, 4000, seed=31)
    r = ref.asoftmax_head(inp.X.numpy(), inp.W.numpy(), inp.y.numpy(), 4, 5.0)

    def cos(a, b):
        a = a.astype(np.float64).ravel(); b = b.astype(np.float64).ravel()
        return float(a @ b / np.sqrt((a @ a) * (b @ b) + 1e-300))
    loss, _, dX, dW = asoftmax_head(inp.X.to(dev), inp.y.to(dev), 4000, 4, 5.0, weights=inp.W.to(dev), mode="bf16")
    torch.cuda.synchronize()
    print("RESULT loss_rel=%.3e cos_dX=%.6f cos_dW=%.6f" % (
        abs(float(loss) - r.loss) / r.loss, cos(dX.cpu().numpy(), r.dX), cos(dW.cpu().numpy(), r.dW)), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
        sys.exit(0)
    cands = [(8192, 1024, 2048), (1024, 8192, 2048), (8192, 1024, 256), (1024, 8192, 256),
             (128, 1024, 2048), (8192, 128, 2048)]
    for lbo, sbo, ks in cands:
        env = dict(os.environ, ASM_UMMA_MN_LBO=str(lbo), ASM_UMMA_MN_SBO=str(sbo), ASM_UMMA_MN_KSTEP=str(ks))
        try:
            out = subprocess.run([sys.executable, __file__, "one"], env=env, capture_output=True, text=True, timeout=120)
            lines = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
            print(f"lbo={lbo} sbo={sbo} kstep={ks}: {lines[0] if lines else 'FAILED rc=%d %s' % (out.returncode, out.stderr[-300:])}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"lbo={lbo} sbo={sbo} kstep={ks}: TIMEOUT", flush=True)
