"""Bring-up probe: two ranks of the NVLink-transport step on one GPU, with host timestamps."""
import os, sys, time
import ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
from oracle import asoftmax_ref as ref
from tf_face_toolbox_b200.synthetic import make_inputs
import test_p2p_gpu as T

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
inp = make_inputs(128, 128, 5000, seed=77)
ranks = T.FakeRanks(inp, G, "bf16", "probe")
t0 = time.time()
for r in range(G):
    ranks.enqueue(r, 5.0)
    print(f"rank {r} enqueued at +{time.time() - t0:.3f}s", flush=True)
torch.cuda.synchronize()
print(f"synchronised at +{time.time() - t0:.3f}s", flush=True)
for h in ranks.handles:
    print("status", h.lib.asm_p2p_status(h.ptr, None), (h.lib.asm_last_error(h.ptr) or b"").decode())
r = ref.asoftmax_head(inp.X.numpy(), inp.W.numpy(), inp.y.numpy(), 4, 5.0)
print("losses", [float(l) for l in ranks.loss], "oracle", r.loss)
