"""Per-kernel CUDA-event times of the single-GPU head at an arbitrary shape (bring-up tool):
    python scripts/kernel_times.py B D C [steps]
Prints the mean in-loop duration of every kernel of the step (asm_set_profiling) in microseconds."""
import ctypes as C
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf_face_toolbox_b200 import asoftmax_head
from tf_face_toolbox_b200.head import get_handle
from tf_face_toolbox_b200.synthetic import make_inputs

B, D, Cn = (int(v) for v in sys.argv[1:4])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 30
mode = sys.argv[5] if len(sys.argv) > 5 else "bf16"
dev = torch.device("cuda:0")
inp = make_inputs(B, D, Cn)
X, y, W = inp.X.to(dev), inp.y.to(dev), inp.W.to(dev)
for _ in range(5):
    asoftmax_head(X, y, Cn, 4, 5.0, weights=W, mode=mode)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    asoftmax_head(X, y, Cn, 4, 5.0, weights=W, mode=mode)
e1.record()
torch.cuda.synchronize()
step_us = e0.elapsed_time(e1) / steps * 1e3
h = get_handle(dev, D, Cn, Cn, 0, B, 4, mode)
h.lib.asm_set_profiling(h.ptr, 1)
ms = (C.c_float * 16)()
names = C.create_string_buffer(16 * 32)
acc = {}
for _ in range(steps):
    asoftmax_head(X, y, Cn, 4, 5.0, weights=W, mode=mode)
    n = h.lib.asm_get_profile(h.ptr, 16, ms, names)
    for i in range(n):
        nm = names.raw[i * 32:(i + 1) * 32].split(b"\0")[0].decode()
        acc.setdefault(nm, []).append(ms[i] * 1e3)
h.lib.asm_set_profiling(h.ptr, 0)
env = {k: v for k, v in os.environ.items() if k.startswith("ASM_") and k != "ASM_B200_LIB"}
units = ((Cn + 255) // 256) * ((B + 255) // 256)
print(f"B={B} D={D} C={Cn} {mode} {env} step={step_us:.1f}us rounds={units / 74:.2f} " +
      " ".join(f"{k[:9]}={sum(v) / len(v):.1f}" for k, v in acc.items()))
