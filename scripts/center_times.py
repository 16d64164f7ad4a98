"""Center-loss kernels alone at the config-5 shape (bring-up / roofline): mean time per call and
achieved GB/s on the algorithmic bytes 3*B*D*4 + 2*D*4 per touched center."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tf_face_toolbox_b200.center import center_loss
B, D, Cn = 2048, 512, 85742
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
X = torch.randn(B, D, generator=g).to(dev)
y = torch.randint(0, Cn, (B,), generator=g).to(torch.int32).to(dev)
cen = (torch.randn(Cn, D, generator=g) * 0.1).to(dev)
acc = torch.zeros(B, D, device=dev)
for _ in range(5):
    center_loss(X, y, cen, 0.95, 0.008, grad_accum=acc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 200
e0.record()
for _ in range(n):
    center_loss(X, y, cen, 0.95, 0.008, grad_accum=acc)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / n * 1e3
touched = len(torch.unique(y))
by = 3 * B * D * 4 + 2 * D * 4 * touched
print(f"center loss B={B} D={D}: {us:.1f} us per call (sort + apply, incl. host-side allocs), "
      f"{by / 1e6:.1f} MB algorithmic -> {by / us / 1e3:.0f} GB/s")
