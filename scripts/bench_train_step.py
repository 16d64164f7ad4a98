"""Head TRAINING step (fwd + bwd + classifier update) at a BASELINE config, on one B200.

Not the headline metric (bench.py measures fwd+bwd without the optimizer); this
measures SURVEY 8f row 1: the optimizer of data_parallel.py:186-196 fused into the dW epilogue
(`asoftmax_head(..., optimizer=FusedOptimizer(...))`) against the same step with the update done
by separate torch ops on the returned dW.  Prints one JSON line.

    python scripts/bench_train_step.py [--workload cfg3] [--steps 50]
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tf_face_toolbox_b200 import FusedOptimizer, asoftmax_head  # noqa: E402
from tf_face_toolbox_b200.head import get_handle  # noqa: E402
from tf_face_toolbox_b200.synthetic import CONFIGS, make_inputs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    cfg = CONFIGS[args.workload]
    B, D, Cn, mode = cfg["B"], cfg["D"], cfg["C"], cfg["mode"]
    dev = torch.device("cuda:0")
    inp = make_inputs(B, D, Cn)
    X, y = inp.X.to(dev), inp.y.to(dev)
    lr, mom, wd = 0.01, 0.9, 5e-4

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    out = {"workload": args.workload, "B": B, "D": D, "C": Cn, "mode": mode, "steps": args.steps}

    # (a) head only (what bench.py times)
    W = inp.W.to(dev).clone()
    out["head_only_ms"] = timed(lambda: asoftmax_head(X, y, Cn, 4, 5.0, weights=W, mode=mode))

    # (b) unfused: dW to HBM, TF-momentum update with torch ops
    acc = torch.zeros_like(W)

    def unfused():
        _, _, _, dW = asoftmax_head(X, y, Cn, 4, 5.0, weights=W, mode=mode)
        acc.mul_(mom).add_(dW).add_(W, alpha=wd)
        W.add_(acc, alpha=-lr)
    out["unfused_momentum_ms"] = timed(unfused)

    # (c) fused in the dW epilogue
    for kind in ("Momentum", "Adam"):
        Wf = inp.W.to(dev).clone()
        opt = FusedOptimizer(kind, lr=lr if kind == "Momentum" else 1e-4, weight_decay=wd)
        out[f"fused_{kind.lower()}_ms"] = timed(
            lambda: asoftmax_head(X, y, Cn, 4, 5.0, weights=Wf, mode=mode, optimizer=opt))
        if kind == "Momentum":
            # per-kernel split of the fused step (library events)
            h = get_handle(dev, D, Cn, Cn, 0, B, 4, mode)
            h.lib.asm_set_profiling(h.ptr, 1)
            ms_buf = (C.c_float * 16)()
            names = C.create_string_buffer(16 * 32)
            acc_k = {}
            for _ in range(10):
                asoftmax_head(X, y, Cn, 4, 5.0, weights=Wf, mode=mode, optimizer=opt)
                n = h.lib.asm_get_profile(h.ptr, 16, ms_buf, names)
                for i in range(max(n, 0)):
                    nm = names.raw[i * 32:(i + 1) * 32].split(b"\0")[0].decode()
                    acc_k.setdefault(nm, []).append(ms_buf[i])
            h.lib.asm_set_profiling(h.ptr, 0)
            out["fused_momentum_kernels_us"] = {k: round(1e3 * sum(v) / len(v), 1) for k, v in acc_k.items()}
            Cp = (Cn + 255) // 256 * 256
            by = 2.0 * B * Cp + 4 * 4.0 * D * Cn      # G'' read + W, state read + write
            dwms = out["fused_momentum_kernels_us"].get("dw_gemm", 0) * 1e-3
            if dwms > 0:
                out["fused_dw_kernel_GBps"] = by / (dwms * 1e-3) / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
