"""Multi-GPU parity check of the class-sharded head (run under torchrun on the GPU box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_check.py
Every rank compares its dX rows / dW shard and the loss with the float64 oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import asoftmax_ref as ref                                   # noqa: E402
from tf_face_toolbox_b200 import ShardedASoftmaxHead                     # noqa: E402
from tf_face_toolbox_b200.synthetic import make_inputs                   # noqa: E402


def cos(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(a @ b / np.sqrt((a @ a) * (b @ b)))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    transports = sys.argv[1:] or ["nccl", "nvlink"]
    cases = [(64, 128, 5000, "fp32", 1e-5), (512, 512, 10572, "bf16", 2e-3), (512, 512, 85742, "bf16", 2e-3)]
    for transport in transports:
      for (B, D, C, mode, tol) in cases:
        inp = make_inputs(B, D, C, seed=77)
        head = ShardedASoftmaxHead(D, C, m=4, mode=mode, device=dev, weights_full=inp.W,
                                   transport=transport, batch_global=B)
        b = B // world
        for it in range(3):          # several steps: exercises the parity double-buffering
            loss, dX, dW = head.step(inp.X[rank * b:(rank + 1) * b].to(dev), inp.y[rank * b:(rank + 1) * b].to(dev), 5.0)
        torch.cuda.synchronize()
        mode = f"{transport}/{mode}"
        r = ref.asoftmax_head(inp.X.numpy(), inp.W.numpy(), inp.y.numpy(), 4, 5.0)
        rel = abs(float(loss) - r.loss) / r.loss
        cx = cos(dX.cpu().numpy(), r.dX[rank * b:(rank + 1) * b])
        cw = cos(dW.cpu().numpy(), r.dW[:, head.lo:head.hi])
        good = rel <= tol and cx >= 0.9999 and cw >= 0.9999
        ok &= good
        print(f"rank {rank}/{world} {mode} B={B} C={C}: loss_rel={rel:.2e} cos_dX={cx:.6f} cos_dW={cw:.6f} {'OK' if good else 'FAIL'}", flush=True)
        # one more step with the optimizer fused into the shard's dW kernel (data_parallel.py:
        # 186-196 without the G-fold redundancy): first momentum step => W -= lr (dW + wd W)
        from tf_face_toolbox_b200 import FusedOptimizer
        lr, wd = 0.05, 5e-4
        W_before = head.weights.clone()
        opt = FusedOptimizer("Momentum", lr=lr, weight_decay=wd)
        loss2, dX2, none_dW = head.step(inp.X[rank * b:(rank + 1) * b].to(dev), inp.y[rank * b:(rank + 1) * b].to(dev),
                                        5.0, optimizer=opt)
        torch.cuda.synchronize()
        expect = W_before - lr * (dW + wd * W_before)
        upd = (head.weights - W_before).double().flatten()
        exp_upd = (expect - W_before).double().flatten()
        cu = float(upd @ exp_upd / torch.sqrt((upd @ upd) * (exp_upd @ exp_upd)))
        good2 = none_dW is None and cu >= 0.9999 and abs(float(loss2) - float(loss)) <= 1e-6 * abs(float(loss))
        ok &= good2
        print(f"rank {rank}/{world} {mode} fused optimizer: cos(update)={cu:.6f} {'OK' if good2 else 'FAIL'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
