"""World-size-2/3 CPU (gloo) tests of the class-sharded head's host logic: the collectives,
shard bounds and the combine of per-shard statistics.  The per-shard compute is injected (a
float64 oracle stand-in) because the product's compute needs a B200; what is under test here
is tf_face_toolbox_b200/sharded.py, i.e. the replacement of nccl.all_sum(grads)
(data_parallel.py:175-181)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import asoftmax_ref as ref
from tf_face_toolbox_b200.sharded import ShardedASoftmaxHead, shard_bounds
from tf_face_toolbox_b200.synthetic import make_inputs


class OracleShard:
    """forward_partial / backward_partial with the C ABI's semantics, in float64 NumPy."""

    def __init__(self, lo, hi, m):
        self.lo, self.hi, self.m = lo, hi, m

    def forward_partial(self, X, y, W, lam):
        self.X, self.y, self.W, self.lam = X.double().numpy(), y.numpy(), W.double().numpy(), lam
        mloc, zloc, fy, owned = ref.sharded_partial_stats(self.X, self.W, self.y, self.lo, self.m, lam)
        return torch.from_numpy(np.stack([mloc, zloc, fy])).float()

    def backward_partial(self, stats_all, X, W):
        st = stats_all.double().numpy()
        M, logZ, loss = ref.sharded_combine([(st[g, 0], st[g, 1], st[g, 2]) for g in range(st.shape[0])])
        Xn, Wn, y = self.X, self.W, self.y
        B = Xn.shape[0]
        n = np.sqrt((Xn * Xn).sum(1))
        c = np.sqrt((Wn * Wn).sum(0))
        What = Wn / c
        S = Xn @ What
        lse = M + logZ
        Gp = np.exp(S - lse[:, None]) / B
        yl = y - self.lo
        own = np.nonzero((yl >= 0) & (yl < Wn.shape[1]))[0]
        r = np.zeros(B)
        if len(own):
            s_y = S[own, yl[own]]
            t = np.clip(s_y / n[own], -1, 1)
            psi, dpsi, _ = ref.psi_kform(t, self.m)
            f_y = (self.lam * s_y + n[own] * psi) / (1 + self.lam)
            g_y = (np.exp(f_y - lse[own]) - 1.0) / B
            Gp[own, yl[own]] = g_y * (self.lam + dpsi) / (1 + self.lam)
            r[own] = g_y * (psi - t * dpsi) / ((1 + self.lam) * n[own])
        dXp = Gp @ What.T + r[:, None] * Xn
        q = (Gp * S).sum(0)
        dW = (Xn.T @ Gp - What * q) / c
        return torch.tensor(loss, dtype=torch.float32), torch.from_numpy(dXp).float(), torch.from_numpy(dW).float()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, D, C, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        inp = make_inputs(B, D, C, seed=5, w_std=0.05)
        lo, hi = shard_bounds(C, world, rank)
        head = ShardedASoftmaxHead(D, C, m=4, mode="fp32", device="cpu", weights_full=inp.W,
                                   shard_compute=OracleShard(lo, hi, 4))
        b = B // world
        loss, dX_local, dW_local = head.step(inp.X[rank * b:(rank + 1) * b], inp.y[rank * b:(rank + 1) * b], 5.0)
        Wfull = head.gather_weights()
        # checkpoint round trip in the reference's layout / naming (saver.py:36-40, 62-72)
        sd = head.state_dict()
        assert list(sd.keys()) == ["classifier/fc_classifier/weights"]
        ptr = head.weights.data_ptr()
        head.load_state_dict({"replicated_0/classifier/fc_classifier/weights": sd[head.VARIABLE_NAME] * 2.0,
                              "replicated_1/classifier/fc_classifier/weights": sd[head.VARIABLE_NAME] * 3.0})
        assert head.weights.data_ptr() == ptr                      # in place
        assert torch.equal(head.weights, inp.W[:, lo:hi] * 2.0)    # tower 0 wins
        head.load_state_dict(sd)
        assert torch.equal(head.weights, inp.W[:, lo:hi])
        try:
            head.load_state_dict({"backbone/conv1/weights": torch.zeros(1)})
            raise AssertionError("missing classifier variable must raise")
        except KeyError:
            pass
        q.put((rank, float(loss), dX_local.numpy(), dW_local.numpy(), lo, hi, Wfull.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_step_equals_unsharded_oracle(world):
    B, D, C = 12, 16, 37
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, D, C, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    inp = make_inputs(B, D, C, seed=5, w_std=0.05)
    full = ref.asoftmax_head(inp.X.numpy(), inp.W.numpy(), inp.y.numpy(), 4, 5.0)
    b = B // world
    for rank, loss, dX_local, dW_local, lo, hi, Wfull in outs:
        assert loss == pytest.approx(full.loss, rel=1e-6)
        np.testing.assert_allclose(dX_local, full.dX[rank * b:(rank + 1) * b], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(dW_local, full.dW[:, lo:hi], rtol=1e-4, atol=1e-7)
        np.testing.assert_array_equal(Wfull, inp.W.numpy())     # checkpoint layout [D, C]


def test_shard_bounds_cover_all_classes():
    for C, G in [(85742, 8), (1000000, 8), (10572, 3), (7, 8)]:
        spans = [shard_bounds(C, G, r) for r in range(G)]
        assert spans[0][0] == 0 and spans[-1][1] == C
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0
    assert shard_bounds(85742, 8, 7) == (75026, 85742)
