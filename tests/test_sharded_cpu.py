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


# ---------------------------------------------------------------------------------------
# Train-loop glue: data-parallel backbone (DDP) + class-parallel head.
# examples/train_sphereface20.py::head_step must reproduce single-process training on the
# full batch -- in particular the `* world` that turns DDP's gradient mean into the sum the
# global-batch mean loss needs (the mirror of data_parallel.py:37 + :179).
# ---------------------------------------------------------------------------------------
def _load_example():
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples",
                        "train_sphereface20.py")
    spec = importlib.util.spec_from_file_location("train_sphereface20", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _tiny_net():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(24, 32), torch.nn.Tanh(), torch.nn.Linear(32, 16))


def _glue_data(B, C):
    g = torch.Generator().manual_seed(11)
    return torch.randn(B, 2, 3, 4, generator=g), torch.randint(0, C, (B,), generator=g, dtype=torch.int32)


def _glue_worker(rank, world, port, B, C, steps, lr, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ex = _load_example()
        net = torch.nn.parallel.DistributedDataParallel(_tiny_net())
        opt = torch.optim.SGD(net.parameters(), lr=lr)
        W0 = make_inputs(B, 16, C, seed=5, w_std=0.05).W
        lo, hi = shard_bounds(C, world, rank)
        head = ShardedASoftmaxHead(16, C, m=4, mode="fp32", device="cpu", weights_full=W0,
                                   shard_compute=OracleShard(lo, hi, 4))

        def head_call(feats, labels):
            loss, dX, dW = head.step(feats, labels, 5.0)
            head.weights.sub_(lr * dW)                       # plain SGD on the shard
            return loss, dX
        images, labels = _glue_data(B, C)
        b = B // world
        for _ in range(steps):
            ex.head_step(net, images[rank * b:(rank + 1) * b], labels[rank * b:(rank + 1) * b], head_call, opt,
                         world=world, clip=0, autocast=False)
        Wfull = head.gather_weights()
        q.put((rank, [p.detach().numpy().copy() for p in net.module.parameters()], Wfull.numpy()))
    finally:
        dist.destroy_process_group()


def test_ddp_backbone_plus_sharded_head_equals_single_process_training():
    B, C, steps, lr, world = 12, 37, 3, 0.5, 2
    # single process, full batch, oracle head
    ex = _load_example()
    net = _tiny_net()
    opt = torch.optim.SGD(net.parameters(), lr=lr)
    W = make_inputs(B, 16, C, seed=5, w_std=0.05).W.clone()
    images, labels = _glue_data(B, C)

    def head_call(feats, y):
        r = ref.asoftmax_head(feats.numpy(), W.numpy(), y.numpy(), 4, 5.0)
        W.sub_(lr * torch.from_numpy(r.dW).float())
        return torch.tensor(r.loss), torch.from_numpy(r.dX).float()
    for _ in range(steps):
        ex.head_step(net, images, labels, head_call, opt, world=1, clip=0, autocast=False)
    want = [p.detach().numpy() for p in net.parameters()]

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_glue_worker, args=(r, world, port, B, C, steps, lr, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, params, Wfull in outs:
        for got, exp in zip(params, want):
            np.testing.assert_allclose(got, exp, rtol=2e-4, atol=1e-6)
        np.testing.assert_allclose(Wfull, W.numpy(), rtol=2e-4, atol=1e-7)
