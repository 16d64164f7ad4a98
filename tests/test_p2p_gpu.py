"""Parity of the NVLink peer-memory transport (asm_p2p_attach / asm_step_p2p) -- the path every
N > 1 benchmark number comes from -- driven on ONE GPU.

`asm_p2p_attach` takes raw device addresses, so G "ranks" can live in one process: G handles,
G streams, G plain device blocks passed to every handle as `peer_bases`.  Each rank's step is
enqueued on its own stream; the kernels synchronise through the release/acquire flag words
exactly as they do across NVLink.  Compared with the float64 oracle: the loss (identical on
every rank), each rank's dX rows, each rank's dW shard -- over 3 consecutive steps (exercises
the step-parity double buffering), eager and replayed from per-rank CUDA graphs, bf16 and fp32,
with and without the fused optimizer.  Replaces, for the head, nccl.all_sum(grads)
(data_parallel.py:175-181) and the tower loop data_parallel.py:203-256.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import asoftmax_ref as ref
from tf_face_toolbox_b200 import FusedOptimizer, _lib
from tf_face_toolbox_b200.head import get_handle
from tf_face_toolbox_b200.sharded import shard_bounds
from tf_face_toolbox_b200.synthetic import make_inputs

pytestmark = pytest.mark.gpu
# Several ranks share ONE GPU here, and a GEMM CTA needs a whole SM.  Two things that are harmless
# with one rank per GPU would starve a peer's kernels: many blocks spinning on a peer, and
# programmatic dependent launch (the next GEMM's CTAs take the SMs early and then wait for the
# norm kernel, which waits for the peer).  Both are switched off for the handles created here.
SHARED_GPU_ENV = {"ASM_P2P_WAITERS": "2", "ASM_PDL": "0"}
LOSS_TOL = {"fp32": 1e-5, "bf16": 2e-3}


def cosine(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(a @ b / np.sqrt((a @ a) * (b @ b)))


class FakeRanks:
    """G ranks of the class-sharded head on one device."""

    def __init__(self, inp, G, mode, tag, m=4):
        self.dev = torch.device("cuda:0")
        self.G, self.mode = G, mode
        B, D = inp.X.shape
        Cn = inp.W.shape[1]
        self.B, self.D, self.Cn, self.b = B, D, Cn, B // G
        assert B % G == 0
        self.handles, self.W, self.streams, self.blocks = [], [], [], []
        self.bounds = [shard_bounds(Cn, G, r) for r in range(G)]
        saved = {k: os.environ.get(k) for k in SHARED_GPU_ENV}
        os.environ.update(SHARED_GPU_ENV)                         # read by asm_create / asm_p2p_attach
        try:
            self._create(inp, G, mode, tag, m)
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v

    def _create(self, inp, G, mode, tag, m):
        B, D, Cn = self.B, self.D, self.Cn
        for r, (lo, hi) in enumerate(self.bounds):
            h = get_handle(self.dev, D, Cn, hi - lo, lo, B, m, mode, r, G, tag=(tag, G, mode, r))
            nbytes = int(h.lib.asm_p2p_bytes(C.byref(h.cfg)))
            assert nbytes > 0
            self.handles.append(h)
            self.blocks.append(torch.zeros(nbytes, dtype=torch.uint8, device=self.dev))
            self.W.append(inp.W[:, lo:hi].contiguous().to(self.dev))
            self.streams.append(torch.cuda.Stream(self.dev))
        ptrs = (C.c_void_p * G)(*[blk.data_ptr() for blk in self.blocks])
        for h in self.handles:
            _lib.check(h.lib.asm_p2p_attach(h.ptr, ptrs), h.ptr)
            _lib.check(h.lib.asm_p2p_set_timeout(h.ptr, 5000), h.ptr)     # a stuck exchange fails in seconds
        self.X = [inp.X[r * self.b:(r + 1) * self.b].contiguous().to(self.dev) for r in range(G)]
        self.y = [inp.y[r * self.b:(r + 1) * self.b].contiguous().to(self.dev) for r in range(G)]
        self.loss = [torch.zeros(1, device=self.dev) for _ in range(G)]
        self.dX = [torch.zeros(self.b, D, device=self.dev) for _ in range(G)]
        self.dW = [torch.zeros_like(w) for w in self.W]
        # First launches are not concurrent: loading a kernel, or growing the local-memory pool for
        # the first kernel with a stack frame, synchronises the whole context -- and a rank that is
        # already spinning on a peer would then wait for ever.  With one rank per process that
        # cannot happen; here every handle first runs the same kernels once with no peer involved
        # (the host-collective form of the step, asm_forward_partial / asm_backward_partial).
        Xall, yall = inp.X.to(self.dev), inp.y.to(self.dev)
        for r, h in enumerate(self.handles):
            stats = torch.zeros(G, 3, B, device=self.dev)
            dXp = torch.zeros(B, D, device=self.dev)
            _lib.check(h.lib.asm_forward_partial(h.ptr, Xall.data_ptr(), B, yall.data_ptr(), 4, self.W[r].data_ptr(),
                                                 5.0, stats[r].data_ptr(), None, None), h.ptr)
            stats[:] = stats[r]
            _lib.check(h.lib.asm_backward_partial(h.ptr, stats.data_ptr(), G, self.loss[r].data_ptr(), dXp.data_ptr(),
                                                  self.dW[r].data_ptr(), None), h.ptr)
            torch.cuda.synchronize()

    def enqueue(self, r, lam, optimizer=None):
        h = self.handles[r]
        st = self.streams[r]
        with torch.cuda.stream(st):
            if optimizer is not None:
                optimizer._arm(h, self.W[r])
            try:
                rc = h.lib.asm_step_p2p(h.ptr, self.X[r].data_ptr(), self.b, self.y[r].data_ptr(), 4,
                                        self.W[r].data_ptr(), lam, self.loss[r].data_ptr(), self.dX[r].data_ptr(),
                                        self.dW[r].data_ptr() if optimizer is None else None,
                                        C.c_void_p(st.cuda_stream))
            finally:
                if optimizer is not None:
                    optimizer._disarm(h)
            _lib.check(rc, h.ptr)

    def step(self, lam, optimizers=None):
        for r in range(self.G):
            self.enqueue(r, lam, None if optimizers is None else optimizers[r])

    def check(self, r_oracle, cos_min=0.9999):
        torch.cuda.synchronize()
        for h in self.handles:
            _lib.check(h.lib.asm_p2p_status(h.ptr, None), h.ptr)         # nobody timed out on a peer
        losses = [float(l) for l in self.loss]
        assert max(losses) == min(losses), losses                  # same bits on every rank
        assert abs(losses[0] - r_oracle.loss) <= LOSS_TOL[self.mode] * abs(r_oracle.loss), (losses[0], r_oracle.loss)
        for r, (lo, hi) in enumerate(self.bounds):
            rows = slice(r * self.b, (r + 1) * self.b)
            assert cosine(self.dX[r].cpu().numpy(), r_oracle.dX[rows]) >= cos_min, ("dX", r)
            if self.dW[r] is not None:
                assert cosine(self.dW[r].cpu().numpy(), r_oracle.dW[:, lo:hi]) >= cos_min, ("dW", r)


# tests/conftest.py sets CUDA_DEVICE_MAX_CONNECTIONS=32 before CUDA starts: with the default of 8
# hardware queues the 2 G streams of G = 8 ranks would alias, and a spinning consumer kernel could
# sit in front of its producer in the same queue.


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
@pytest.mark.parametrize("G", [2, 4, 8])
def test_p2p_step_eager_three_steps(G, mode):
    B, D, Cn = 128, 128, 5000
    inp = make_inputs(B, D, Cn, seed=77)
    ranks = FakeRanks(inp, G, mode, "p2p-eager")
    for it, lam in enumerate((5.0, 0.0, 1000 / 1.12)):            # the buffers alternate by step parity
        ranks.step(lam)
        ranks.check(ref.asoftmax_head(inp.X.numpy(), inp.W.numpy(), inp.y.numpy(), 4, lam))


@pytest.mark.parametrize("G", [2, 8])
def test_p2p_step_cfg3_shard_shapes(G):
    """BASELINE config 3 at full size (C = 85,742, D = 512, batch 512, bf16) through the
    transport the N = 2 / N = 8 benchmark lines use."""
    inp = make_inputs(512, 512, 85742)
    ranks = FakeRanks(inp, G, "bf16", "p2p-cfg3")
    r = ref.asoftmax_head(inp.X.numpy(), inp.W.numpy(), inp.y.numpy(), 4, 5.0)
    for _ in range(2):
        ranks.step(5.0)
    ranks.check(r)


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_p2p_step_graph_replay(mode):
    """One CUDA graph per rank (what ShardedASoftmaxHead.capture builds), replayed on G streams;
    lambda is read from a device scalar, so one captured graph serves every step."""
    G, B, D, Cn = 4, 128, 128, 5000
    inp = make_inputs(B, D, Cn, seed=78)
    ranks = FakeRanks(inp, G, mode, "p2p-graph")
    lam_dev = torch.zeros(1, device=ranks.dev)
    for h in ranks.handles:
        _lib.check(h.lib.asm_set_lambda_device(h.ptr, lam_dev.data_ptr()), h.ptr)
    ranks.step(0.0)                                               # warm-up outside capture (TMA maps)
    torch.cuda.synchronize()
    graphs = []
    for r in range(G):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=ranks.streams[r]):
            ranks.enqueue(r, 0.0)
        graphs.append(g)
    for lam in (5.0, 0.0, 1000 / 1.12, 5.0):
        lam_dev.fill_(lam)
        torch.cuda.synchronize()
        for r in range(G):
            with torch.cuda.stream(ranks.streams[r]):
                graphs[r].replay()
        ranks.check(ref.asoftmax_head(inp.X.numpy(), inp.W.numpy(), inp.y.numpy(), 4, lam))
    for h in ranks.handles:
        h.lib.asm_set_lambda_device(h.ptr, None)


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_p2p_step_with_fused_optimizer(mode):
    """Each rank updates its own class columns inside its dW kernel (data_parallel.py:186-196
    without the G-fold redundancy): first momentum step == W - lr (dW + wd W)."""
    G, B, D, Cn = 4, 128, 128, 5000
    lr, wd = 0.05, 5e-4
    inp = make_inputs(B, D, Cn, seed=79)
    ranks = FakeRanks(inp, G, mode, "p2p-opt")
    r = ref.asoftmax_head(inp.X.numpy(), inp.W.numpy(), inp.y.numpy(), 4, 5.0)
    ranks.step(5.0)
    ranks.check(r)
    W0 = [w.clone() for w in ranks.W]
    opts = [FusedOptimizer("Momentum", lr=lr, weight_decay=wd) for _ in range(G)]
    saved_dW, ranks.dW = ranks.dW, [None] * G
    ranks.step(5.0, optimizers=opts)
    ranks.check(r)                                                # loss and dX unchanged by the fusion
    for g, (lo, hi) in enumerate(ranks.bounds):
        upd = (ranks.W[g] - W0[g]).double().cpu().numpy()
        want = -lr * (r.dW[:, lo:hi] + wd * inp.W[:, lo:hi].double().numpy())
        assert cosine(upd, want) >= 0.9999, g
        np.testing.assert_allclose(upd, want, rtol=0, atol=(2e-3 if mode == "fp32" else 3e-2) * np.abs(want).max())
    ranks.dW = saved_dW


def test_p2p_step_with_bf16_embeddings():
    """The transport publishes and gathers bf16 rows when the embeddings are handed over as bf16."""
    G, B, D, Cn = 4, 128, 128, 5000
    inp = make_inputs(B, D, Cn, seed=80)
    inp = type(inp)(inp.X.to(torch.bfloat16).float(), inp.W, inp.y)      # values exactly representable in bf16
    ranks = FakeRanks(inp, G, "bf16", "p2p-x16")
    ranks.step(5.0)
    r = ref.asoftmax_head(inp.X.numpy(), inp.W.numpy(), inp.y.numpy(), 4, 5.0)
    ranks.check(r)
    want = [t.clone() for t in ranks.dX]
    for g, h in enumerate(ranks.handles):
        _lib.check(h.lib.asm_set_embedding_dtype(h.ptr, 2), h.ptr)
        ranks.X[g] = ranks.X[g].to(torch.bfloat16)
    ranks.step(5.0)
    ranks.check(r)
    for a, b in zip(want, ranks.dX):      # same operand bits, row norms summed in a different order
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-7 * float(a.abs().max()) + 1e-12)
