"""Reference-format checkpoint files (TF V2 tensor bundle; saver.py:30-80, train.py:188-213)
written and parsed without TensorFlow: known-answer checksums, table structure, round trips."""
import os
import struct

import numpy as np
import pytest

from tf_face_toolbox_b200 import checkpoint as ck


def test_crc32c_known_answers_and_chunked_path():
    assert ck.crc32c(b"123456789") == 0xE3069283                  # the CRC-32C check value
    assert ck.crc32c(b"") == 0
    assert ck.crc32c(bytes(32)) == 0x8A9136AA                     # RFC 3720 test vector: 32 zero bytes
    assert ck.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43            # RFC 3720: 32 bytes of 0xFF
    rng = np.random.default_rng(0)
    big = rng.integers(0, 256, size=(1 << 18) + 12345, dtype=np.uint8).tobytes()   # chunked + tail path
    assert ck.crc32c(big) == ck._crc_bytes(big) ^ 0xFFFFFFFF


def test_bundle_round_trip_and_layout(tmp_path):
    rng = np.random.default_rng(1)
    tensors = {ck.WEIGHTS: rng.standard_normal((64, 301)).astype(np.float32),
               ck.WEIGHTS + "/Momentum": rng.standard_normal((64, 301)).astype(np.float32),
               "global_step": np.int64(1234)}
    prefix = str(tmp_path / "model.ckpt-1234")
    ck.write_bundle(prefix, tensors)
    idx = open(prefix + ".index", "rb").read()
    assert struct.unpack_from("<Q", idx, len(idx) - 8)[0] == ck.TABLE_MAGIC       # SSTable footer
    assert os.path.getsize(prefix + ".data-00000-of-00001") == 2 * 64 * 301 * 4 + 8
    assert ck.latest_checkpoint(str(tmp_path)) == prefix
    back = ck.read_bundle(prefix)
    assert set(back) == set(tensors)
    for k in tensors:
        np.testing.assert_array_equal(back[k], np.asarray(tensors[k]))
    # header entry: key "" -> BundleHeaderProto{num_shards: 1, version{producer: 1}}
    items = ck._read_table(prefix + ".index")
    assert items[b""] == bytes([0x08, 0x01, 0x1A, 0x02, 0x08, 0x01])
    e = ck._parse_entry(items[ck.WEIGHTS.encode()])
    assert e["dtype"] == ck.DT_FLOAT and e["shape"] == [64, 301] and e["size"] == 64 * 301 * 4


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "m")
    ck.write_bundle(prefix, {ck.WEIGHTS: np.ones((8, 8), np.float32)})
    with open(prefix + ".data-00000-of-00001", "r+b") as fh:
        fh.seek(17)
        fh.write(b"\x7f")
    with pytest.raises(ValueError):
        ck.read_bundle(prefix)


def test_head_save_and_restore_with_tower_prefix(tmp_path):
    """save_head / load_head on a head-like object; a checkpoint whose variable still carries the
    tower prefix `replicated_0/` restores too (restore_op strips it, saver.py:62-72)."""
    import torch
    from tf_face_toolbox_b200 import LambdaState

    class Head:                                     # the attributes save_head / load_head rely on
        def __init__(self, w):
            self.weights, self.lambda_state, self.rank = w, LambdaState(), 0
    w = torch.randn(16, 40)
    h = Head(w.clone())
    h.lambda_state.iteration = 77
    prefix = str(tmp_path / "model.ckpt-77")
    ck.save_head(prefix, h)
    h2 = Head(torch.zeros(16, 40))
    assert ck.load_head(prefix, h2) == 77
    assert torch.equal(h2.weights, w) and h2.lambda_state.iteration == 77
    ck.write_bundle(str(tmp_path / "tower"), {"replicated_0/" + ck.WEIGHTS: w.numpy()})
    h3 = Head(torch.zeros(16, 40))
    ck.load_head(str(tmp_path / "tower"), h3)
    assert torch.equal(h3.weights, w)


# ---------------------------------------------------------------------------------------
# Sharded save / restore (world size 2, gloo): every rank holds a class shard of the weights and
# of the Adam slots; save_head gathers them into the reference's single [D, C] tensors (rank 0
# writes), load_head hands every rank its slice back (saver.py:30-80, data_parallel.py:186-196).
# ---------------------------------------------------------------------------------------
def _ckpt_worker(rank, world, port, prefix, q):
    import os
    import torch
    import torch.distributed as dist
    from tf_face_toolbox_b200 import _lib
    from tf_face_toolbox_b200.sharded import ShardedASoftmaxHead, shard_bounds

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        D, C = 8, 37                                             # 37 classes: ragged shards (19 + 18)
        g = torch.Generator().manual_seed(3)
        W, M, V = (torch.randn(D, C, generator=g) for _ in range(3))
        lo, hi = shard_bounds(C, world, rank)

        class Opt:                                               # what save_head / load_head read of an optimizer
            kind, beta1, beta2 = _lib.OPT_ADAM, 0.5, 0.999

            def __init__(self, s0, s1, step):
                self.state0, self.state1, self.step = s0, s1, step

        head = ShardedASoftmaxHead(D, C, m=4, mode="fp32", device="cpu", weights_full=W, shard_compute=object())
        ck.save_head(prefix, head, Opt(M[:, lo:hi].contiguous(), V[:, lo:hi].contiguous(), 6), global_step=41)
        dist.barrier()
        t = ck.read_bundle(prefix)                               # the file holds the FULL tensors, reference names
        assert sorted(t) == sorted([ck.WEIGHTS, ck.WEIGHTS + "/Adam", ck.WEIGHTS + "/Adam_1", "beta1_power",
                                    "beta2_power", "global_step"])
        assert t[ck.WEIGHTS].shape == (D, C) and np.array_equal(t[ck.WEIGHTS], W.numpy())
        assert np.array_equal(t[ck.WEIGHTS + "/Adam"], M.numpy()) and np.array_equal(t[ck.WEIGHTS + "/Adam_1"], V.numpy())
        fresh = ShardedASoftmaxHead(D, C, m=4, mode="fp32", device="cpu", weights_full=torch.zeros(D, C),
                                    shard_compute=object())
        opt = Opt(None, None, 0)
        step = ck.load_head(prefix, fresh, opt)
        ok = (step == 41 and torch.equal(fresh.weights, W[:, lo:hi]) and torch.equal(opt.state0, M[:, lo:hi])
              and torch.equal(opt.state1, V[:, lo:hi]) and opt.step == 6)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_sharded_head_checkpoint_round_trip_world2(tmp_path):
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    prefix = str(tmp_path / "model.ckpt-41")
    procs = [ctx.Process(target=_ckpt_worker, args=(r, 2, port, prefix, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert outs == [(0, True), (1, True)]
