"""Reference-format checkpoint files for the head: the TensorFlow "V2" tensor-bundle layout that
`tf.train.Saver(..., builder=DataParallelSaverBuilder)` writes (train.py:188, :245-248;
saver.py:30-80), produced and parsed without TensorFlow.

What the reference stores for the classifier is ONE tensor, `classifier/fc_classifier/weights`
[D, C] fp32, under the tower-0 variable name with the `replicated_0/` prefix stripped
(saver.py:36-40), plus the optimizer's slot variables (`.../Momentum`, or `.../Adam`, `.../Adam_1`
with the `beta1_power` / `beta2_power` accumulators).  `save_head` / `load_head` move exactly those
between a (class-sharded) head and files `<prefix>.index` + `<prefix>.data-00000-of-00001`.

Layout written here (tensorflow/core/util/tensor_bundle, lib/io/table -- public formats):
  data file   raw little-endian tensor bytes, one after the other
  index file  an SSTable (LevelDB table format): sorted key -> value blocks with restart arrays,
              every block followed by a 1-byte compression tag (0) and a masked CRC-32C; an empty
              metaindex block; an index block; a 48-byte footer ending in the magic
              0xdb4775248b80fb57.  Key "" maps to BundleHeaderProto{num_shards: 1, version{producer: 1}},
              every tensor name to BundleEntryProto{dtype, shape, shard_id: 0, offset, size, crc32c}.
TensorFlow is not installed in this environment, so the files are validated by this module's own
reader (checksums, protobuf fields, round trip); reading them with TensorFlow itself is untested.
Host-side file plumbing only: no part of the GPU hot path.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Optional

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
DT_FLOAT, DT_INT64 = 1, 9
_DTYPES = {DT_FLOAT: np.dtype("<f4"), DT_INT64: np.dtype("<i8")}

# ---------------------------------------------------------------------------- CRC-32C (Castagnoli)
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        poly = 0x82F63B78
        tab = np.zeros(256, dtype=np.uint32)
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ poly if c & 1 else c >> 1
            tab[i] = c
        _CRC_TABLE = tab
    return _CRC_TABLE


def _crc_bytes(data: bytes, crc: int = 0xFFFFFFFF) -> int:
    tab = _crc_table().tolist()
    for b in data:
        crc = tab[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc


def _gf2_times(mat, vec):
    out, i = 0, 0
    while vec:
        if vec & 1:
            out ^= mat[i]
        vec >>= 1
        i += 1
    return out


def _zeros_operator(nbytes: int):
    """32 x 32 GF(2) matrix that advances a CRC-32C register through `nbytes` zero bytes (the
    crc32_combine construction), returned as four 256-entry tables."""
    odd = [0x82F63B78] + [1 << i for i in range(31)]             # one zero BIT
    even = [_gf2_times(odd, odd[i]) for i in range(32)]           # two bits
    odd = [_gf2_times(even, even[i]) for i in range(32)]          # four bits
    result = [1 << i for i in range(32)]                          # identity
    n = nbytes
    while n:                                                      # square: 8, 16, 32 ... bits = 1, 2, 4 ... bytes
        even = [_gf2_times(odd, odd[i]) for i in range(32)]
        if n & 1:
            result = [_gf2_times(even, result[i]) for i in range(32)]
        n >>= 1
        if not n:
            break
        odd = [_gf2_times(even, even[i]) for i in range(32)]
        if n & 1:
            result = [_gf2_times(odd, result[i]) for i in range(32)]
        n >>= 1
    return [[_gf2_times(result, v << (8 * k)) for v in range(256)] for k in range(4)]


def crc32c(data: bytes) -> int:
    """CRC-32C (Castagnoli).  Large buffers are cut into equal chunks whose CRCs advance together,
    one NumPy step per byte position, and are then chained with the zero-run operator
    (crc(A || B) = shift(crc(A), |B|) ^ crc(B)), so a 176 MB weight tensor takes about a second."""
    n = len(data)
    if n < (1 << 16):
        return _crc_bytes(data) ^ 0xFFFFFFFF
    L = 4096
    nchunk = n // L
    arr = np.frombuffer(data, dtype=np.uint8, count=nchunk * L).reshape(nchunk, L)
    tab = _crc_table()
    crcs = np.full(nchunk, 0xFFFFFFFF, dtype=np.uint32)
    for i in range(L):
        crcs = tab[(crcs ^ arr[:, i]) & 0xFF] ^ (crcs >> np.uint32(8))
    crcs ^= np.uint32(0xFFFFFFFF)
    T = _zeros_operator(L)
    total = 0
    for c in crcs.tolist():                                       # chain the chunk CRCs in order
        total = (T[0][total & 0xFF] ^ T[1][(total >> 8) & 0xFF] ^ T[2][(total >> 16) & 0xFF] ^ T[3][total >> 24]) ^ c
    tail = data[nchunk * L:]
    if tail:
        Tt = _zeros_operator(len(tail))
        total = (Tt[0][total & 0xFF] ^ Tt[1][(total >> 8) & 0xFF] ^ Tt[2][(total >> 16) & 0xFF] ^ Tt[3][total >> 24]) \
            ^ (_crc_bytes(tail) ^ 0xFFFFFFFF)
    return total & 0xFFFFFFFF


def masked_crc(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---------------------------------------------------------------------------- varints / protobuf
def _varint(n: int) -> bytes:
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf: bytes, pos: int):
    shift = n = 0
    while True:
        b = buf[pos]
        pos += 1
        n |= (b & 0x7F) << shift
        if not b & 0x80:
            return n, pos
        shift += 7


def _field(num: int, wire: int, payload: bytes) -> bytes:
    return _varint((num << 3) | wire) + payload


def _header_proto() -> bytes:
    version = _field(1, 0, _varint(1))                                  # VersionDef.producer = 1
    return _field(1, 0, _varint(1)) + _field(3, 2, _varint(len(version)) + version)   # num_shards = 1, version


def _entry_proto(dtype: int, shape, offset: int, size: int, crc: int) -> bytes:
    dims = b"".join(_field(2, 2, (lambda d: _varint(len(d)) + d)(_field(1, 0, _varint(int(s))))) for s in shape)
    out = _field(1, 0, _varint(dtype)) + _field(2, 2, _varint(len(dims)) + dims)
    if offset:
        out += _field(4, 0, _varint(offset))
    out += _field(5, 0, _varint(size)) + _field(6, 5, struct.pack("<I", crc))
    return out


def _parse_entry(buf: bytes) -> dict:
    pos, e = 0, dict(dtype=0, shape=[], shard_id=0, offset=0, size=0, crc32c=0)
    while pos < len(buf):
        tag, pos = _read_varint(buf, pos)
        num, wire = tag >> 3, tag & 7
        if wire == 0:
            v, pos = _read_varint(buf, pos)
            name = {1: "dtype", 3: "shard_id", 4: "offset", 5: "size"}.get(num)
            if name:
                e[name] = v
        elif wire == 5:
            (v,) = struct.unpack_from("<I", buf, pos)
            pos += 4
            if num == 6:
                e["crc32c"] = v
        elif wire == 2:
            ln, pos = _read_varint(buf, pos)
            sub = buf[pos:pos + ln]
            pos += ln
            if num == 2:                                                 # TensorShapeProto
                sp = 0
                while sp < len(sub):
                    t2, sp = _read_varint(sub, sp)
                    l2, sp = _read_varint(sub, sp)
                    dim = sub[sp:sp + l2]
                    sp += l2
                    if t2 >> 3 == 2:
                        dp, size = 0, 0
                        while dp < len(dim):
                            t3, dp = _read_varint(dim, dp)
                            if t3 & 7 == 0:
                                v3, dp = _read_varint(dim, dp)
                                if t3 >> 3 == 1:
                                    size = v3
                            else:
                                l3, dp = _read_varint(dim, dp)
                                dp += l3
                        e["shape"].append(size)
        else:
            raise ValueError("unsupported protobuf wire type in BundleEntryProto")
    return e


# ---------------------------------------------------------------------------- SSTable
def _block(entries) -> bytes:
    """One table block: prefix-compressed entries (restart interval 16) + restart array."""
    out, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % 16 == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        out += _varint(shared) + _varint(len(k) - shared) + _varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _with_trailer(block: bytes) -> bytes:
    return block + b"\x00" + struct.pack("<I", masked_crc(block + b"\x00"))


def _write_table(path: str, items: Dict[bytes, bytes]) -> None:
    keys = sorted(items)
    data = _block([(k, items[k]) for k in keys])
    f = bytearray()
    data_off, data_size = 0, len(data)
    f += _with_trailer(data)
    meta = _block([])
    meta_off, meta_size = len(f), len(meta)
    f += _with_trailer(meta)
    index = _block([(keys[-1] + b"\x00", _varint(data_off) + _varint(data_size))])   # key >= last key of the block
    idx_off, idx_size = len(f), len(index)
    f += _with_trailer(index)
    footer = _varint(meta_off) + _varint(meta_size) + _varint(idx_off) + _varint(idx_size)
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    f += footer
    with open(path, "wb") as fh:
        fh.write(bytes(f))


def _read_block(buf: bytes, off: int, size: int):
    block = buf[off:off + size]
    tag, (crc,) = buf[off + size], struct.unpack_from("<I", buf, off + size + 1)
    if tag != 0:
        raise ValueError("compressed table blocks are not supported")
    if masked_crc(block + b"\x00") != crc:
        raise ValueError("table block checksum mismatch")
    (nrest,) = struct.unpack_from("<I", block, len(block) - 4)
    end = len(block) - 4 - 4 * nrest
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _read_varint(block, pos)
        non_shared, pos = _read_varint(block, pos)
        vlen, pos = _read_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def _read_table(path: str) -> Dict[bytes, bytes]:
    buf = open(path, "rb").read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != TABLE_MAGIC:
        raise ValueError(f"{path}: not a tensor-bundle index (bad table magic)")
    foot = buf[-48:]
    pos = 0
    _, pos = _read_varint(foot, pos)
    _, pos = _read_varint(foot, pos)
    idx_off, pos = _read_varint(foot, pos)
    idx_size, pos = _read_varint(foot, pos)
    items = {}
    for _k, handle in _read_block(buf, idx_off, idx_size):
        off, hp = _read_varint(handle, 0)
        size, _ = _read_varint(handle, hp)
        items.update(_read_block(buf, off, size))
    return items


# ---------------------------------------------------------------------------- bundle API
def write_bundle(prefix: str, tensors: Dict[str, np.ndarray]) -> None:
    """Write `<prefix>.index`, `<prefix>.data-00000-of-00001` and the `checkpoint` state file."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)) or ".", exist_ok=True)
    items = {b"": _header_proto()}
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as fh:
        for name in sorted(tensors):
            a = np.asarray(tensors[name])
            dt = DT_INT64 if a.dtype.kind in "iu" else DT_FLOAT
            raw = np.ascontiguousarray(a.astype(_DTYPES[dt])).tobytes()
            fh.write(raw)
            items[name.encode()] = _entry_proto(dt, a.shape, offset, len(raw), masked_crc(raw))
            offset += len(raw)
    _write_table(prefix + ".index", items)
    with open(os.path.join(os.path.dirname(os.path.abspath(prefix)), "checkpoint"), "w") as fh:
        base = os.path.basename(prefix)
        fh.write(f'model_checkpoint_path: "{base}"\nall_model_checkpoint_paths: "{base}"\n')


def read_bundle(prefix: str, verify: bool = True) -> Dict[str, np.ndarray]:
    items = _read_table(prefix + ".index")
    data = np.memmap(prefix + ".data-00000-of-00001", dtype=np.uint8, mode="r")
    out = {}
    for key, val in items.items():
        if key == b"":
            continue
        e = _parse_entry(val)
        raw = bytes(data[e["offset"]:e["offset"] + e["size"]])
        if verify and masked_crc(raw) != e["crc32c"]:
            raise ValueError(f"checksum mismatch for tensor {key.decode()}")
        out[key.decode()] = np.frombuffer(raw, dtype=_DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
    return out


def latest_checkpoint(model_dir: str) -> Optional[str]:
    """tf.train.latest_checkpoint: the prefix named by `<model_dir>/checkpoint` (train.py:207)."""
    path = os.path.join(model_dir, "checkpoint")
    if not os.path.exists(path):
        return None
    for line in open(path):
        if line.startswith("model_checkpoint_path:"):
            return os.path.join(model_dir, line.split(":", 1)[1].strip().strip('"'))
    return None


# ---------------------------------------------------------------------------- the head
WEIGHTS = "classifier/fc_classifier/weights"          # nets/sphere.py:84-90, prefix stripped (saver.py:36-40)


def head_tensors(weights_full, optimizer=None, state0_full=None, state1_full=None, global_step=None):
    """The reference-named tensors of the head: weights [D, C] and, with an optimizer, its slots
    (TF names: Momentum -> '<var>/Momentum'; Adam -> '<var>/Adam', '<var>/Adam_1', 'beta1_power',
    'beta2_power'); `global_step` (train.py:157) drives the lambda schedule."""
    import torch
    t = {WEIGHTS: weights_full.detach().cpu().numpy() if isinstance(weights_full, torch.Tensor) else weights_full}

    def host(x):
        return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)
    if optimizer is not None and state0_full is not None:
        from . import _lib
        if optimizer.kind == _lib.OPT_MOMENTUM:
            t[WEIGHTS + "/Momentum"] = host(state0_full)
        else:
            t[WEIGHTS + "/Adam"] = host(state0_full)
            t[WEIGHTS + "/Adam_1"] = host(state1_full)
            t["beta1_power"] = np.float32(optimizer.beta1 ** optimizer.step)
            t["beta2_power"] = np.float32(optimizer.beta2 ** optimizer.step)
    if global_step is not None:
        t["global_step"] = np.int64(global_step)
    return t


def save_head(prefix: str, head, optimizer=None, global_step=None) -> None:
    """Collective for a ShardedASoftmaxHead (every rank calls it; rank 0 writes): gathers the class
    shards -- and the optimizer slots, which are sharded the same way -- into the single [D, C]
    tensors the reference's saver writes, then writes the bundle."""
    full = head.gather_weights() if hasattr(head, "gather_weights") else head.weights
    s0 = s1 = None
    if optimizer is not None and optimizer.state0 is not None:
        s0 = head.gather_shard(optimizer.state0) if hasattr(head, "gather_shard") else optimizer.state0
        if optimizer.state1 is not None:
            s1 = head.gather_shard(optimizer.state1) if hasattr(head, "gather_shard") else optimizer.state1
    if getattr(head, "rank", 0) == 0:
        step = global_step if global_step is not None else getattr(getattr(head, "lambda_state", None), "iteration", None)
        write_bundle(prefix, head_tensors(full, optimizer, s0, s1, step))


def load_head(prefix: str, head, optimizer=None):
    """Inverse of save_head: every rank reads the bundle and keeps its class slice (weights and
    optimizer slots); accepts `replicated_<k>/`-prefixed names as restore_op does (saver.py:62-72).
    Returns the restored global_step (or None)."""
    import torch
    t = read_bundle(prefix)

    def find(name):
        for k in sorted(t):
            base = "/".join(k.split("/")[1:]) if k.startswith("replicated_") else k
            if base == name:
                return t[k]
        return None
    w = find(WEIGHTS)
    if w is None:
        raise KeyError(f"no '{WEIGHTS}' in {prefix}")
    lo, hi = getattr(head, "lo", 0), getattr(head, "hi", w.shape[1])
    if hasattr(head, "load_weights"):
        head.load_weights(torch.from_numpy(w))
    else:
        head.weights.copy_(torch.from_numpy(w).to(head.weights.device))
    if optimizer is not None:
        from . import _lib
        names = [WEIGHTS + "/Momentum"] if optimizer.kind == _lib.OPT_MOMENTUM else [WEIGHTS + "/Adam", WEIGHTS + "/Adam_1"]
        slots = [find(n) for n in names]
        if all(s is not None for s in slots):
            dev = head.weights.device
            optimizer.state0 = torch.from_numpy(slots[0][:, lo:hi].copy()).to(dev)
            optimizer.state1 = torch.from_numpy(slots[1][:, lo:hi].copy()).to(dev) if len(slots) > 1 else None
            if optimizer.kind == _lib.OPT_ADAM and find("beta1_power") is not None:
                import math
                optimizer.step = int(round(math.log(float(find("beta1_power"))) / math.log(optimizer.beta1)))
    gs = find("global_step")
    if gs is not None and getattr(head, "lambda_state", None) is not None:
        head.lambda_state.iteration = int(gs)
    return None if gs is None else int(gs)
