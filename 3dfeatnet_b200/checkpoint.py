"""Weights in and out: the reference's `initialize_model` (inference.py:183-217, train.py:187-232) for a world without
TensorFlow.

The reference restores `tf.train.Saver` checkpoints (`<prefix>.index` + `<prefix>.data-00000-of-00001`, the "tensor
bundle" format) into variables named by TF scopes.  Here the weights are a flat
{TF variable name: array} dict, and this module
  * reads / writes that dict as `.npz` (the offline-converted form: `np.savez(path, **{n: reader.get_tensor(n) ...})`),
  * reads (and, for round-trip tests, writes) the tensor-bundle format directly with a small pure-Python parser of the
    LevelDB-style table that `.index` is (uncompressed blocks; BundleEntryProto values) -- NOTE: no TF installation and no
    real 3DFeat-Net checkpoint is available offline, so the bundle reader is validated against the published format
    and this module's own writer only,
  * applies the reference's restore rules: scopes in `restore_exclude` keep their initial values (train.sh stage 2
    re-initialises `detection`), `ignore_missing_vars` tolerates absent names, anything else missing is an error.
Name mapping: conv kernels are stored by TF as (1,1,Cin,Cout) and used here as (Cin,Cout); the two EMA shadows TF creates
per batch-norm (`.../moments/Squeeze/ExponentialMovingAverage`, `.../Squeeze_1/...`) are the moving mean / variance.
Optimizer slots (`.../Adam`, `.../Adam_1`, `beta1_power`, ...) and `global_step` are ignored.
"""
import os
import struct

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}  # tensorflow/core/framework/types.proto


# ------------------------------------------------------------------------------------------------ name mapping
def canonical_name(name):
    """TF variable name -> key of this repository's parameter dict, or None for variables that are not model weights."""
    name = name.split(":")[0]
    if name.endswith(("/Adam", "/Adam_1")) or name in ("global_step", "beta1_power", "beta2_power"):
        return None
    if name.endswith("/ExponentialMovingAverage"):
        scope = name.split("/bn/")[0]
        if "Squeeze_1" in name:
            return scope + "/bn/moving_variance"
        if "Squeeze" in name:
            return scope + "/bn/moving_mean"
        return None
    return name


def to_model_arrays(raw):
    """{checkpoint name: array} -> {parameter name: float32 array in this repository's layout}."""
    out = {}
    for name, arr in raw.items():
        key = canonical_name(name)
        if key is None:
            continue
        a = np.asarray(arr, dtype=np.float32)
        if key.endswith("/conv2d/weights") and a.ndim == 4:
            if a.shape[0] != 1 or a.shape[1] != 1:
                raise ValueError("%s: only 1x1 kernels are supported, got %s" % (name, a.shape))
            a = a.reshape(a.shape[2], a.shape[3])
        out[key] = np.ascontiguousarray(a)
    return out


# ------------------------------------------------------------------------------------------------ npz
def save_npz(weights, path, tf_layout=True):
    """Write a {name: tensor/array} dict.  tf_layout=True stores conv kernels as (1,1,Cin,Cout) like a TF checkpoint."""
    arrays = {}
    for k, v in weights.items():
        a = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
        if tf_layout and k.endswith("/conv2d/weights") and a.ndim == 2:
            a = a.reshape(1, 1, *a.shape)
        arrays[k] = a.astype(np.float32)
    np.savez(path, **arrays)


def load_npz(path):
    with np.load(path) as z:
        return to_model_arrays({k: z[k] for k in z.files})


# ------------------------------------------------------------------------------------------------ tensor bundle
def _varint(buf, pos):
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_block(data, offset, size):
    """Entries of one table block: list of (key bytes, value bytes).  `size` excludes the 5-byte trailer."""
    ctype = data[offset + size]
    if ctype != 0:
        raise ValueError("compressed table blocks (type %d) are not supported; TF writes bundle indices uncompressed" % ctype)
    blk = data[offset:offset + size]
    (num_restarts,) = struct.unpack_from("<I", blk, len(blk) - 4)
    limit = len(blk) - 4 - 4 * num_restarts
    pos, key, out = 0, b"", []
    while pos < limit:
        shared, pos = _varint(blk, pos)
        non_shared, pos = _varint(blk, pos)
        vlen, pos = _varint(blk, pos)
        key = key[:shared] + bytes(blk[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(blk[pos:pos + vlen])))
        pos += vlen
    return out


def _parse_entry(buf):
    """BundleEntryProto (tensorflow/core/protobuf/tensor_bundle.proto): dtype=1, shape=2, shard_id=3, offset=4, size=5."""
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "slices": 0}
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            v, pos = _varint(buf, pos)
            if field == 1:
                e["dtype"] = v
            elif field == 3:
                e["shard_id"] = v
            elif field == 4:
                e["offset"] = v
            elif field == 5:
                e["size"] = v
        elif wire == 5:
            pos += 4  # crc32c
        elif wire == 1:
            pos += 8
        elif wire == 2:
            ln, pos = _varint(buf, pos)
            sub = buf[pos:pos + ln]
            pos += ln
            if field == 2:  # TensorShapeProto: repeated Dim dim = 2 { int64 size = 1; }
                sp = 0
                while sp < len(sub):
                    t2, sp = _varint(sub, sp)
                    if t2 & 7 == 2:
                        l2, sp = _varint(sub, sp)
                        dim = sub[sp:sp + l2]
                        sp += l2
                        if t2 >> 3 == 2:
                            dp, size = 0, 0
                            while dp < len(dim):
                                t3, dp = _varint(dim, dp)
                                if t3 & 7 == 0:
                                    v3, dp = _varint(dim, dp)
                                    if t3 >> 3 == 1:
                                        size = v3
                                elif t3 & 7 == 2:
                                    l3, dp = _varint(dim, dp)
                                    dp += l3
                            e["shape"].append(size)
                    elif t2 & 7 == 0:
                        _, sp = _varint(sub, sp)
            elif field == 7:
                e["slices"] += 1
        else:
            raise ValueError("unexpected wire type %d in BundleEntryProto" % wire)
    return e


def read_tf_bundle(prefix):
    """{variable name: ndarray} of a TF tensor-bundle checkpoint `<prefix>.index` / `<prefix>.data-XXXXX-of-YYYYY`."""
    with open(prefix + ".index", "rb") as f:
        data = f.read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != _MAGIC:
        raise ValueError("%s.index is not a TF tensor-bundle index (bad table magic)" % prefix)
    pos = len(data) - 48
    _, pos = _varint(data, pos)  # metaindex handle
    _, pos = _varint(data, pos)
    ioff, pos = _varint(data, pos)
    isize, pos = _varint(data, pos)
    entries = []
    for _, handle in _read_block(data, ioff, isize):
        boff, hp = _varint(handle, 0)
        bsize, hp = _varint(handle, hp)
        entries += _read_block(data, boff, bsize)
    num_shards, shards, out = 1, {}, {}
    for key, val in entries:
        if key == b"":  # BundleHeaderProto: num_shards = 1
            hp = 0
            while hp < len(val):
                tag, hp = _varint(val, hp)
                if tag & 7 == 0:
                    v, hp = _varint(val, hp)
                    if tag >> 3 == 1:
                        num_shards = v
                elif tag & 7 == 2:
                    ln, hp = _varint(val, hp)
                    hp += ln
            continue
        e = _parse_entry(val)
        if e["slices"]:
            raise ValueError("%s: partitioned (sliced) variables are not supported" % key.decode())
        if e["dtype"] not in _DTYPES:
            continue  # strings etc.: not model weights
        if e["shard_id"] not in shards:
            shards[e["shard_id"]] = np.memmap("%s.data-%05d-of-%05d" % (prefix, e["shard_id"], num_shards), dtype=np.uint8, mode="r")
        raw = shards[e["shard_id"]][e["offset"]:e["offset"] + e["size"]]
        out[key.decode()] = np.frombuffer(bytes(raw), dtype=_DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
    return out


def _crc32c_table():
    tbl = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tbl.append(c)
    return tbl


_CRC_TABLE = _crc32c_table()


def _masked_crc32c(buf):
    c = 0xFFFFFFFF
    for b in buf:
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    c ^= 0xFFFFFFFF
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _block(entries):
    body, restarts = bytearray(), []
    for key, val in entries:  # restart at every entry (shared prefix 0): simplest valid encoding
        restarts.append(len(body))
        body += _put_varint(0) + _put_varint(len(key)) + _put_varint(len(val)) + key + val
    if not restarts:
        restarts = [0]
    for r in restarts:
        body += struct.pack("<I", r)
    body += struct.pack("<I", len(restarts))
    return bytes(body)


def write_tf_bundle(prefix, arrays):
    """Write {name: float32 array} as a single-shard tensor bundle (tests / interchange).  Layout follows
    tensorflow/core/util/tensor_bundle: data file = concatenated little-endian tensors, index = one uncompressed table."""
    names = sorted(arrays)
    offset, entries = 0, [(b"", b"\x08\x01" + b"\x1a\x02\x08\x01")]  # header: num_shards=1, version{producer=1}
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for n in names:
            a = np.asarray(arrays[n], dtype=np.float32)
            raw = a.tobytes()  # C order
            f.write(raw)
            shape = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(s) for s in a.shape))
            val = (b"\x08\x01" + b"\x12" + _put_varint(len(shape)) + shape + b"\x20" + _put_varint(offset) + b"\x28" + _put_varint(len(raw))
                   + b"\x35" + struct.pack("<I", _masked_crc32c(raw)))
            entries.append((n.encode(), val))
            offset += len(raw)
    out = bytearray()

    def emit(block):
        off = len(out)
        out.extend(block)
        out.extend(b"\x00" + struct.pack("<I", _masked_crc32c(block + b"\x00")))
        return _put_varint(off) + _put_varint(len(block))

    data_handle = emit(_block(entries))
    meta_handle = emit(_block([]))
    index_handle = emit(_block([(entries[-1][0] + b"\xff", data_handle)]))
    footer = meta_handle + index_handle
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC))
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))


# ------------------------------------------------------------------------------------------------ restore
def latest_checkpoint(checkpoint_dir):
    """tf.train.latest_checkpoint for `--checkpoint <dir>` (inference.py:189-190, train.py:192-193): the prefix named by the
    directory's TF `checkpoint` state file (`model_checkpoint_path: "..."`) when there is one, otherwise the highest-step
    `<name>.ckpt-<step>` found as `.npz` (what train() writes) or as a TF bundle (`.index`); None when there is none."""
    import re

    if not checkpoint_dir or not os.path.isdir(checkpoint_dir):
        return None
    state = os.path.join(checkpoint_dir, "checkpoint")
    if os.path.isfile(state):
        with open(state, "r") as f:
            m = re.search(r'^model_checkpoint_path:\s*"(.*)"\s*$', f.read(), re.M)
        if m:
            p = m.group(1)
            p = p if os.path.isabs(p) else os.path.join(checkpoint_dir, p)
            if os.path.exists(p + ".index") or os.path.exists(p) or os.path.exists(p + ".npz"):
                return p
    best, best_step = None, -1
    for f in os.listdir(checkpoint_dir):
        m = re.match(r"^(.*\.ckpt-(\d+))(\.npz|\.index)$", f)
        if m and int(m.group(2)) > best_step:
            best_step = int(m.group(2))
            best = os.path.join(checkpoint_dir, f if m.group(3) == ".npz" else m.group(1))
    return best


def load_checkpoint(path):
    """`.npz` file, a TF checkpoint prefix (`<path>.index` exists), or a checkpoint DIRECTORY (resolved like
    tf.train.latest_checkpoint, as the reference does when os.path.isdir(checkpoint)) -> {parameter name: float32 array}."""
    if os.path.isdir(path):
        resolved = latest_checkpoint(path)
        if resolved is None:
            raise FileNotFoundError("no checkpoint in directory %s" % path)
        path = resolved
    if os.path.exists(path + ".index"):
        return to_model_arrays(read_tf_bundle(path))
    if os.path.exists(path):
        return load_npz(path)
    if os.path.exists(path + ".npz"):
        return load_npz(path + ".npz")
    raise FileNotFoundError("no checkpoint at %s (.npz or TF bundle prefix)" % path)


def initialize_model(model, checkpoint, ignore_missing_vars=False, restore_exclude=None):
    """The reference's initialize_model (inference.py:183-217): restore every model variable from `checkpoint` except
    those under the scopes in `restore_exclude`; a variable absent from the checkpoint is an error unless
    ignore_missing_vars.  `model` is a Feat3dNet (its `.weights` dict is updated in place).  Returns the restored names."""
    import torch

    if checkpoint is None:
        return []
    ckpt = load_checkpoint(checkpoint)
    import re
    # tf.get_collection(GLOBAL_VARIABLES, scope=e) keeps the variables whose name re.match()es e: a regex PREFIX match
    excluded = lambda name: any(re.match(e, name) for e in (restore_exclude or []))
    restored, missing = [], []
    with torch.no_grad():
        for name, dst in model.weights.items():
            if excluded(name):
                continue
            if name not in ckpt:
                missing.append(name)
                continue
            src = torch.as_tensor(ckpt[name])
            if tuple(src.shape) != tuple(dst.shape):
                raise ValueError("checkpoint variable %s has shape %s, the model expects %s" % (name, tuple(src.shape), tuple(dst.shape)))
            dst.copy_(src.to(dst.device))
            restored.append(name)
    if missing and not ignore_missing_vars:
        raise KeyError("variables missing from checkpoint %s: %s" % (checkpoint, ", ".join(missing[:8]) + (" ..." if len(missing) > 8 else "")))
    model.invalidate()
    return restored
