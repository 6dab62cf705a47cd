"""Pure-Python reader / writer of TensorFlow "V2" checkpoints (tensor bundles), without TensorFlow.

The reference restores its voices with `tf.train.Saver().restore(sess, tf.train.latest_checkpoint(dir))`
(`synthesize.py:302-316`, `train.py:196-207`): a `<prefix>.index` file (a LevelDB-format sorted table that maps
variable names to BundleEntryProto records) plus `<prefix>.data-00000-of-00001` (raw little-endian tensor bytes).
The variable names and kernel layouts of `VariableStore` are the reference's, so a checkpoint written by the
reference loads by name, and a checkpoint written here can be restored by the reference (synthesize.py's restore by
scope as well as train.py's Supervisor, which wants every global variable: Adam slots, beta powers, global_step).

Formats restated from their public definitions (nothing of this can be executed against TensorFlow in this image,
so it is pinned by round trips and by the byte-level known answers in tests/test_host.py only):
  * table: leveldb `table_format.md` — data blocks of prefix-compressed entries with restart points, each followed by
    a 1-byte compression type and a masked CRC32C; an index block; a 48-byte footer with magic 0xdb4775248b80fb57.
    TensorFlow writes bundle indexes uncompressed (tensor_bundle.cc: options.compression = kNoCompression).
  * records: `tensor_bundle.proto` — key "" -> BundleHeaderProto{num_shards=1, endianness=2, version=3};
    name -> BundleEntryProto{dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6}.
"""
import os
import struct
from collections import OrderedDict

import numpy as np

_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}          # DT_FLOAT, DT_DOUBLE, DT_INT32, DT_INT64
_DTYPE_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


# ------------------------------------------------------------------------------------------------ CRC32C (Castagnoli)
def _make_tables():
    poly = 0x82F63B78
    t0 = []
    for n in range(256):
        c = n
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        t0.append(c)
    tables = [t0]
    for k in range(1, 8):
        prev = tables[k - 1]
        tables.append([(prev[n] >> 8) ^ t0[prev[n] & 0xFF] for n in range(256)])
    return tables


_T = _make_tables()


_native_crc = None


def _native():
    """`oph_crc32c` of the C-ABI library when it has been built (False otherwise: the pure-Python loop below is the same
    function, ~100x slower -- noticeable only on the 100-300 MB data files of full checkpoints)."""
    global _native_crc
    if _native_crc is None:
        try:
            from . import _lib
            _native_crc = _lib.load().oph_crc32c
        except Exception:  # noqa: BLE001 -- library not built: checkpoints still work
            _native_crc = False
    return _native_crc


def crc32c(data, crc=0):
    """CRC-32C of a bytes-like object (slicing-by-8)."""
    if len(data) >= 4096 and _native():
        mv = memoryview(data).cast("B")
        if mv.readonly:
            return int(_native_crc(bytes(mv) if not isinstance(data, bytes) else data, len(mv), crc))
        import ctypes
        return int(_native_crc((ctypes.c_char * len(mv)).from_buffer(mv), len(mv), crc))
    return _crc32c_py(data, crc)


def _crc32c_py(data, crc=0):
    t0, t1, t2, t3, t4, t5, t6, t7 = _T
    c = crc ^ 0xFFFFFFFF
    mv = memoryview(data).cast("B")
    n8 = len(mv) // 8 * 8
    if n8:
        words = struct.unpack("<%dQ" % (n8 // 8), mv[:n8])
        for w in words:
            w ^= c
            c = (t7[w & 0xFF] ^ t6[(w >> 8) & 0xFF] ^ t5[(w >> 16) & 0xFF] ^ t4[(w >> 24) & 0xFF] ^
                 t3[(w >> 32) & 0xFF] ^ t2[(w >> 40) & 0xFF] ^ t1[(w >> 48) & 0xFF] ^ t0[(w >> 56) & 0xFF])
    for b in mv[n8:]:
        c = t0[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data):
    """LevelDB / TensorFlow store CRCs rotated and offset so that CRCs of CRCs stay well distributed."""
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ varints / protobuf
def _put_varint(out, v):
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)


def _get_varint(buf, pos):
    shift = v = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if b < 0x80:
            return v, pos
        shift += 7


def _parse_fields(buf):
    """Flat protobuf decode: [(field number, wire type, value)]; value is int or bytes."""
    pos, out = 0, []
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.append((f, wt, v))
    return out


def _encode_entry(dtype_code, shape, offset, size, crc):
    dims = bytearray()
    for d in shape:                                    # TensorShapeProto.dim (field 2) { size = 1 }
        dim = bytearray([0x08])
        _put_varint(dim, int(d))
        dims += bytes([0x12])
        _put_varint(dims, len(dim))
        dims += dim
    e = bytearray([0x08])
    _put_varint(e, dtype_code)
    e += bytes([0x12])
    _put_varint(e, len(dims))
    e += dims
    if offset:
        e += bytes([0x20])
        _put_varint(e, offset)
    e += bytes([0x28])
    _put_varint(e, size)
    e += bytes([0x35]) + struct.pack("<I", crc)
    return bytes(e)


def _decode_entry(buf):
    ent = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
    for f, _wt, v in _parse_fields(buf):
        if f == 1:
            ent["dtype"] = v
        elif f == 2:
            dims = []
            for f2, _w2, v2 in _parse_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _w3, v3 in _parse_fields(v2):
                        if f3 == 1:
                            size = v3
                    dims.append(size)
            ent["shape"] = tuple(dims)
        elif f == 3:
            ent["shard_id"] = v
        elif f == 4:
            ent["offset"] = v
        elif f == 5:
            ent["size"] = v
        elif f == 6:
            ent["crc32c"] = v
        elif f == 7:
            ent["sliced"] = True
    return ent


# ------------------------------------------------------------------------------------------------ sorted table
def _read_block(f, offset, size, verify=True):
    f.seek(offset)
    raw = f.read(size + 5)
    contents, ctype, crc = raw[:size], raw[size], struct.unpack_from("<I", raw, size + 1)[0]
    if verify and masked_crc32c(raw[:size + 1]) != crc:
        raise IOError("checkpoint index: block checksum mismatch at offset %d" % offset)
    if ctype != 0:
        raise IOError("checkpoint index: compressed blocks (type %d) are not supported" % ctype)
    return contents


def _block_entries(block):
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def _build_block(items, restart_interval):
    out, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            m = min(len(k), len(last))
            while shared < m and k[shared] == last[shared]:
                shared += 1
        _put_varint(out, shared)
        _put_varint(out, len(k) - shared)
        _put_varint(out, len(v))
        out += k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _write_block(f, contents):
    offset = f.tell()
    f.write(contents)
    f.write(b"\x00")
    f.write(struct.pack("<I", masked_crc32c(contents + b"\x00")))
    h = bytearray()
    _put_varint(h, offset)
    _put_varint(h, len(contents))
    return bytes(h)


# ------------------------------------------------------------------------------------------------ public API
def list_variables(prefix, verify=True, with_header=False):
    """OrderedDict name -> entry dict (dtype code, shape, shard_id, offset, size, crc32c) of `<prefix>.index`
    (with_header: also the number of data shards from the bundle header)."""
    path = prefix + ".index"
    out = OrderedDict()
    num_shards = 1
    with open(path, "rb") as f:
        f.seek(0, os.SEEK_END)
        total = f.tell()
        if total < 48:
            raise IOError("%s is too short to be a checkpoint index" % path)
        f.seek(total - 48)
        footer = f.read(48)
        if struct.unpack_from("<Q", footer, 40)[0] != _MAGIC:
            raise IOError("%s: bad table magic (not a TensorFlow V2 checkpoint index)" % path)
        _mo, p = _get_varint(footer, 0)
        _ms, p = _get_varint(footer, p)
        io_, p = _get_varint(footer, p)
        is_, p = _get_varint(footer, p)
        for _sep, handle in _block_entries(_read_block(f, io_, is_, verify)):
            bo, q = _get_varint(handle, 0)
            bs, q = _get_varint(handle, q)
            for key, val in _block_entries(_read_block(f, bo, bs, verify)):
                if key == b"":
                    hdr = {fn: v for fn, _wt, v in _parse_fields(val)}
                    if hdr.get(2, 0) != 0:
                        raise IOError("big-endian checkpoints are not supported")
                    num_shards = hdr.get(1, 1)
                    continue
                out[key.decode()] = _decode_entry(val)
    return (out, num_shards) if with_header else out


def read_checkpoint(prefix, names=None, verify_data=False):
    """{name: numpy array} for every (or the requested) variable of the checkpoint `<prefix>.index/.data-*`."""
    entries, num = list_variables(prefix, with_header=True)
    out = OrderedDict()
    files = {}
    try:
        for name, e in entries.items():
            if names is not None and name not in names:
                continue
            if e["sliced"]:
                raise IOError("%s: partitioned variables are not supported" % name)
            if e["dtype"] not in _DTYPES:
                continue                                # strings etc. (not part of the model)
            sid = e["shard_id"]
            if sid not in files:
                files[sid] = open("%s.data-%05d-of-%05d" % (prefix, sid, num), "rb")
            fh = files[sid]
            fh.seek(e["offset"])
            raw = fh.read(e["size"])
            if len(raw) != e["size"]:
                raise IOError("%s: data file is truncated" % name)
            if verify_data and e["crc32c"] is not None and masked_crc32c(raw) != e["crc32c"]:
                raise IOError("%s: tensor checksum mismatch" % name)
            out[name] = np.frombuffer(raw, dtype=np.dtype(_DTYPES[e["dtype"]]).newbyteorder("<")).reshape(e["shape"]).copy()
    finally:
        for fh in files.values():
            fh.close()
    return out


def write_checkpoint(prefix, tensors, block_size=4096):
    """Write {name: array} as a one-shard V2 checkpoint that `tf.train.Saver().restore` accepts."""
    names = sorted(tensors, key=lambda n: n.encode())
    items, offset = [], 0
    # both files are written under temporary names and moved into place at the end: an interrupted save never leaves a
    # half-written checkpoint under the name the `checkpoint` state file points to
    final_data, final_index = prefix + ".data-00000-of-00001", prefix + ".index"
    tmp_data, tmp_index = final_data + ".tmp%d" % os.getpid(), final_index + ".tmp%d" % os.getpid()
    with open(tmp_data, "wb") as df:
        for n in names:
            a = np.asarray(tensors[n], order="C")               # (ascontiguousarray would turn scalars into shape [1])
            code = _DTYPE_CODES.get(a.dtype)
            if code is None:
                raise TypeError("%s: dtype %s is not supported" % (n, a.dtype))
            raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
            df.write(raw)
            items.append((n.encode(), _encode_entry(code, a.shape, offset, len(raw), masked_crc32c(raw))))
            offset += len(raw)
    header = bytes([0x08, 0x01, 0x1A, 0x02, 0x08, 0x01])       # num_shards = 1, (little endian), version { producer: 1 }
    items = [(b"", header)] + items
    with open(tmp_index, "wb") as f:
        index, cur, cur_bytes = [], [], 0
        for k, v in items:
            cur.append((k, v))
            cur_bytes += len(k) + len(v) + 3
            if cur_bytes >= block_size:
                index.append((cur[-1][0], _write_block(f, _build_block(cur, 16))))
                cur, cur_bytes = [], 0
        if cur:
            index.append((cur[-1][0], _write_block(f, _build_block(cur, 16))))
        meta = _write_block(f, _build_block([], 1))
        idx = _write_block(f, _build_block(index, 1))
        footer = bytearray(meta + idx)
        footer += b"\x00" * (40 - len(footer))
        footer += struct.pack("<Q", _MAGIC)
        f.write(bytes(footer))
    os.replace(tmp_data, final_data)
    os.replace(tmp_index, final_index)


def latest_checkpoint(directory):
    """`tf.train.latest_checkpoint`: the prefix named by the `checkpoint` state file (first line), else None."""
    state = os.path.join(directory, "checkpoint")
    if not os.path.exists(state):
        return None
    with open(state) as f:
        for line in f:
            if line.startswith("model_checkpoint_path:"):
                p = line.split(":", 1)[1].strip().strip('"')
                return p if os.path.isabs(p) else os.path.join(directory, p)
    return None


def restore(store, prefix, strict=True, with_optimizer=True):
    """Load a reference checkpoint into a VariableStore by variable name (synthesize.py:302-316).

    Model variables must all be present when strict.  With with_optimizer the Adam slots (`<var>/Adam`, `<var>/Adam_1`,
    tf.train.AdamOptimizer's first / second moments) and `global_step` are restored too when the store has optimiser
    state (train.py:196-207, `restart_from_savepath`).  Returns the list of checkpoint names that were not used."""
    import torch
    ck = read_checkpoint(prefix)
    store.finalize()
    used = set()
    values = {}
    for n in store.vars:
        if n in ck:
            values[n] = ck[n]
            used.add(n)
        elif strict:
            raise KeyError("variable %s is missing from checkpoint %s" % (n, prefix))
    store.load_state_dict(values, strict=True)
    if with_optimizer and store.m_flat is not None:
        for n in store.vars:
            o, cnt = store.offsets[n], int(np.prod(store.specs[n][0]))
            for slot, flat in (("Adam", store.m_flat), ("Adam_1", store.v_flat)):
                key = "%s/%s" % (n, slot)
                if key in ck:
                    flat[o:o + cnt].copy_(torch.as_tensor(np.asarray(ck[key], np.float32).reshape(-1)).to(flat.device))
                    used.add(key)
        if "global_step" in ck:
            store.global_step.fill_(int(np.asarray(ck["global_step"]).reshape(-1)[0]))
            used.add("global_step")
        # Adam's beta powers are functions of global_step here (adam_prepare_kernel recomputes them every step)
        used.update(n for n in ("beta1_power", "beta2_power") if n in ck)
    return [n for n in ck if n not in used]


def save(store, prefix, with_optimizer=True):
    """Write a VariableStore as a reference-compatible checkpoint (names, layouts, Adam slots, global_step)."""
    tensors = OrderedDict(store.state_dict())
    if with_optimizer and store.m_flat is not None:
        for n in store.vars:
            o, cnt = store.offsets[n], int(np.prod(store.specs[n][0]))
            shape = store.specs[n][0]
            tensors["%s/Adam" % n] = store.m_flat[o:o + cnt].detach().cpu().numpy().reshape(shape).copy()
            tensors["%s/Adam_1" % n] = store.v_flat[o:o + cnt].detach().cpu().numpy().reshape(shape).copy()
        step = int(store.global_step.item())
        tensors["global_step"] = np.asarray(step, dtype=np.int32)
        # tf.train.AdamOptimizer's non-slot variables: train.py resumes through tf.train.Supervisor, which restores ALL
        # global variables and fails with NotFound without them.  After t updates they hold beta^(t+1).
        hp = getattr(store, "hp", None)
        b1, b2 = (getattr(hp, "beta1", 0.9), getattr(hp, "beta2", 0.999)) if hp is not None else (0.9, 0.999)
        tensors["beta1_power"] = np.asarray(b1 ** (step + 1), dtype=np.float32)
        tensors["beta2_power"] = np.asarray(b2 ** (step + 1), dtype=np.float32)
    write_checkpoint(prefix, tensors)
    with open(os.path.join(os.path.dirname(os.path.abspath(prefix)), "checkpoint"), "w") as f:
        base = os.path.basename(prefix)
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))


def _read_state(directory):
    """(latest, all paths) of the `checkpoint` state file (text-format CheckpointState proto), paths as written."""
    state = os.path.join(directory, "checkpoint")
    latest, paths = None, []
    if os.path.exists(state):
        with open(state) as f:
            for line in f:
                key, _, val = line.partition(":")
                val = val.strip().strip('"')
                if key.strip() == "model_checkpoint_path":
                    latest = val
                elif key.strip() == "all_model_checkpoint_paths":
                    paths.append(val)
    return latest, paths


class Saver(object):
    """The part of `tf.train.Saver` that train.py relies on through `tf.train.Supervisor` (train.py:190, 296-297):
    `save(store, prefix)` writes `<prefix>.index` / `<prefix>.data-00000-of-00001`, keeps the `max_to_keep` (5) most
    recent checkpoints of the directory, deletes older ones and maintains the `checkpoint` state file that
    `latest_checkpoint` reads.  Checkpoints listed by an existing state file are adopted, so a resumed run keeps
    rotating the files of the previous run."""

    def __init__(self, max_to_keep=5):
        self.max_to_keep = max_to_keep
        self._dirs = {}

    def _known(self, directory):
        if directory not in self._dirs:
            _, paths = _read_state(directory)
            self._dirs[directory] = [p for p in paths
                                     if os.path.exists((p if os.path.isabs(p) else os.path.join(directory, p)) + ".index")]
        return self._dirs[directory]

    def save(self, store, prefix, with_optimizer=True):
        directory = os.path.dirname(os.path.abspath(prefix))
        os.makedirs(directory, exist_ok=True)
        known = self._known(directory)
        save(store, prefix, with_optimizer=with_optimizer)          # (also writes a one-entry state file)
        base = os.path.basename(prefix)
        if base in known:
            known.remove(base)
        known.append(base)
        while self.max_to_keep and len(known) > self.max_to_keep:
            old = known.pop(0)
            old = old if os.path.isabs(old) else os.path.join(directory, old)
            for suffix in (".index", ".data-00000-of-00001", ".meta"):
                if os.path.exists(old + suffix):
                    os.remove(old + suffix)
        with open(os.path.join(directory, "checkpoint"), "w") as f:
            f.write('model_checkpoint_path: "%s"\n' % base)
            for p in known:
                f.write('all_model_checkpoint_paths: "%s"\n' % p)
        return prefix
