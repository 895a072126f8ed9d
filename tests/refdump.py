"""Reader for the binary dumps written by oracle/ref_harness.cpp (test infrastructure only).

Record layout: [u32 name_len][name][u8 dtype 'd'|'i'|'q'][u32 ndim][i64 dims...][raw little-endian data],
after an 8-byte magic "AMDGDUMP".  Per-element arrays are sorted by ascending Hash::hash_key.
"""
import lzma
import struct
import numpy as np

_DT = {b"d": np.float64, b"i": np.int32, b"q": np.int64}


def load(path):
    out = {}
    opener = lzma.open if path.endswith(".xz") else open
    with opener(path, "rb") as f:
        buf = f.read()
    assert buf[:8] == b"AMDGDUMP", "not an AMDG dump"
    pos = 8
    while pos < len(buf):
        (nl,) = struct.unpack_from("<I", buf, pos); pos += 4
        name = buf[pos:pos + nl].decode(); pos += nl
        dt = _DT[buf[pos:pos + 1]]; pos += 1
        (nd,) = struct.unpack_from("<I", buf, pos); pos += 4
        dims = struct.unpack_from("<%dq" % nd, buf, pos); pos += 8 * nd
        n = int(np.prod(dims)) if nd else 1
        arr = np.frombuffer(buf, dtype=dt, count=n, offset=pos).reshape(dims).copy()
        pos += n * arr.itemsize
        out[name] = arr
    return out


_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return x ^ (x >> np.uint64(31))


def field(seed, hash_key, level, vec_num, block):
    """The stateless pseudo-random coefficient field of ref_harness.cpp::field_value:
    U(-1,1) * 2^-(sum of levels), keyed by (hash_key, vec, local index).  Returns [n_elem, vec_num, block]."""
    with np.errstate(over="ignore"):
        hk = hash_key.astype(np.int64).astype(np.uint64) & np.uint64(0xFFFFFFFF)
        key = (hk[:, None, None] << np.uint64(20)) + (np.arange(vec_num, dtype=np.uint64)[None, :, None] << np.uint64(16)) \
            + np.arange(block, dtype=np.uint64)[None, None, :]
        h = _splitmix64(np.uint64(seed) ^ _splitmix64(key))
    u = (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    scale = np.ldexp(1.0, -level.sum(axis=1).astype(np.int64))
    return (2.0 * u - 1.0) * scale[:, None, None]
