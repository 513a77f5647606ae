"""Exports tests/golden/ringzk_n512.npz -- plus a few raw ring products and the representative pins of SURVEY.md 8(c) --
to tests/golden/ringzk_n512.bin, a flat little-endian file a dependency-free Rust test can parse
(shim/src/golden_vectors.rs: the crate-side test that pins these vectors against poly-ring-xnp1).

Format: magic b"RZKGOLD1", u32 entry count, then per entry
    u32 name length, name (ASCII), u8 dtype (0 = i8, 1 = i32, 2 = i64, 3 = u8), u32 ndim, u32 dims[ndim], data.

    python tests/golden/export_flat.py      # a second of work; the output is committed
The extra products are computed by the same pure-Python big-int restatement (oracle/pyref.py) that made the npz."""
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402

DT = {np.dtype(np.int8): 0, np.dtype(np.int32): 1, np.dtype(np.int64): 2, np.dtype(np.uint8): 3}
Q = 3515337053


def extra_vectors(G):
    """Raw Polynomial `*` `+` `-` pins (mat.rs:109-110, 135-136, 160-161, 176): large x large, small x large, the
    sparse challenge, x^(N-1) * x = -1 (the negacyclic wrap), and extreme residues."""
    P = pyref.Params(N=512)
    N, half = 512, (Q - 1) // 2
    a = [int(v) for v in G["a1p"][0, 0]]
    b = [int(v) for v in G["a2p"][0, 0]]
    r = [int(v) for v in G["r"][0, 1]]
    d = [int(v) for v in G["d"][0]]
    hi = [half if i % 2 == 0 else -half for i in range(N)]
    xn1 = [0] * (N - 1) + [1]
    x1 = [0, 1] + [0] * (N - 2)
    out = {
        "p_ab": pyref.p_mul(a, b, P), "p_ar": pyref.p_mul(a, r, P), "p_ad": pyref.p_mul(a, d, P),
        "p_hh": pyref.p_mul(hi, hi, P), "p_wrap": pyref.p_mul(xn1, x1, P),
        "p_a_plus_b": pyref.p_add(a, b, P), "p_a_minus_b": pyref.p_sub(a, b, P), "p_hi": hi,
        # ZqI64::<Q>::from(v).into::<i64>() for these v (canonical centred representative)
        "rep_in": [half + 1, -(half + 1), Q, -Q, Q + 5, half, -half, 2 ** 40 + 17, -(2 ** 40 + 17)],
    }
    out["rep_out"] = [pyref.center(v, Q) for v in out["rep_in"]]
    assert out["p_wrap"][0] == -1 and not any(out["p_wrap"][1:])
    assert out["rep_out"][0] == -half
    res = {}
    for k, v in out.items():
        arr = np.asarray(v, dtype=np.int64)
        res[k] = arr if k.startswith("rep") else arr.astype(np.int32)
    return res


def main():
    G = dict(np.load(os.path.join(HERE, "ringzk_n512.npz")))
    G.update(extra_vectors(G))
    entries = []
    for name in sorted(G):
        a = np.ascontiguousarray(G[name])
        if a.dtype == np.bool_:
            a = a.astype(np.uint8)
        if a.ndim == 0:
            a = a.reshape(1)
        entries.append((name, a))
    path = os.path.join(HERE, "ringzk_n512.bin")
    with open(path, "wb") as f:
        f.write(b"RZKGOLD1")
        f.write(struct.pack("<I", len(entries)))
        for name, a in entries:
            f.write(struct.pack("<I", len(name)) + name.encode())
            f.write(struct.pack("<BI", DT[a.dtype], a.ndim))
            f.write(struct.pack("<%dI" % a.ndim, *a.shape))
            f.write(a.astype(a.dtype.newbyteorder("<")).tobytes())
    print("wrote", path, os.path.getsize(path), "bytes,", len(entries), "entries")


def load_flat(path=None):
    """Parser twin of the Rust one (used by tests/test_golden_oracle.py to check the export round-trips)."""
    path = path or os.path.join(HERE, "ringzk_n512.bin")
    buf = open(path, "rb").read()
    assert buf[:8] == b"RZKGOLD1"
    (n,), off = struct.unpack_from("<I", buf, 8), 12
    inv = {v: k for k, v in DT.items()}
    out = {}
    for _ in range(n):
        (ln,) = struct.unpack_from("<I", buf, off); off += 4
        name = buf[off:off + ln].decode(); off += ln
        dt, nd = struct.unpack_from("<BI", buf, off); off += 5
        dims = struct.unpack_from("<%dI" % nd, buf, off); off += 4 * nd
        cnt = int(np.prod(dims))
        a = np.frombuffer(buf, dtype=inv[dt].newbyteorder("<"), count=cnt, offset=off).reshape(dims)
        off += cnt * inv[dt].itemsize
        out[name] = a
    assert off == len(buf)
    return out


if __name__ == "__main__":
    main()
