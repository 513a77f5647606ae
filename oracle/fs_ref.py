"""CPU restatement of the Fiat-Shamir challenge of docs/FIAT_SHAMIR.md -- TEST INFRASTRUCTURE (checker for the device kernel
ring-zk_b200/csrc/rzk_fs.cuh; never shipped or measured).  The reference has no Fiat-Shamir transform (README.md:16 names it as a
possibility); the challenge space is challenge_space.rs:12-33."""
import hashlib
import struct

import numpy as np


def key_digest(a11, a12, a22):
    h = hashlib.shake_128()
    for p in (a11, a12, a22):
        h.update(np.asarray(p).astype("<i4").tobytes())
    return h.digest(32)


def prefix(tag: str, digest: bytes, q: int, N: int, kappa: int, T: int = 0, b: int = 1, session: bytes = b"") -> bytes:
    t = tag.encode()
    assert len(t) <= 32 and len(digest) == 32 and len(session) % 8 == 0
    return t.ljust(32, b"\0") + digest + struct.pack("<QIIII", q, N, kappa, T, b) + session


def sample_in_ball(stream: bytes, N: int, kappa: int):
    """stream: enough SHAKE output; words are consumed 8 bytes at a time as the device does"""
    d = np.zeros(N, np.int8)
    signs = int.from_bytes(stream[:8], "little")
    pos = 8
    cand = []
    for i in range(N - kappa, N):
        while True:
            if not cand:
                w = int.from_bytes(stream[pos:pos + 8], "little"); pos += 8
                cand = [(w >> (16 * k)) & 0xFFFF for k in range(4)]
            j = cand.pop(0) & 0x1FF
            if j <= i:
                break
        d[i] = d[j]
        d[j] = -1 if (signs & 1) else 1
        signs >>= 1
    return d


def challenge(pre: bytes, polys, N: int = 512, kappa: int = 36):
    """polys: iterable of coefficient arrays (any integer dtype; int8 is widened), absorbed as int32 LE"""
    assert len(pre) % 8 == 0
    h = hashlib.shake_128()
    h.update(pre)
    for p in polys:
        h.update(np.asarray(p).reshape(-1).astype("<i4").tobytes())
    return sample_in_ball(h.digest(8 + 8 * 2048), N, kappa)
