"""numpy restatement of the optional on-device samplers (ring-zk_b200/csrc/rzk_sample.cuh): Philox4x32-10 and the
three draws built on it.  Test infrastructure: the GPU tests compare the device output with these bit for bit
(small, challenge) or up to the last-ulp differences of log / cos between libm and CUDA (gaussian)."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox(c0, c1, c2, c3, seed):
    """vectorised Philox4x32-10; counters are uint64 arrays holding 32-bit values; returns four uint64 arrays"""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return c0, c1, c2, c3


def sample_small(n_polys, b, seed, tag, N=512):
    rng = 2 * b + 1
    thresh = (2 ** 32) % rng
    poly = np.repeat(np.arange(n_polys, dtype=np.uint64), N)
    i = np.tile(np.arange(N, dtype=np.uint64), n_polys)
    out = np.zeros(n_polys * N, np.int64)
    todo = np.ones(n_polys * N, bool)
    for attempt in range(8):
        if not todo.any():
            break
        words = philox(i[todo] >> np.uint64(2), poly[todo] & MASK, poly[todo] >> np.uint64(32), np.uint64((tag << 8) | attempt), seed)
        lane = (i[todo] & np.uint64(3)).astype(np.int64)
        u = np.choose(lane, words)
        m = u * np.uint64(rng)
        out[todo] = (m >> np.uint64(32)).astype(np.int64)
        acc = (m & MASK) >= np.uint64(thresh)
        idx = np.nonzero(todo)[0]
        todo[idx[acc]] = False
    return (out - b).astype(np.int8).reshape(n_polys, N)


def sample_gaussian(n_polys, sigma, seed, tag, N=512):
    poly = np.repeat(np.arange(n_polys, dtype=np.uint64), N // 2)
    g = np.tile(np.arange(N // 2, dtype=np.uint64), n_polys)
    x, y, z, w = philox(g, poly & MASK, poly >> np.uint64(32), np.uint64(tag << 8), seed)
    u1 = ((((x << np.uint64(32)) | y) >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)
    u2 = ((((z << np.uint64(32)) | w) >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)
    rad = sigma * np.sqrt(-2.0 * np.log(u1))
    ang = 6.283185307179586476925286766559 * u2
    out = np.empty((n_polys * N // 2, 2), np.int64)
    out[:, 0] = np.trunc(rad * np.cos(ang)).astype(np.int64)
    out[:, 1] = np.trunc(rad * np.sin(ang)).astype(np.int64)
    return out.reshape(n_polys, N).astype(np.int32)


def sample_challenge(n_items, kappa, seed, tag, N=512):
    out = np.zeros((n_items, N), np.int8)
    want = min(kappa, N)
    for it in range(n_items):
        have, block = 0, 0
        while have < want:
            ws = philox(np.uint64(block), np.uint64(it & 0xFFFFFFFF), np.uint64(it >> 32), np.uint64(tag << 8), seed)
            for wv in ws:
                wv = int(wv)
                if have >= want:
                    break
                pos = (wv >> 23) & (N - 1)
                if out[it, pos] == 0:
                    out[it, pos] = 1 if (wv & 1) else -1
                    have += 1
            block += 1
    return out
