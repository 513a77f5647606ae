"""ctypes binding of the C ABI in include/ringzk_b200.h (libringzk_b200.so).

This is the only compute path of the package: every call goes to the hand-written
sm_100a kernels through the C ABI.  There is no CPU fallback -- if the shared library
is missing or no CUDA device is present, loading / Engine() raises.

Host entry points take numpy arrays (C-contiguous; int32 / int8 as documented in the
header).  `*_dev` entry points take torch CUDA tensors (or anything with data_ptr())
and enqueue on a CUDA stream without synchronising.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("RZK_LIB_PATH") or os.path.join(_HERE, "_build", "libringzk_b200.so")   # RZK_LIB_PATH: A/B builds during development

RZK_OK, RZK_ERR_INVALID, RZK_ERR_UNSUPPORTED, RZK_ERR_CUDA, RZK_ERR_RANGE, RZK_ERR_NOKEY = range(6)
_ERRNAMES = {1: "RZK_ERR_INVALID", 2: "RZK_ERR_UNSUPPORTED", 3: "RZK_ERR_CUDA", 4: "RZK_ERR_RANGE", 5: "RZK_ERR_NOKEY"}

SOURCES = [os.path.join(_HERE, "csrc", f) for f in ("rzk_engine.cu", "rzk_tables.cpp")]
HEADERS = [os.path.join(_HERE, "csrc", f) for f in
           ("rzk_arith.cuh", "rzk_vm.h", "rzk_vm_exec.cuh", "rzk_programs.h", "rzk_tables.h", "rzk_sparse.cuh", "rzk_sample.cuh", "rzk_wire.cuh")] + \
          [os.path.join(_ROOT, "include", "ringzk_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
              "-Xcompiler", "-fPIC", "--shared"]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA extension in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    deps = SOURCES + HEADERS
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps)
    if stale:
        os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
        cmd = ["nvcc"] + NVCC_FLAGS
        if os.path.exists("/usr/bin/g++"):
            cmd += ["-ccbin", "/usr/bin/g++"]
        cmd += ["-o", LIB_PATH] + SOURCES
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            print(res.stdout, res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed building libringzk_b200.so")
    return LIB_PATH


def pack_r2(r):
    """rzk_pack_r2 (plain CPU loop in the library): int8 entries in [-2, 1] -> two's-complement 2-bit fields, four per byte."""
    r = np.ascontiguousarray(r, dtype=np.int8)
    out = np.empty(r.shape[:-1] + (r.shape[-1] // 4,), np.uint8)
    rc = lib().rzk_pack_r2(r.size, r.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise RzkError(rc, "rzk_pack_r2: an entry outside [-2, 1] (or a count that is not a multiple of 4)")
    return out


class RzkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{_ERRNAMES.get(code, code)}: {msg}")
        self.code = code


class RzkParams(C.Structure):
    _fields_ = [("q", C.c_int64), ("b", C.c_int64), ("N", C.c_int32), ("n", C.c_int32),
                ("k", C.c_int32), ("l", C.c_int32), ("kappa", C.c_int32)]


_lib = None
_VP = C.c_void_p

# name -> argument kinds after the engine handle ('z' size_t, 'u' uint32, 'p' pointer)
_SIGS = {
    "rzk_commit_batch": "zpppp",
    "rzk_commit_batch_r2": "zpppp",
    "rzk_unpack_r2_dev": "zppp",
    "rzk_commitment_verify_batch": "zppppp",
    "rzk_commitment_verify_batch_dev": "zpppppp",
    "rzk_open_commit_batch": "zpppppp",
    "rzk_open_respond_batch": "zpppp",
    "rzk_open_verify_batch": "zppppp",
    "rzk_linear_commit_batch": "z" + "p" * 13,
    "rzk_linear_respond_batch": "z" + "p" * 7,
    "rzk_linear_verify_batch": "z" + "p" * 10,
    "rzk_sum_commit_batch": "zu" + "p" * 13,
    "rzk_sum_respond_batch": "zu" + "p" * 7,
    "rzk_sum_verify_batch": "zu" + "p" * 10,
    "rzk_commit_batch_dev": "zppppp",
    "rzk_open_commit_batch_dev": "zppppppp",
    "rzk_open_respond_batch_dev": "zppppp",
    "rzk_open_verify_batch_dev": "zpppuppp",
    "rzk_linear_commit_batch_dev": "z" + "p" * 14,
    "rzk_linear_respond_batch_dev": "z" + "p" * 8,
    "rzk_linear_verify_batch_dev": "z" + "p" * 11,
    "rzk_sum_commit_batch_dev": "zu" + "p" * 14,
    "rzk_sum_respond_batch_dev": "zu" + "p" * 8,
    "rzk_sum_verify_batch_dev": "zu" + "p" * 11,
    "rzk_flags_to_bitmap_dev": "zpppp",
    "rzk_sample_small_dev": "ziqupp",
    "rzk_sample_gaussian_dev": "zdqupp",
    "rzk_sample_challenge_dev": "ziqupp",
    "rzk_pack_i64": "zpp",
    "rzk_unpack_i64": "zpp",
    "rzk_sync": "p",
}
_GROUP_HOST = ["rzk_commit_batch", "rzk_commit_batch_r2", "rzk_commitment_verify_batch", "rzk_open_commit_batch", "rzk_open_respond_batch",
               "rzk_open_verify_batch", "rzk_linear_commit_batch", "rzk_linear_respond_batch", "rzk_linear_verify_batch",
               "rzk_sum_commit_batch", "rzk_sum_respond_batch", "rzk_sum_verify_batch"]
_GROUP_SIGS = {n.replace("rzk_", "rzk_group_", 1): _SIGS[n] for n in _GROUP_HOST}
EXPORTS = sorted(list(_SIGS) + list(_GROUP_SIGS) + ["rzk_group_create", "rzk_group_destroy", "rzk_group_size", "rzk_group_last_error",
                                                    "rzk_group_set_key", "rzk_group_kernel_launches"] + ["rzk_default_params", "rzk_create", "rzk_destroy", "rzk_last_error", "rzk_device", "rzk_pack_r2",
                                "rzk_sigma", "rzk_commit_bound", "rzk_verify_bound", "rzk_small_limit",
                                "rzk_set_key", "rzk_host_alloc", "rzk_host_free", "rzk_kernel_launches",
                                "rzk_wire_layout", "rzk_wire_pack_dev", "rzk_wire_unpack_dev",
                                "rzk_fs_challenge_dev", "rzk_open_prove_fs_batch_dev", "rzk_open_verify_fs_batch_dev",
                                "rzk_wire_pack", "rzk_wire_unpack", "rzk_open_prove_fs_batch", "rzk_open_verify_fs_batch"])


def lib():
    """Load libringzk_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() (nvcc, sm_100a). "
                           "There is no CPU fallback for the engine.")
    L = C.CDLL(LIB_PATH)
    kinds = {"z": C.c_size_t, "u": C.c_uint32, "p": _VP, "i": C.c_int32, "q": C.c_uint64, "d": C.c_double}
    for name, sig in _SIGS.items():
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = [_VP] + [kinds[c] for c in sig]
    L.rzk_default_params.restype = RzkParams
    L.rzk_default_params.argtypes = [C.c_int32]
    L.rzk_pack_r2.restype = C.c_int
    L.rzk_pack_r2.argtypes = [C.c_size_t, _VP, _VP]
    L.rzk_create.restype = C.c_int
    L.rzk_create.argtypes = [C.POINTER(RzkParams), C.c_int, C.POINTER(_VP)]
    L.rzk_destroy.argtypes = [_VP]
    L.rzk_destroy.restype = None
    L.rzk_last_error.restype = C.c_char_p
    L.rzk_last_error.argtypes = [_VP]
    L.rzk_device.argtypes = [_VP]
    for n in ("rzk_sigma", "rzk_commit_bound", "rzk_verify_bound", "rzk_kernel_launches"):
        getattr(L, n).restype = C.c_uint64
        getattr(L, n).argtypes = [_VP]
    L.rzk_small_limit.restype = C.c_uint32
    L.rzk_small_limit.argtypes = [_VP]
    L.rzk_set_key.restype = C.c_int
    L.rzk_set_key.argtypes = [_VP, _VP, _VP]
    L.rzk_host_alloc.restype = _VP
    L.rzk_host_alloc.argtypes = [C.c_size_t]
    L.rzk_host_free.argtypes = [_VP]
    L.rzk_host_free.restype = None
    for name, sig in _GROUP_SIGS.items():
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = [_VP] + [kinds[c] for c in sig]
    L.rzk_group_create.restype = C.c_int
    L.rzk_group_create.argtypes = [C.POINTER(RzkParams), C.POINTER(C.c_int), C.c_int, C.POINTER(_VP)]
    L.rzk_group_destroy.argtypes = [_VP]
    L.rzk_group_destroy.restype = None
    L.rzk_group_size.argtypes = [_VP]
    L.rzk_group_last_error.restype = C.c_char_p
    L.rzk_group_last_error.argtypes = [_VP]
    L.rzk_group_set_key.restype = C.c_int
    L.rzk_group_set_key.argtypes = [_VP, _VP, _VP]
    L.rzk_group_kernel_launches.restype = C.c_uint64
    L.rzk_group_kernel_launches.argtypes = [_VP]
    L.rzk_wire_layout.restype = C.c_int
    L.rzk_wire_layout.argtypes = [C.c_int, C.c_uint32, _VP, C.c_size_t, C.POINTER(C.c_size_t)]
    L.rzk_wire_pack_dev.restype = C.c_int
    L.rzk_wire_pack_dev.argtypes = [_VP, C.c_size_t, _VP, C.c_size_t, _VP, C.c_int, C.c_int, C.c_int, _VP, C.c_size_t, _VP,
                                    C.POINTER(C.c_uint64), _VP]
    L.rzk_wire_unpack_dev.restype = C.c_int
    L.rzk_wire_unpack_dev.argtypes = [_VP, C.c_size_t, _VP, C.c_size_t, _VP, C.c_int, C.c_int, _VP, C.c_size_t, _VP, _VP, _VP]
    L.rzk_fs_challenge_dev.restype = C.c_int
    L.rzk_fs_challenge_dev.argtypes = [_VP, C.c_size_t, C.c_char_p, C.c_size_t, _VP, C.c_int, _VP, _VP]
    L.rzk_open_prove_fs_batch_dev.restype = C.c_int
    L.rzk_open_prove_fs_batch_dev.argtypes = [_VP, C.c_size_t, _VP, _VP, _VP, C.c_char_p, C.c_size_t, _VP, _VP, _VP, _VP, _VP, _VP]
    L.rzk_open_verify_fs_batch_dev.restype = C.c_int
    L.rzk_open_verify_fs_batch_dev.argtypes = [_VP, C.c_size_t, _VP, _VP, _VP, C.c_char_p, C.c_size_t, _VP, _VP, _VP]
    L.rzk_open_prove_fs_batch.restype = C.c_int
    L.rzk_open_prove_fs_batch.argtypes = [_VP, C.c_size_t, _VP, _VP, _VP, C.c_char_p, C.c_size_t, _VP, _VP, _VP, _VP, _VP]
    L.rzk_open_verify_fs_batch.restype = C.c_int
    L.rzk_open_verify_fs_batch.argtypes = [_VP, C.c_size_t, _VP, _VP, _VP, C.c_char_p, C.c_size_t, _VP]
    L.rzk_wire_pack.restype = C.c_int
    L.rzk_wire_pack.argtypes = [_VP, C.c_size_t, _VP, C.c_size_t, _VP, C.c_int, C.c_int, C.c_int, _VP, C.c_size_t, _VP, C.POINTER(C.c_uint64)]
    L.rzk_wire_unpack.restype = C.c_int
    L.rzk_wire_unpack.argtypes = [_VP, C.c_size_t, _VP, C.c_size_t, _VP, C.c_int, C.c_int, _VP, C.c_size_t, _VP, _VP]
    _lib = L
    return L


# ---- wire format (include/ringzk_b200.h, "wire format of the messages") ----
class WireTok(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("stream", C.c_uint32), ("poly", C.c_uint32), ("value", C.c_uint32)]


class WireStream(C.Structure):
    _fields_ = [("base", _VP), ("polys_per_item", C.c_uint32), ("dtype", C.c_uint32)]


WIRE_END, WIRE_LEN, WIRE_POLY, WIRE_TAG = range(4)
(MSG_COMMITMENT, MSG_OPENING, MSG_OPENING_F, MSG_OPEN_COMMITMENT, MSG_CHALLENGE, MSG_OPEN_RESPONSE, MSG_LINEAR_COMMITMENT,
 MSG_LINEAR_RESPONSE, MSG_SUM_COMMITMENT, MSG_SUM_RESPONSE) = range(1, 11)


def wire_layout(kind: int, T: int = 0):
    """Token list of a message kind (rzk_wire_layout; host logic only, no GPU): list of (kind, stream, poly, value)."""
    L = lib()
    n = C.c_size_t(0)
    rc = L.rzk_wire_layout(kind, T, None, 0, C.byref(n))
    if rc != RZK_OK:
        raise RzkError(rc, "rzk_wire_layout: unknown message kind or T out of range")
    toks = (WireTok * n.value)()
    rc = L.rzk_wire_layout(kind, T, toks, n.value, C.byref(n))
    if rc != RZK_OK:
        raise RzkError(rc, "rzk_wire_layout")
    return [(t.kind, t.stream, t.poly, t.value) for t in toks]


def _ptr(a, dtype=None):
    """Host numpy array or device tensor -> raw pointer (with dtype / contiguity checks)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if dtype is not None and a.dtype != dtype:
            raise TypeError(f"expected {dtype}, got {a.dtype}")
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return a.data_ptr()
    if isinstance(a, int):
        return a
    raise TypeError(type(a))


class Engine:
    """One engine per CUDA device (rzk_create / rzk_destroy)."""

    def __init__(self, N=512, device=-1, params: RzkParams | None = None):
        L = lib()
        self.L = L
        self.params = params if params is not None else L.rzk_default_params(N)
        h = _VP()
        rc = L.rzk_create(C.byref(self.params), device, C.byref(h))
        if rc != RZK_OK:
            raise RzkError(rc, (L.rzk_last_error(None) or b"").decode())
        self.h = h
        self.N = self.params.N

    def close(self):
        if getattr(self, "h", None):
            self.L.rzk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, name, *args):
        rc = getattr(self.L, name)(self.h, *args)
        if rc != RZK_OK:
            raise RzkError(rc, self.last_error())

    def last_error(self):
        return (self.L.rzk_last_error(self.h) or b"").decode()

    # ---- scalars ----
    @property
    def device(self):
        return self.L.rzk_device(self.h)

    def sigma(self):
        return int(self.L.rzk_sigma(self.h))

    def commit_bound(self):
        return int(self.L.rzk_commit_bound(self.h))

    def verify_bound(self):
        return int(self.L.rzk_verify_bound(self.h))

    def small_limit(self):
        return int(self.L.rzk_small_limit(self.h))

    def kernel_launches(self):
        return int(self.L.rzk_kernel_launches(self.h))

    # ---- key ----
    def set_key(self, a1, a2):
        """a1 [n][k][N], a2 [l][k][N] int64 as the reference stores them (commit.rs:19-60)."""
        a1 = np.ascontiguousarray(a1, dtype=np.int64)
        a2 = np.ascontiguousarray(a2, dtype=np.int64)
        rc = self.L.rzk_set_key(self.h, a1.ctypes.data, a2.ctypes.data)
        if rc != RZK_OK:
            raise RzkError(rc, (self.L.rzk_last_error(self.h) or b"").decode())

    def set_key_blocks(self, a1p, a2p):
        """Random blocks only: a1p [1][2][N], a2p [1][1][N] -> [1 | a1p], [0 | 1 | a2p]."""
        N = self.N
        a1 = np.zeros((1, 3, N), np.int64)
        a2 = np.zeros((1, 3, N), np.int64)
        a1[0, 0, 0] = 1
        a1[0, 1:] = np.asarray(a1p, dtype=np.int64).reshape(2, N)
        a2[0, 1, 0] = 1
        a2[0, 2] = np.asarray(a2p, dtype=np.int64).reshape(N)
        self.set_key(a1, a2)
        return a1, a2

    # ---- host API (numpy) ----
    def commit(self, x, r):
        B, N = x.shape[0], self.N
        c = np.empty((B, 2, N), np.int32)
        ok = np.zeros((B + 7) // 8, np.uint8)
        self._call("rzk_commit_batch", B, _ptr(x, np.int32), _ptr(r, np.int8), _ptr(c), _ptr(ok))
        return c, ok

    def commit_r2(self, x, r2):
        """rzk_commit_batch_r2: the randomness packed at 2 bits per coefficient (pack_r2), [B][3][N/4] uint8."""
        B, N = x.shape[0], self.N
        c = np.empty((B, 2, N), np.int32)
        ok = np.zeros((B + 7) // 8, np.uint8)
        self._call("rzk_commit_batch_r2", B, _ptr(x, np.int32), _ptr(r2, np.uint8), _ptr(c), _ptr(ok))
        return c, ok

    def commitment_verify(self, c, x, r, f=None):
        """Commitment::verify (commit.rs:173-210) for a batch; f None or [B][N] int8."""
        B = c.shape[0]
        bm = np.zeros((B + 7) // 8, np.uint8)
        self._call("rzk_commitment_verify_batch", B, _ptr(c, np.int32), _ptr(x, np.int32), _ptr(r, np.int8),
                   _ptr(f, np.int8) if f is not None else None, _ptr(bm))
        return bm

    def open_commit(self, x, r, y):
        B, N = x.shape[0], self.N
        c = np.empty((B, 2, N), np.int32)
        t = np.empty((B, 1, N), np.int32)
        ok = np.zeros((B + 7) // 8, np.uint8)
        self._call("rzk_open_commit_batch", B, _ptr(x, np.int32), _ptr(r, np.int8), _ptr(y, np.int32),
                   _ptr(c), _ptr(t), _ptr(ok))
        return c, t, ok

    def open_respond(self, y, r, d):
        B = y.shape[0]
        z = np.empty((B, 3, self.N), np.int32)
        self._call("rzk_open_respond_batch", B, _ptr(y, np.int32), _ptr(r, np.int8), _ptr(d, np.int8), _ptr(z))
        return z

    def open_verify(self, z, t, c1, d):
        B = z.shape[0]
        bm = np.zeros((B + 7) // 8, np.uint8)
        self._call("rzk_open_verify_batch", B, _ptr(z, np.int32), _ptr(t, np.int32), _ptr(c1, np.int32),
                   _ptr(d, np.int8), _ptr(bm))
        return bm

    def linear_commit(self, g, x, rp, r, y, yp):
        B, N = x.shape[0], self.N
        o = dict(gx=np.empty((B, 1, N), np.int32), cp=np.empty((B, 2, N), np.int32), c=np.empty((B, 2, N), np.int32),
                 t=np.empty((B, 1, N), np.int32), tp=np.empty((B, 1, N), np.int32), u=np.empty((B, 1, N), np.int32),
                 ok=np.zeros((B + 7) // 8, np.uint8))
        self._call("rzk_linear_commit_batch", B, _ptr(g, np.int32), _ptr(x, np.int32), _ptr(rp, np.int8), _ptr(r, np.int8),
                   _ptr(y, np.int32), _ptr(yp, np.int32), _ptr(o["gx"]), _ptr(o["cp"]), _ptr(o["c"]), _ptr(o["t"]),
                   _ptr(o["tp"]), _ptr(o["u"]), _ptr(o["ok"]))
        return o

    def linear_respond(self, y, yp, r, rp, d):
        B = y.shape[0]
        z = np.empty((B, 3, self.N), np.int32)
        zp = np.empty((B, 3, self.N), np.int32)
        self._call("rzk_linear_respond_batch", B, _ptr(y, np.int32), _ptr(yp, np.int32), _ptr(r, np.int8),
                   _ptr(rp, np.int8), _ptr(d, np.int8), _ptr(z), _ptr(zp))
        return z, zp

    def linear_verify(self, z, zp, c, cp, g, t, tp, u, d):
        B = z.shape[0]
        bm = np.zeros((B + 7) // 8, np.uint8)
        self._call("rzk_linear_verify_batch", B, _ptr(z, np.int32), _ptr(zp, np.int32), _ptr(c, np.int32),
                   _ptr(cp, np.int32), _ptr(g, np.int32), _ptr(t, np.int32), _ptr(tp, np.int32), _ptr(u, np.int32),
                   _ptr(d, np.int8), _ptr(bm))
        return bm

    def sum_commit(self, gs, xs, rp, rs, ys, yp):
        B, T, N = gs.shape[0], gs.shape[1], self.N
        o = dict(xp=np.empty((B, 1, N), np.int32), cp=np.empty((B, 2, N), np.int32), cs=np.empty((B, T, 2, N), np.int32),
                 ts=np.empty((B, T, 1, N), np.int32), tp=np.empty((B, 1, N), np.int32), u=np.empty((B, 1, N), np.int32),
                 ok=np.zeros((B + 7) // 8, np.uint8))
        self._call("rzk_sum_commit_batch", B, T, _ptr(gs, np.int32), _ptr(xs, np.int32), _ptr(rp, np.int8),
                   _ptr(rs, np.int8), _ptr(ys, np.int32), _ptr(yp, np.int32), _ptr(o["xp"]), _ptr(o["cp"]),
                   _ptr(o["cs"]), _ptr(o["ts"]), _ptr(o["tp"]), _ptr(o["u"]), _ptr(o["ok"]))
        return o

    def sum_respond(self, ys, yp, rs, rp, d):
        B, T = ys.shape[0], ys.shape[1]
        zs = np.empty((B, T, 3, self.N), np.int32)
        zp = np.empty((B, 3, self.N), np.int32)
        self._call("rzk_sum_respond_batch", B, T, _ptr(ys, np.int32), _ptr(yp, np.int32), _ptr(rs, np.int8),
                   _ptr(rp, np.int8), _ptr(d, np.int8), _ptr(zs), _ptr(zp))
        return zs, zp

    def sum_verify(self, zs, zp, cs, cp, gs, ts, tp, u, d):
        B, T = zs.shape[0], zs.shape[1]
        bm = np.zeros((B + 7) // 8, np.uint8)
        self._call("rzk_sum_verify_batch", B, T, _ptr(zs, np.int32), _ptr(zp, np.int32), _ptr(cs, np.int32),
                   _ptr(cp, np.int32), _ptr(gs, np.int32), _ptr(ts, np.int32), _ptr(tp, np.int32), _ptr(u, np.int32),
                   _ptr(d, np.int8), _ptr(bm))
        return bm

    def pack_i64(self, a):
        a = np.ascontiguousarray(a, dtype=np.int64)
        out = np.empty(a.shape, np.int32)
        self._call("rzk_pack_i64", a.size, _ptr(a), _ptr(out))
        return out

    def unpack_i64(self, a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        out = np.empty(a.shape, np.int64)
        self._call("rzk_unpack_i64", a.size, _ptr(a), _ptr(out))
        return out

    # ---- device API (torch tensors / raw pointers) ----
    def dev(self, name, *args, stream=0):
        """Call rzk_<name>_dev; tensors are passed by data_ptr(); ints stay ints."""
        sig = _SIGS[f"rzk_{name}_dev"]
        conv = []
        for kind, a in zip(sig, list(args) + [stream]):
            conv.append(a if kind in "zuiqd" else _ptr(a))
        self._call(f"rzk_{name}_dev", *conv)

    def sync(self, stream=0):
        self._call("rzk_sync", stream)

    # ---- Fiat-Shamir challenges on the device (docs/FIAT_SHAMIR.md) ----
    def fs_challenge(self, prefix: bytes, segs, d, stream=0):
        """d [B][N] int8 (device) = SampleInBall(SHAKE128(prefix || segs...)), one hash per item."""
        import torch
        cs = (WireStream * len(segs))()
        for i, a in enumerate(segs):
            cs[i] = WireStream(a.data_ptr(), int(np.prod(a.shape[1:-1])) if a.dim() > 2 else 1, 1 if a.dtype == torch.int8 else 0)
        rc = self.L.rzk_fs_challenge_dev(self.h, segs[0].shape[0], prefix, len(prefix), cs, len(cs), d.data_ptr(), stream)
        if rc != RZK_OK:
            raise RzkError(rc, self.last_error())

    def open_prove_fs(self, x, r, y, prefix: bytes, c, t, d, z, flags, stream=0):
        rc = self.L.rzk_open_prove_fs_batch_dev(self.h, x.shape[0], _ptr(x), _ptr(r), _ptr(y), prefix, len(prefix), _ptr(c), _ptr(t),
                                                _ptr(d), _ptr(z), _ptr(flags), stream)
        if rc != RZK_OK:
            raise RzkError(rc, self.last_error())

    def open_verify_fs(self, c, t, z, prefix: bytes, d, flags, stream=0):
        rc = self.L.rzk_open_verify_fs_batch_dev(self.h, c.shape[0], _ptr(c), _ptr(t), _ptr(z), prefix, len(prefix), _ptr(d), _ptr(flags), stream)
        if rc != RZK_OK:
            raise RzkError(rc, self.last_error())

    # ---- host forms of the Fiat-Shamir proofs and of the wire format (numpy arrays) ----
    def open_prove_fs_host(self, x, r, y, prefix: bytes):
        """rzk_open_prove_fs_batch: numpy x [B][1][N] i32, r [B][3][N] i8, y [B][3][N] i32 -> dict(c, t, d, z, ok bitmap)"""
        B, N = x.shape[0], self.N
        c, t, z = np.empty((B, 2, N), np.int32), np.empty((B, 1, N), np.int32), np.empty((B, 3, N), np.int32)
        d, ok = np.empty((B, N), np.int8), np.zeros((B + 7) // 8, np.uint8)
        rc = self.L.rzk_open_prove_fs_batch(self.h, B, _ptr(x, np.int32), _ptr(r, np.int8), _ptr(y, np.int32), prefix, len(prefix),
                                            _ptr(c), _ptr(t), _ptr(d), _ptr(z), _ptr(ok))
        if rc != RZK_OK:
            raise RzkError(rc, self.last_error())
        return dict(c=c, t=t, d=d, z=z, ok=ok)

    def open_verify_fs_host(self, c, t, z, prefix: bytes):
        B = c.shape[0]
        bm = np.zeros((B + 7) // 8, np.uint8)
        rc = self.L.rzk_open_verify_fs_batch(self.h, B, _ptr(c, np.int32), _ptr(t, np.int32), _ptr(z, np.int32), prefix, len(prefix), _ptr(bm))
        if rc != RZK_OK:
            raise RzkError(rc, self.last_error())
        return bm

    @staticmethod
    def _wire_args_host(kind, T, arrays):
        toks = wire_layout(kind, T)
        ctoks = (WireTok * len(toks))(*[WireTok(*t) for t in toks])
        cs = (WireStream * len(arrays))()
        for i, a in enumerate(arrays):
            if a.dtype not in (np.int32, np.int8) or not a.flags["C_CONTIGUOUS"]:
                raise TypeError("wire streams are C-contiguous int32 / int8 arrays [B][polys][N]")
            cs[i] = WireStream(a.ctypes.data, int(np.prod(a.shape[1:-1])) if a.ndim > 2 else 1, 1 if a.dtype == np.int8 else 0)
        return ctoks, cs

    def wire_pack_host(self, kind, arrays, T=0, elem_bytes=8, trim=True):
        """rzk_wire_pack: numpy arrays in -> (bytes, offsets [B + 1] uint64)"""
        B = arrays[0].shape[0]
        ctoks, cs = self._wire_args_host(kind, T, arrays)
        off = np.zeros(B + 1, np.uint64)
        total = C.c_uint64(0)
        args = (self.h, B, ctoks, len(ctoks), cs, len(cs), elem_bytes, 1 if trim else 0)
        rc = self.L.rzk_wire_pack(*args, None, 0, off.ctypes.data, C.byref(total))
        if rc != RZK_OK:
            raise RzkError(rc, self.last_error())
        out = np.zeros(max(int(total.value), 1), np.uint8)
        rc = self.L.rzk_wire_pack(*args, out.ctypes.data, int(total.value), off.ctypes.data, C.byref(total))
        if rc != RZK_OK:
            raise RzkError(rc, self.last_error())
        return out[: int(total.value)].tobytes(), off

    def wire_unpack_host(self, kind, data: bytes, offsets, arrays, T=0, elem_bytes=8):
        """rzk_wire_unpack into the preallocated numpy `arrays`; returns the ok bitmap (bit i: item i parsed)"""
        B = len(offsets) - 1
        ctoks, cs = self._wire_args_host(kind, T, arrays)
        buf = np.frombuffer(data, np.uint8)
        off = np.ascontiguousarray(offsets, np.uint64)
        ok = np.zeros((B + 7) // 8, np.uint8)
        rc = self.L.rzk_wire_unpack(self.h, B, ctoks, len(ctoks), cs, len(cs), elem_bytes, buf.ctypes.data, buf.size, off.ctypes.data, ok.ctypes.data)
        if rc != RZK_OK:
            raise RzkError(rc, self.last_error())
        return ok

    # ---- wire format: device tensors in, device bytes out (and back) ----
    @staticmethod
    def _wire_args(kind, T, streams):
        toks = wire_layout(kind, T)
        ctoks = (WireTok * len(toks))(*[WireTok(*t) for t in toks])
        cs = (WireStream * len(streams))()
        for i, a in enumerate(streams):
            import torch
            if a.dtype not in (torch.int32, torch.int8) or not a.is_contiguous():
                raise TypeError("wire streams are contiguous int32 / int8 device tensors [B][polys][N]")
            cs[i] = WireStream(a.data_ptr(), int(np.prod(a.shape[1:-1])) if a.dim() > 2 else 1, 1 if a.dtype == torch.int8 else 0)
        return ctoks, cs

    def wire_pack(self, kind, streams, T=0, elem_bytes=8, trim=True, stream=0):
        """Messages of `kind` for the B items of `streams` (device tensors, numbered as in the header) in the reference's
        bincode layout: returns (bytes: uint8 device tensor, offsets: int64 device tensor of B + 1 entries)."""
        import torch
        B = streams[0].shape[0]
        ctoks, cs = self._wire_args(kind, T, streams)
        offsets = torch.empty(B + 1, dtype=torch.int64, device=streams[0].device)
        total = C.c_uint64(0)
        args = (self.h, B, ctoks, len(ctoks), cs, len(cs), elem_bytes, 1 if trim else 0)
        rc = self.L.rzk_wire_pack_dev(*args, None, 0, offsets.data_ptr(), C.byref(total), stream)
        if rc != RZK_OK:
            raise RzkError(rc, self.last_error())
        out = torch.empty(max(int(total.value), 1), dtype=torch.uint8, device=streams[0].device)
        rc = self.L.rzk_wire_pack_dev(*args, out.data_ptr(), int(total.value), offsets.data_ptr(), C.byref(total), stream)
        if rc != RZK_OK:
            raise RzkError(rc, self.last_error())
        return out[: int(total.value)], offsets

    def wire_unpack(self, kind, data, offsets, streams, T=0, elem_bytes=8, stream=0):
        """Parses B messages (uint8 device tensor + int64 offsets [B + 1]) into the preallocated `streams`;
        returns the per-item flags (bit 0: malformed)."""
        import torch
        B = offsets.numel() - 1
        ctoks, cs = self._wire_args(kind, T, streams)
        flags = torch.zeros(max(B, 1), dtype=torch.int32, device=data.device)
        rc = self.L.rzk_wire_unpack_dev(self.h, B, ctoks, len(ctoks), cs, len(cs), elem_bytes, data.data_ptr(), data.numel(),
                                        offsets.data_ptr(), flags.data_ptr(), stream)
        if rc != RZK_OK:
            raise RzkError(rc, self.last_error())
        return flags[:B]


def unpack_bitmap(bm, B):
    """bitmap (uint8, LSB first) -> bool array of length B."""
    return np.unpackbits(np.asarray(bm, dtype=np.uint8), bitorder="little")[:B].astype(bool)



class Group(Engine):
    """Several devices behind one handle (rzk_group_*): one engine and one host worker thread per listed device,
    the batch split into contiguous 8-aligned item ranges.  Same host methods as Engine (commit, open_*, linear_*,
    sum_*, commitment_verify); the `_dev` entry points belong to single engines."""

    def __init__(self, device_ids, N=512, params: RzkParams | None = None):
        self.L = lib()
        self.params = params if params is not None else self.L.rzk_default_params(N)
        self.N = self.params.N
        ids = (C.c_int * len(device_ids))(*device_ids)
        h = _VP()
        rc = self.L.rzk_group_create(C.byref(self.params), ids, len(device_ids), C.byref(h))
        if rc != RZK_OK:
            raise RzkError(rc, (self.L.rzk_last_error(None) or b"").decode())
        self.h = h
        self.devices = list(device_ids)

    def close(self):
        if getattr(self, "h", None):
            self.L.rzk_group_destroy(self.h)
            self.h = None

    def _call(self, name, *args):
        gname = name.replace("rzk_", "rzk_group_", 1)
        if gname not in _GROUP_SIGS:
            raise AttributeError(f"{name} has no group form")
        rc = getattr(self.L, gname)(self.h, *args)
        if rc != RZK_OK:
            raise RzkError(rc, (self.L.rzk_group_last_error(self.h) or b"").decode())

    def size(self):
        return int(self.L.rzk_group_size(self.h))

    def kernel_launches(self):
        return int(self.L.rzk_group_kernel_launches(self.h))

    def set_key(self, a1, a2):
        a1 = np.ascontiguousarray(a1, dtype=np.int64)
        a2 = np.ascontiguousarray(a2, dtype=np.int64)
        rc = self.L.rzk_group_set_key(self.h, a1.ctypes.data, a2.ctypes.data)
        if rc != RZK_OK:
            raise RzkError(rc, (self.L.rzk_group_last_error(self.h) or b"").decode())
