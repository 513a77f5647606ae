"""The C-ABI library loads and exports every symbol include/ringzk_b200.h declares.
No compute calls are made here (no GPU in this tier)."""
import ctypes
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
engine = importlib.import_module("ring-zk_b200.engine")


def header_functions():
    src = open(os.path.join(ROOT, "include", "ringzk_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rzk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_header():
    engine.build()
    L = ctypes.CDLL(engine.LIB_PATH)
    names = header_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"missing export {n}"
    # the binding and the header agree on the surface
    assert set(engine.EXPORTS) == set(names)


def test_default_params_and_no_device_error():
    L = engine.lib()
    P = L.rzk_default_params(512)
    assert (P.q, P.b, P.N, P.n, P.k, P.l, P.kappa) == (3515337053, 1, 512, 1, 3, 1, 36)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        # the product path fails loudly without a CUDA device: no CPU fallback
        with pytest.raises(engine.RzkError) as ei:
            engine.Engine(N=512)
        assert ei.value.code == engine.RZK_ERR_CUDA


def test_pack_r2_host_helper():
    """rzk_pack_r2 (pure CPU): two's-complement 2-bit fields, low bits first; entries outside [-2, 1] are refused."""
    import numpy as np
    r = np.array([0, 1, -1, -2, 1, 1, 0, -1], np.int8)
    out = engine.pack_r2(r)
    assert out.tolist() == [0 | (1 << 2) | (3 << 4) | (2 << 6), 1 | (1 << 2) | (0 << 4) | (3 << 6)]
    rng = np.random.default_rng(0)
    big = rng.integers(-2, 2, size=(5, 3, 512)).astype(np.int8)
    p = engine.pack_r2(big)
    fields = np.stack([(p >> (2 * k)) & 3 for k in range(4)], axis=-1).reshape(big.shape).astype(np.int8)
    assert (((fields ^ 2) - 2) == big).all()
    with pytest.raises(engine.RzkError):
        engine.pack_r2(np.array([0, 0, 0, 2], np.int8))
    with pytest.raises(engine.RzkError):
        engine.pack_r2(np.array([0, 0, 0], np.int8))


def test_unsupported_params_rejected():
    L = engine.lib()
    P = L.rzk_default_params(16)
    h = ctypes.c_void_p()
    assert L.rzk_create(ctypes.byref(P), -1, ctypes.byref(h)) == engine.RZK_ERR_UNSUPPORTED
    assert b"N=512" in L.rzk_last_error(None)


@pytest.mark.parametrize("b,kappa,admissible", [(1, 36, True), (2, 36, True), (2, 37, True), (3, 36, False), (5, 36, False),
                                                 (11, 36, False), (127, 1, False), (74, 1, True), (75, 1, False), (24, 3, True)])
def test_parameter_admissibility(b, kappa, admissible):
    """rzk_create derives the exactness limits from the actual (b, kappa): sigma = 429 b kappa / 36 must leave
    rzk_small_limit() = 320,245 at least 10 sigma away (b * kappa <= 74), and A1.z of a response at the norm bound must
    stay inside the two-prime CRT range.  Anything else is RZK_ERR_UNSUPPORTED (the shim then keeps the CPU path),
    never a silently wrapped result.  The check precedes the device query, so it is testable without a GPU."""
    L = engine.lib()
    P = L.rzk_default_params(512)
    P.b, P.kappa = b, kappa
    h = ctypes.c_void_p()
    rc = L.rzk_create(ctypes.byref(P), -1, ctypes.byref(h))
    if h:
        L.rzk_destroy(h)
    if admissible:
        assert rc in (engine.RZK_OK, engine.RZK_ERR_CUDA), L.rzk_last_error(None)
    else:
        assert rc == engine.RZK_ERR_UNSUPPORTED
        assert b"b*kappa" in L.rzk_last_error(None) or b"<= 127" in L.rzk_last_error(None)


def test_rust_ffi_block_is_generated_from_the_header_and_complete():
    """shim/src/b200/ffi.rs declares exactly the header's functions (it is regenerated here and compared byte for byte), and every
    `ffi::rzk_*` call in the hand-written shim modules names a declared function with the right number of arguments."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(ROOT, "tools", "gen_rust_ffi.py"))
    gen = importlib.util.module_from_spec(spec); spec.loader.exec_module(gen)
    text, names = gen.generate()
    assert open(os.path.join(ROOT, "shim", "src", "b200", "ffi.rs")).read() == text, "run python tools/gen_rust_ffi.py"
    assert sorted(names) == header_functions() and len(names) == len(set(names))
    arity = {m.group(1): (0 if not m.group(2).strip() else m.group(2).count(":")) for m in re.finditer(r"pub fn (rzk_\w+)\((.*?)\)", text)}
    used = set()
    for dirpath, _, files in os.walk(os.path.join(ROOT, "shim")):
        for f in files:
            if f.endswith(".rs") and f != "ffi.rs":
                src = open(os.path.join(dirpath, f)).read()
                for m in re.finditer(r"ffi::(rzk_\w+)\(", src):
                    name = m.group(1)
                    assert name in arity, f"{f}: {name} is not in the header"
                    # count top-level commas of the call
                    depth, i, commas = 1, m.end(), 0
                    while depth:
                        ch = src[i]
                        depth += ch in "([{"
                        depth -= ch in ")]}"
                        commas += (ch == "," and depth == 1)
                        i += 1
                    nargs = commas + (1 if src[m.end():i - 1].strip() else 0)
                    assert nargs == arity[name], f"{f}: {name} called with {nargs} arguments, declared with {arity[name]}"
                    used.add(name)
    # the shim drives every host protocol entry point, single engine and group
    host = [n for n in names if n.endswith("_batch")]
    assert set(host) <= used, sorted(set(host) - used)
