"""Driver for the *compiled reference* (oracle/_ref): the reference's own latok.c,
built unmodified by `make -C oracle ref`, exposed as the extension module
``latok.latok`` with its three functions (latok.c:373-378).

TEST INFRASTRUCTURE ONLY (same rule as oracle.py).

Two layers:
  * ``ext()``        -- the compiled extension itself (travels to the GPU box as a
                        built artefact; used for parity pinning and as the
                        cpu_baseline of kind "reference").
  * ``ref_python()`` -- the reference's own Python glue (default_tokenizer.py),
                        importable only where /root/reference exists (the build
                        container); used to pin the glue restated below.
The glue below restates default_tokenizer.py:113-191 on top of ``ext()`` so that
the compiled reference can be driven on a box that has no reference checkout.
"""
from __future__ import annotations

import importlib
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_PKG = HERE / "_ref"
REF_SRC = Path("/root/reference")

_ext = None
_refpy = None


def available() -> bool:
    return any((REF_PKG / "latok").glob("latok*.so"))


def _import_pkg():
    if "latok" in sys.modules and not str(getattr(sys.modules["latok"], "__file__", "")).startswith(str(REF_PKG)):
        raise RuntimeError("another 'latok' package is already imported; run the reference driver in its own process")
    if str(REF_PKG) not in sys.path:
        sys.path.insert(0, str(REF_PKG))
    return importlib.import_module("latok")


def ext():
    """The compiled reference extension module (latok.latok)."""
    global _ext
    if _ext is None:
        if not available():
            raise RuntimeError("oracle/_ref is not built (make -C oracle ref needs /root/reference)")
        _import_pkg()
        _ext = importlib.import_module("latok.latok")
    return _ext


def ref_python():
    """The reference's own default_tokenizer module, or None where /root/reference is absent."""
    global _refpy
    if _refpy is None:
        if not (REF_SRC / "latok" / "core" / "default_tokenizer.py").exists():
            return None
        ext()
        pkg = _import_pkg()
        if str(REF_SRC / "latok") not in pkg.__path__:
            pkg.__path__.append(str(REF_SRC / "latok"))
        _refpy = importlib.import_module("latok.core.default_tokenizer")
    return _refpy


def _combo(idx_lists):
    n = max(len(r) for r in idx_lists)
    m = np.full((len(idx_lists), n), -1, dtype=np.int8)
    for i, r in enumerate(idx_lists):
        m[i, :len(r)] = r
    return m


# column numbers from latok/core/offsets.py:24-49
_C_SPLIT = _combo([[5], [6], [20], [4, 17], [4, 16]])               # default_tokenizer.py:49-55
_C_MASK = _combo([[7, 18, 13], [11, 18, 21, 23], [8, 14, 15], [9, 22, 24, 12]])  # :80-91
_C_SYM = _combo([[6, 19]])                                             # :100-102


def gen_split_mask(m: np.ndarray) -> np.ndarray:
    """default_tokenizer.py:113-134 on the compiled extension."""
    e = ext()
    mt = m.T
    splits = e._combine_matrix_rows(mt, _C_SPLIT) * e._gen_block_mask(e._combine_matrix_rows(mt, _C_MASK), mt[5])
    splits += e._combine_matrix_rows(mt, _C_SYM)
    splits[0] = 1
    return splits


def split_positions(text: str) -> np.ndarray:
    """The like-for-like array-output call of BASELINE.md section 3 (i)."""
    return np.nonzero(gen_split_mask(ext()._gen_parse_matrix(text)))[0]


def tokenize(text: str):
    """default_tokenizer.py:137-160."""
    nz = split_positions(text)
    s, e = nz[0], 0
    for e in nz[1:]:
        tok = text[s:e].strip()
        if tok:
            yield tok
        s = e
    tok = text[e:].strip()
    if tok:
        yield tok


def featurize_arrays(text: str):
    """default_tokenizer.py:163-191 as arrays: (spans[T,2], feats[T,25]).  Valid only
    while every position fits int8 (len(text) <= 127)."""
    e = ext()
    m = e._gen_parse_matrix(text)
    nz = np.nonzero(gen_split_mask(m))[0]
    bounds = list(nz) + [len(text)]
    sp, ft = [], []
    for a, b in zip(bounds[:-1], bounds[1:]):
        if text[a:b].strip():
            sp.append((a, b))
            ft.append(e._combine_matrix_rows(m, np.arange(a, b, dtype=np.int8)))
    return (np.array(sp, dtype=np.int32).reshape(-1, 2),
            np.array(ft, dtype=np.int8).reshape(-1, 25))
