"""ctypes front end for oracle/latok_oracle.c (the CPU parity checker).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Nothing under latok_b200/
imports this module.

Strings are handed to the C restatement as arrays of code points
(``str -> utf-32 -> uint32``), i.e. the same view of the text the reference gets
from PyUnicode_READ (latok.c:47-55,79).  That keeps the oracle independent of
the UTF-8 decoder inside the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
# LATOK_ORACLE_LIB: an oracle built over other class ranges (tools/ucd_check.py checks a newer-UCD library with it)
LIB_PATH = Path(os.environ.get("LATOK_ORACLE_LIB") or HERE / "_build" / "liblatok_oracle.so")
NFEAT = 25

# feature columns (latok/core/offsets.py:24-49)
(ALPHA, ALPHA_NUM, NUM, LOWER, UPPER, SPACE, SYMBOL, TWITTER, CHAR_AT, CHAR_COLON, CHAR_SLASH,
 CHAR_PERIOD, PREV_ALPHA, NEXT_ALPHA, PREV_ALPHA_NUM, NEXT_ALPHA_NUM, PREV_LOWER, NEXT_LOWER,
 PREV_SPACE, NEXT_SPACE, PREV_SYMBOL, NEXT_AT, NEXT_SLASH, AFTER_NEXT_ALPHA,
 AFTER_NEXT_SLASH) = range(NFEAT)


def build() -> Path:
    """Compile the C restatement (and, when the reference checkout is present, oracle/_ref)."""
    subprocess.run(["make", "-s", "-C", str(HERE), "oracle"], check=True)
    if os.path.isdir("/root/reference"):
        subprocess.run(["make", "-s", "-C", str(HERE), "ref"], check=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        L = C.CDLL(str(LIB_PATH))
        i8p, i32p, i64p, u32p, u8p = (C.POINTER(t) for t in (C.c_int8, C.c_int32, C.c_int64, C.c_uint32, C.c_uint8))
        L.lo_base_features.restype = C.c_uint16
        L.lo_base_features.argtypes = [C.c_uint32]
        L.lo_parse_matrix.argtypes = [u32p, C.c_int64, i8p]
        L.lo_combine_rows.argtypes = [i8p, C.c_int64, C.c_int64, C.c_int64, i8p, C.c_int, C.c_int, i8p]
        L.lo_block_mask.argtypes = [i8p, C.c_int64, i8p, C.c_int64, C.c_int64, i8p]
        L.lo_split_mask.restype = C.c_int
        L.lo_split_mask.argtypes = [i8p, C.c_int64] + [i8p, C.c_int, C.c_int] * 3 + [i8p]
        L.lo_spans.restype = C.c_int64
        L.lo_spans.argtypes = [i8p, i8p, C.c_int64, i32p, i32p]
        L.lo_token_feats.argtypes = [i8p, i32p, C.c_int64, i8p]
        L.lo_tokenize_batch_cps.restype = C.c_int64
        L.lo_tokenize_batch_cps.argtypes = ([u32p, i64p, C.c_int64] + [i8p, C.c_int, C.c_int] * 3
                                            + [i8p, i8p, i32p, i64p, i8p])
        L.lo_decode_utf8.restype = C.c_int64
        L.lo_decode_utf8.argtypes = [u8p, i64p, C.c_int64, u32p, i64p]
        _lib = L
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def combo(idx_lists) -> np.ndarray:
    """build_combo_matrix layout (latok_utils.py:27-56): int8, rows padded with -1."""
    n = max(len(r) for r in idx_lists)
    m = np.full((len(idx_lists), n), -1, dtype=np.int8)
    for i, r in enumerate(idx_lists):
        m[i, :len(r)] = r
    return m


# the default tokenizer's three rule matrices (default_tokenizer.py:49-55, 80-91, 100-102)
C_SPLIT = combo([[SPACE], [SYMBOL], [PREV_SYMBOL], [UPPER, NEXT_LOWER], [UPPER, PREV_LOWER]])
C_MASK = combo([[TWITTER, PREV_SPACE, NEXT_ALPHA],
                [CHAR_PERIOD, PREV_SPACE, NEXT_AT, AFTER_NEXT_ALPHA],
                [CHAR_AT, PREV_ALPHA_NUM, NEXT_ALPHA_NUM],
                [CHAR_COLON, NEXT_SLASH, AFTER_NEXT_SLASH, PREV_ALPHA]])
C_SYM = combo([[SYMBOL, NEXT_SPACE]])
DEFAULT_RULES = (C_SPLIT, C_MASK, C_SYM)


def _rule_args(rules):
    out = []
    for r in rules:
        r = np.ascontiguousarray(r, dtype=np.int8)
        out += [_p(r, C.c_int8), r.shape[0], r.shape[1]]
    return out, rules


def codepoints(text: str) -> np.ndarray:
    return np.frombuffer(text.encode("utf-32-le", "surrogatepass"), dtype=np.uint32).copy()


def base_features(cp: int) -> int:
    return int(lib().lo_base_features(cp))


def parse_matrix(text: str) -> np.ndarray:
    cps = codepoints(text)
    m = np.empty((len(cps), NFEAT), dtype=np.int8)
    lib().lo_parse_matrix(_p(cps, C.c_uint32), len(cps), _p(m, C.c_int8))
    return m


def combine_rows(m: np.ndarray, idx: np.ndarray) -> np.ndarray:
    assert m.dtype == np.int8 and m.ndim == 2
    idx = np.ascontiguousarray(idx, dtype=np.int8)
    out = np.empty(m.shape[1], dtype=np.int8)
    rows, cols = (idx.shape[0], idx.shape[1]) if idx.ndim == 2 else (idx.shape[0], 0)
    lib().lo_combine_rows(C.cast(m.ctypes.data, C.POINTER(C.c_int8)), m.shape[1], m.strides[0], m.strides[1],
                          _p(idx, C.c_int8), rows, cols, _p(out, C.c_int8))
    return out


def block_mask(a1: np.ndarray, a2: np.ndarray) -> np.ndarray:
    a1 = np.ascontiguousarray(a1, dtype=np.int8)
    a2 = np.ascontiguousarray(a2, dtype=np.int8)
    assert a1.shape == a2.shape and a1.ndim == 1
    out = np.empty(len(a1), dtype=np.int8)
    lib().lo_block_mask(_p(a1, C.c_int8), 1, _p(a2, C.c_int8), 1, len(a1), _p(out, C.c_int8))
    return out


def split_mask(m: np.ndarray, rules=DEFAULT_RULES) -> np.ndarray:
    m = np.ascontiguousarray(m, dtype=np.int8)
    if m.shape[0] == 0:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")  # default_tokenizer.py:132
    out = np.empty(m.shape[0], dtype=np.int8)
    args, keep = _rule_args(rules)
    lib().lo_split_mask(_p(m, C.c_int8), m.shape[0], *args, _p(out, C.c_int8))
    return out


def spans(splits: np.ndarray, m: np.ndarray):
    L = len(splits)
    sp = np.empty((max(L, 1), 2), dtype=np.int32)
    tr = np.empty((max(L, 1), 2), dtype=np.int32)
    m = np.ascontiguousarray(m, dtype=np.int8)
    splits = np.ascontiguousarray(splits, dtype=np.int8)
    T = lib().lo_spans(_p(splits, C.c_int8), _p(m, C.c_int8), L, _p(sp, C.c_int32), _p(tr, C.c_int32))
    return sp[:T].copy(), tr[:T].copy()


def token_feats(m: np.ndarray, sp: np.ndarray) -> np.ndarray:
    m = np.ascontiguousarray(m, dtype=np.int8)
    sp = np.ascontiguousarray(sp, dtype=np.int32)
    out = np.empty((len(sp), NFEAT), dtype=np.int8)
    lib().lo_token_feats(_p(m, C.c_int8), _p(sp, C.c_int32), len(sp), _p(out, C.c_int8))
    return out


def tokens(text: str, rules=DEFAULT_RULES):
    """list(tokenize(text)) of the reference (default_tokenizer.py:137-160)."""
    m = parse_matrix(text)
    s = split_mask(m, rules)
    _, tr = spans(s, m)
    return [text[a:b] for a, b in tr]


def tokenize_batch(texts, rules=DEFAULT_RULES, matrix=False, feats=True):
    """Whole-batch oracle: dict of splits[C], char_offsets[S+1], spans[T,2], tok_offsets[S+1],
    tok_feats[T,25] (if feats) and matrix[C,25] (if matrix)."""
    cps_list = [codepoints(t) for t in texts]
    S = len(texts)
    char_off = np.zeros(S + 1, dtype=np.int64)
    if S:
        char_off[1:] = np.cumsum([len(c) for c in cps_list])
    Cn = int(char_off[-1])
    cps = np.concatenate(cps_list).astype(np.uint32) if Cn else np.zeros(0, np.uint32)
    return tokenize_batch_cps(cps, char_off, rules, matrix, feats)


def tokenize_batch_cps(cps, char_off, rules=DEFAULT_RULES, matrix=False, feats=True):
    S = len(char_off) - 1
    Cn = int(char_off[-1])
    cps = np.ascontiguousarray(cps, dtype=np.uint32)
    char_off = np.ascontiguousarray(char_off, dtype=np.int64)
    splits = np.empty(Cn, dtype=np.int8)
    mat = np.empty((Cn, NFEAT), dtype=np.int8) if matrix else None
    sp = np.empty((max(Cn, 1), 2), dtype=np.int32)
    tok_off = np.zeros(S + 1, dtype=np.int64)
    tf = np.empty((max(Cn, 1), NFEAT), dtype=np.int8) if feats else None
    args, keep = _rule_args(rules)
    T = lib().lo_tokenize_batch_cps(_p(cps, C.c_uint32), _p(char_off, C.c_int64), S, *args,
                                    _p(splits, C.c_int8), _p(mat, C.c_int8), _p(sp, C.c_int32),
                                    _p(tok_off, C.c_int64), _p(tf, C.c_int8))
    out = {"splits": splits, "char_offsets": char_off, "spans": sp[:T].copy(), "tok_offsets": tok_off,
           "n_chars": Cn, "n_tokens": int(T)}
    if feats:
        out["tok_feats"] = tf[:T].copy()
    if matrix:
        out["matrix"] = mat
    return out


def decode_utf8(buf: np.ndarray, byte_off: np.ndarray):
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    byte_off = np.ascontiguousarray(byte_off, dtype=np.int64)
    S = len(byte_off) - 1
    cps = np.empty(max(len(buf), 1), dtype=np.uint32)
    char_off = np.empty(S + 1, dtype=np.int64)
    n = lib().lo_decode_utf8(_p(buf, C.c_uint8), _p(byte_off, C.c_int64), S, _p(cps, C.c_uint32), _p(char_off, C.c_int64))
    return cps[:n], char_off


def tokenize_batch_utf8(buf, byte_off, rules=DEFAULT_RULES, matrix=False, feats=True):
    """UTF-8 in -> arrays out; the like-for-like CPU leg of bench.py (kind='port')."""
    cps, char_off = decode_utf8(buf, byte_off)
    return tokenize_batch_cps(cps, char_off, rules, matrix, feats)
