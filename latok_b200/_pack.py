"""list[str] -> (flat UTF-8 uint8 buffer, int64 offsets[S+1]) through the C packer `_pypack` (csrc/latok_pypack.c,
built in-tree by latok_b200.build next to liblatok_b200.so).  Host-side marshalling only."""
from __future__ import annotations

import importlib.util
import sysconfig
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

PKG = Path(__file__).resolve().parent
EXT_PATH = PKG / ("_pypack" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))
_mod = None


def load():
    global _mod
    if _mod is None:
        if not EXT_PATH.exists():
            from . import build as _build
            _build.build_pypack()
        spec = importlib.util.spec_from_file_location("latok_b200._pypack", EXT_PATH)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _mod = m
    return _mod


_ratio = [192.0]      # UTF-8 bytes per string seen so far (sizes the buffer of the next call)


def pack(texts: Sequence[str], out: Optional[np.ndarray] = None):
    m = load()
    if not isinstance(texts, (list, tuple)):
        texts = list(texts)
    n = len(texts)
    offsets = np.empty(n + 1, dtype=np.int64)
    if out is not None and out.dtype == np.uint8 and out.flags["C_CONTIGUOUS"]:
        buf = out
    else:
        buf = np.empty(int(n * _ratio[0] * 1.25) + 4096, dtype=np.uint8)      # (untouched pages cost nothing)
    total = m.utf8_pack(texts, offsets.ctypes.data, buf.ctypes.data, buf.size)
    if total > buf.size:                                                      # did not fit: exact size, once more
        buf = np.empty(total, dtype=np.uint8)
        m.utf8_pack(texts, offsets.ctypes.data, buf.ctypes.data, total)
    if n:
        _ratio[0] = max(16.0, total / n)
    return buf[:total], offsets


def slice_tokens(texts: Sequence[str], spans: np.ndarray, tok_offsets: np.ndarray):
    """[[text[s:e].strip() ...] per string] from the span array of a batch (one C loop instead of a Python loop per token)."""
    m = load()
    if not isinstance(texts, (list, tuple)):
        texts = list(texts)
    spans = np.ascontiguousarray(spans)
    tok_offsets = np.ascontiguousarray(tok_offsets, dtype=np.int64)
    if spans.dtype not in (np.int32, np.uint16) or len(tok_offsets) != len(texts) + 1:
        raise ValueError("spans must be int32 or uint16 [T,2], tok_offsets int64[len(texts)+1]")
    if len(texts) and int(tok_offsets[-1]) > len(spans):
        raise ValueError("tok_offsets end beyond the span array")
    return m.slice_tokens(texts, spans.ctypes.data, tok_offsets.ctypes.data, 1 if spans.dtype == np.int32 else 0)
