"""Batch front end: packs strings into (flat UTF-8 bytes, int64 offsets) and drives the C ABI.

One ``Engine`` = one GPU (one ``latok_b200_engine`` handle).  All arithmetic of the tokenization
path happens in liblatok_b200.so; this module only marshals NumPy buffers in and out.
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass
from typing import Iterable, Iterator, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import FEATS, MATRIX, SPANS, SPANS16, SPLITS, NUM_FEATURES


def pack_strings(texts: Sequence[str], out: Optional[np.ndarray] = None):
    """list[str] -> (uint8 buffer, int64 offsets[S+1]).  The packing loop runs in C (``_pypack``: reads the
    ``str`` objects in place like the reference's per-string entry, latok.c:47-55, and writes UTF-8 straight into
    the buffer).  Lone surrogates (legal in ``str``) are kept the way 'surrogatepass' encodes them.
    `out`: optional uint8 buffer (e.g. pinned memory) to pack into when it is large enough."""
    from . import _pack
    return _pack.pack(texts, out)


def pack_strings_python(texts: Sequence[str]):
    """The same result by a Python loop (kept as the check of the C packer and as its fallback-free reference)."""
    enc = [t.encode("utf-8", "surrogatepass") for t in texts]
    offsets = np.zeros(len(enc) + 1, dtype=np.int64)
    if enc:
        np.cumsum([len(b) for b in enc], out=offsets[1:])
    buf = np.frombuffer(b"".join(enc), dtype=np.uint8)
    return buf, offsets


@dataclass
class BatchResult:
    """Arrays for one batch (any field not requested at submit is None).

    splits[C] int8, char_offsets[S+1] int64, spans[T,2] int32 (string-relative, untrimmed, as in
    LaToken.start_idx/end_idx; uint16 with SPANS16), tok_offsets[S+1] int64, tok_feats[T,25] int8, matrix[C,25] int8.
    """
    n_strings: int
    n_chars: int
    n_tokens: int
    splits: Optional[np.ndarray] = None
    char_offsets: Optional[np.ndarray] = None
    spans: Optional[np.ndarray] = None
    tok_offsets: Optional[np.ndarray] = None
    tok_feats: Optional[np.ndarray] = None
    matrix: Optional[np.ndarray] = None
    kernel_ms: float = 0.0
    lookahead_walks: int = 0

    def string_splits(self, i: int) -> np.ndarray:
        return self.splits[self.char_offsets[i]:self.char_offsets[i + 1]]

    def string_spans(self, i: int) -> np.ndarray:
        return self.spans[self.tok_offsets[i]:self.tok_offsets[i + 1]]

    def string_feats(self, i: int) -> np.ndarray:
        return self.tok_feats[self.tok_offsets[i]:self.tok_offsets[i + 1]]

    def string_matrix(self, i: int) -> np.ndarray:
        return self.matrix[self.char_offsets[i]:self.char_offsets[i + 1]]


class Engine:
    """One GPU's tokenizer.  The handle serves one host thread at a time: every public method takes the engine's
    lock, and the compound calls (`run`, `run_packed`, `stream`, the three extension functions) hold it from submit
    to fetch, so threads that share an engine (e.g. the process-wide default one) cannot interleave."""

    def __init__(self, device: int = 0, max_batch_bytes: int = 0, max_strings: int = 0):
        self._L = _lib.load()
        h = C.c_void_p()
        _lib.check(self._L.latok_b200_create(device, max_batch_bytes, max_strings, C.byref(h)))
        self._h = h
        self.device = device
        self._lock = threading.RLock()
        self._flight = []            # (buf, offsets, n_strings, what) of the batches in flight, oldest first
        self._depth = 1

    def close(self):
        if getattr(self, "_h", None):
            self._L.latok_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- rules (build_combo_matrix layout, latok_utils.py:27-56) -------------------------------
    def set_rules(self, c_split=None, c_mask=None, c_sym=None):
        with self._lock:
            if c_split is None and c_mask is None and c_sym is None:
                _lib.check(self._L.latok_b200_set_rules(self._h, None, 0, 0, None, 0, 0, None, 0, 0))
                return
            mats = []
            for m in (c_split, c_mask, c_sym):
                if m is None:
                    raise ValueError("must specify split, mask and sym combo matrices")
                m = np.ascontiguousarray(m, dtype=np.int8)
                if m.ndim != 2:
                    raise ValueError("combo matrices must be 2d")
                mats.append(m)
            args = []
            for m in mats:
                args += [m.ctypes.data, m.shape[0], m.shape[1]]
            _lib.check(self._L.latok_b200_set_rules(self._h, *args))

    # ---- batch path ---------------------------------------------------------------------------
    def set_pipeline_depth(self, depth: int):
        """1: a submit replaces the batch in flight.  2: two batches in flight; fetch() reads the oldest, release()
        retires it (the in-library double buffering: copy-in + kernels of batch i+1 overlap copy-out of batch i)."""
        with self._lock:
            _lib.check(self._L.latok_b200_set_pipeline_depth(self._h, depth))
            self._depth = depth
            self._flight = []            # (the library drops the batches in flight)

    def _push(self, rec):
        if self._depth == 1:
            self._flight = [rec]
        else:
            self._flight.append(rec)

    def submit(self, buf: np.ndarray, offsets: np.ndarray, what: int = SPLITS | SPANS):
        buf = np.ascontiguousarray(buf, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        if offsets.ndim != 1 or len(offsets) < 1:
            raise ValueError("offsets must be a 1d array with n_strings + 1 entries")
        if int(offsets[-1]) > len(buf):
            raise ValueError("offsets end beyond the UTF-8 buffer")
        with self._lock:
            _lib.check(self._L.latok_b200_submit(self._h, buf.ctypes.data if len(buf) else None, offsets.ctypes.data,
                                                 len(offsets) - 1, what))
            self._push((buf, offsets, len(offsets) - 1, what | (SPANS if what & SPANS16 else 0)))

    def submit_device(self, d_utf8: int, d_offsets: int, n_strings: int, n_bytes: int, what: int = SPLITS | SPANS):
        """Text already in device memory (raw device pointers, e.g. tensor.data_ptr())."""
        with self._lock:
            _lib.check(self._L.latok_b200_submit_device(self._h, d_utf8, d_offsets, n_strings, n_bytes, what))
            self._push((None, None, n_strings, what | (SPANS if what & SPANS16 else 0)))

    def sizes(self):
        with self._lock:
            c, t = C.c_int64(0), C.c_int64(0)
            _lib.check(self._L.latok_b200_sizes(self._h, C.byref(c), C.byref(t)))
            return c.value, t.value

    def fetch(self, out: Optional[BatchResult] = None) -> BatchResult:
        """Results of the (oldest) batch in flight as fresh NumPy arrays."""
        with self._lock:
            if not self._flight:
                raise RuntimeError("no batch submitted")
            n_chars, n_tokens = self.sizes()
            _, _, S, w = self._flight[0]
            r = out or BatchResult(S, n_chars, n_tokens)
            r.n_strings, r.n_chars, r.n_tokens = S, n_chars, n_tokens
            if w & SPLITS:
                r.splits = np.empty(n_chars, dtype=np.int8)
            if w & (SPANS | FEATS):
                r.tok_offsets = np.empty(S + 1, dtype=np.int64)
            if w & SPANS:
                r.spans = np.empty((n_tokens, 2), dtype=np.uint16 if w & SPANS16 else np.int32)
            if w & FEATS:
                r.tok_feats = np.empty((n_tokens, NUM_FEATURES), dtype=np.int8)
            if w & MATRIX:
                r.matrix = np.empty((n_chars, NUM_FEATURES), dtype=np.int8)
            r.char_offsets = np.empty(S + 1, dtype=np.int64)

            def ptr(a):
                return None if a is None else a.ctypes.data
            _lib.check(self._L.latok_b200_fetch(self._h, n_chars, n_tokens, S, ptr(r.splits), ptr(r.char_offsets),
                                                ptr(r.spans), ptr(r.tok_offsets), ptr(r.tok_feats), ptr(r.matrix)))
            ms, walks = C.c_float(0), C.c_int64(0)
            _lib.check(self._L.latok_b200_last_stats(self._h, C.byref(ms), C.byref(walks)))
            r.kernel_ms, r.lookahead_walks = ms.value, walks.value
            return r

    def release(self):
        """Pipeline depth 2: retire the oldest batch in flight."""
        with self._lock:
            _lib.check(self._L.latok_b200_release(self._h))
            if self._flight:
                self._flight.pop(0)

    def token_bytes(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """int64 [T,2] byte ranges of the tokens of the (oldest) batch in flight in its flat UTF-8 buffer, trimmed like
        the reference's `text[s:e].strip()` (default_tokenizer.py:151-158): `buf[b:e]` is the token text."""
        with self._lock:
            if not self._flight or not self._flight[0][3] & SPANS:
                raise RuntimeError("spans were not requested at submit")
            _, n_tokens = self.sizes()
            if out is None:
                out = np.empty((n_tokens, 2), dtype=np.int64)
            if out.dtype != np.int64 or out.shape != (n_tokens, 2) or not out.flags["C_CONTIGUOUS"]:
                raise ValueError("out must be a C-contiguous int64 array of shape (n_tokens, 2)")
            _lib.check(self._L.latok_b200_fetch_token_bytes(self._h, n_tokens, out.ctypes.data if n_tokens else None, 0))
            return out

    def token_bytes_ms(self) -> float:
        ms = C.c_float(0)
        _lib.check(self._L.latok_b200_token_bytes_ms(self._h, C.byref(ms)))
        return ms.value

    def run(self, texts: Sequence[str], what: int = SPLITS | SPANS) -> BatchResult:
        buf, offsets = pack_strings(texts)
        return self.run_packed(buf, offsets, what)

    def run_packed(self, buf, offsets, what: int = SPLITS | SPANS) -> BatchResult:
        with self._lock:
            self.submit(buf, offsets, what)
            r = self.fetch()
            if self._depth == 2:
                self.release()
            return r

    def stream(self, batches: Iterable[Tuple[np.ndarray, np.ndarray]], what: int = SPLITS | SPANS) -> Iterator[BatchResult]:
        """Pipelined pass over a sequence of packed batches on one host thread: batch i+1 is staged, copied in and
        tokenized while the results of batch i are copied out (pipeline depth 2 inside the library)."""
        with self._lock:
            if self._flight and self._depth == 2:
                raise RuntimeError("batches are in flight")
            self.set_pipeline_depth(2)
            try:
                pending = 0
                for buf, offsets in batches:
                    self.submit(buf, offsets, what)
                    pending += 1
                    if pending == 2:
                        r = self.fetch()
                        self.release()
                        pending -= 1
                        yield r
                while pending:
                    r = self.fetch()
                    self.release()
                    pending -= 1
                    yield r
            finally:
                while self._flight:
                    self.release()
                self.set_pipeline_depth(1)

    # ---- measurement hooks ---------------------------------------------------------------------
    def timer_begin(self):
        _lib.check(self._L.latok_b200_timer_begin(self._h))

    def timer_end(self) -> float:
        ms = C.c_float(0)
        _lib.check(self._L.latok_b200_timer_end(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        n = C.c_int64(0)
        _lib.check(self._L.latok_b200_launch_count(self._h, C.byref(n)))
        return n.value

    # ---- the extension module's three functions ---------------------------------------------------
    def gen_parse_matrix(self, text: str) -> np.ndarray:
        b = text.encode("utf-8", "surrogatepass")
        out = np.empty((len(text), NUM_FEATURES), dtype=np.int8)
        n = C.c_int64(0)
        src = np.frombuffer(b, dtype=np.uint8)
        with self._lock:
            if self._depth == 2 and self._flight:
                raise RuntimeError("batches are in flight")
            _lib.check(self._L.latok_b200_gen_parse_matrix(self._h, src.ctypes.data if len(b) else None, len(b),
                                                           C.byref(n), out.ctypes.data if len(text) else None))
            self._flight = []
        if n.value != len(text):
            raise _lib.LatokCudaError(f"device decoded {n.value} characters, expected {len(text)}")
        return out

    def gen_block_mask(self, a1: np.ndarray, a2: np.ndarray) -> np.ndarray:
        a1 = np.asarray(a1)
        a2 = np.asarray(a2)
        if a1.ndim != 1 or a2.ndim != 1:
            raise ValueError("must specify 1d numpy array args")            # latok.c:157-160
        if a1.size != a2.size:
            raise ValueError("must specify 1d numpy arrays of matching length")  # latok.c:167-170
        x1 = (a1 != 0).astype(np.int8)   # the reference only tests for non-zero (PyArray_Nonzero, latok.c:178,198)
        x2 = (a2 != 0).astype(np.int8)
        out = np.empty(a1.size, dtype=np.int8)
        with self._lock:
            _lib.check(self._L.latok_b200_gen_block_mask(self._h, x1.ctypes.data, 1, x2.ctypes.data, 1, a1.size,
                                                         out.ctypes.data))
        return out

    def combine_matrix_rows(self, m: np.ndarray, idxs: np.ndarray) -> np.ndarray:
        m = np.asarray(m)
        idxs = np.asarray(idxs)
        if m.dtype != np.int8 or idxs.dtype != np.int8:
            # the reference dereferences NULL here; the boundary validates instead
            raise ValueError("m and idxs must be int8 arrays")
        if m.ndim != 2 or idxs.ndim > 2 or idxs.ndim < 1:
            raise ValueError("must specify 2d numpy array args")              # latok.c:309-312
        if any(s < 0 for s in m.strides):
            m = np.ascontiguousarray(m)
        idxs = np.ascontiguousarray(idxs)
        out = np.empty(m.shape[1], dtype=np.int8)
        ir, ic = (idxs.shape[0], idxs.shape[1]) if idxs.ndim == 2 else (idxs.shape[0], 0)
        if idxs.ndim == 2 and ic == 0:
            out[:] = 0
            return out
        with self._lock:
            _lib.check(self._L.latok_b200_combine_matrix_rows(self._h, m.ctypes.data, m.shape[0], m.shape[1],
                                                              m.strides[0], m.strides[1], idxs.ctypes.data, ir, ic,
                                                              out.ctypes.data))
        return out


_default = {}
_default_lock = threading.Lock()


def default_engine(device: int = 0) -> Engine:
    """Process-wide engine used by the drop-in single-string functions (shared by all threads; the engine's own
    lock serialises them)."""
    with _default_lock:
        e = _default.get(device)
        if e is None:
            e = _default[device] = Engine(device)
        return e
