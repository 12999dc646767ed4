"""The default LaTok tokenizer on the GPU (mirror of the reference's
latok/core/default_tokenizer.py).

Single-string entry points keep the reference's names, arguments and results:

    gen_split_mask(m)   int8[L,25] feature matrix -> int8[L] split mask          (:113-134)
    tokenize(text)      generator of token strings                               (:137-160)
    featurize(text)     generator of LaToken(text, start_idx, end_idx, features)  (:163-191)

and the batch entry points are the ones that make sense on a GPU: a whole list of strings (or a
packed UTF-8 buffer + offsets) goes through one kernel pass and comes back as flat arrays.

    tokenize_batch(texts)  -> list[list[str]]
    featurize_batch(texts) -> list[list[LaToken]]
    split_mask_batch(texts)-> list[int8 arrays]
    batch_arrays(texts, ...) -> engine.BatchResult (splits, spans, CSR offsets, token features)
    tokenize_packed(buf, offsets) -> PackedTokens (trimmed token byte ranges; .tokens(i), .to_arrow())

Deliberate, documented deviations from the reference:
  * Q5: token feature vectors are the sum over *all* characters of the span; the reference's
    int8 row indices wrap/raise for positions >= 128.  Identical wherever the reference is defined.
  * Q1: the empty string raises IndexError in the single-string functions exactly like the
    reference (``splits[0] = 1`` on an empty array); in a batch it simply has no characters/tokens.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from . import offsets as oft
from .latok_utils import LaToken, build_combo_matrix, gen_block_mask
from ..latok import _combine_matrix_rows, _gen_parse_matrix
from ..engine import FEATS, MATRIX, SPANS, SPLITS, BatchResult, default_engine


def build_split_combo_matrix():
    """Split on whitespace, on a symbol, after a symbol, and at camelCase boundaries
    (an upper-case letter next to a lower-case one) -- default_tokenizer.py:39-55."""
    return build_combo_matrix([
        [oft.SPACE_IDX], [oft.SYMBOL_IDX], [oft.PREV_SYMBOL_IDX],
        [oft.UPPER_IDX, oft.NEXT_LOWER_IDX], [oft.UPPER_IDX, oft.PREV_LOWER_IDX],
    ])


def build_mask_combo_matrix():
    """Marks that protect a whitespace-delimited block from being split: #tag @user $X ^y after a
    space, ``.@user``, e-mail ``a@b`` and ``scheme://`` -- default_tokenizer.py:58-91."""
    return build_combo_matrix([
        [oft.TWITTER_IDX, oft.PREV_SPACE_IDX, oft.NEXT_ALPHA_IDX],
        [oft.CHAR_PERIOD_IDX, oft.PREV_SPACE_IDX, oft.NEXT_AT_IDX, oft.AFTER_NEXT_ALPHA_IDX],
        [oft.CHAR_AT_IDX, oft.PREV_ALPHA_NUM_IDX, oft.NEXT_ALPHA_NUM_IDX],
        [oft.CHAR_COLON_IDX, oft.NEXT_SLASH_IDX, oft.AFTER_NEXT_SLASH_IDX, oft.PREV_ALPHA_IDX],
    ])


def build_symbol_combo_matrix():
    """A symbol followed by whitespace still splits inside a protected block -- default_tokenizer.py:94-102."""
    return build_combo_matrix([[oft.SYMBOL_IDX, oft.NEXT_SPACE_IDX]])


C_SPLIT = build_split_combo_matrix()
C_MASK = build_mask_combo_matrix()
C_SYM = build_symbol_combo_matrix()


def gen_split_mask(m: np.ndarray) -> np.ndarray:
    """Feature matrix -> split mask, composed from the extension functions exactly as the
    reference composes them (default_tokenizer.py:113-134); every array op below runs on the GPU
    except the two int8 element-wise NumPy lines the reference also does in NumPy."""
    feats = m.T
    splits = _combine_matrix_rows(feats, C_SPLIT) * gen_block_mask(_combine_matrix_rows(feats, C_MASK),
                                                                   feats[oft.SPACE_IDX])
    splits += _combine_matrix_rows(feats, C_SYM)
    splits[0] = 1
    return splits


def batch_arrays(texts: Sequence[str], splits=True, spans=True, feats=False, matrix=False, engine=None) -> BatchResult:
    """One kernel pass over a list of strings; returns the flat result arrays."""
    what = (SPLITS if splits else 0) | (SPANS if spans else 0) | (FEATS if feats else 0) | (MATRIX if matrix else 0)
    return (engine or default_engine()).run(texts, what)


def _token_texts(text: str, spans: np.ndarray) -> List[str]:
    # `if token` as in the reference's loop (default_tokenizer.py:152-158): the kernel already drops the spans that
    # are one whitespace character, the guard keeps the two in step should a class table's SPACE set ever differ
    # from str.strip()'s (tests/test_host_cpu.py checks that they agree for the shipped table)
    return [tok for tok in (text[s:e].strip() for s, e in spans) if tok]


def tokenize_batch(texts: Sequence[str], engine=None) -> List[List[str]]:
    """[list(tokenize(t)) for t in texts]: one kernel pass, then the token strings cut by one C loop
    (`text[s:e].strip()`, dropped when empty, default_tokenizer.py:151-158)."""
    from .. import _pack
    if not isinstance(texts, (list, tuple)):
        texts = list(texts)
    r = batch_arrays(texts, splits=False, spans=True, engine=engine)
    return _pack.slice_tokens(texts, r.spans, r.tok_offsets)


class PackedTokens:
    """Tokens of a packed batch as byte ranges of its UTF-8 buffer: the reference's
    ``text[s:e].strip()`` loop (default_tokenizer.py:151-158) without creating Python strings.

    buf uint8[B]; byte_spans int64[T,2] (trimmed: ``buf[b:e]`` is the token text); tok_offsets int64[S+1]."""

    def __init__(self, buf, byte_spans, tok_offsets):
        self.buf, self.byte_spans, self.tok_offsets = buf, byte_spans, tok_offsets

    def __len__(self):
        return len(self.tok_offsets) - 1

    def tokens(self, i: int) -> List[str]:
        mv = memoryview(self.buf)
        return [bytes(mv[b:e]).decode("utf-8", "surrogatepass")
                for b, e in self.byte_spans[self.tok_offsets[i]:self.tok_offsets[i + 1]]]

    def to_arrow(self):
        """pyarrow ListArray<string>: one list of tokens per input string (token bytes gathered with NumPy)."""
        import pyarrow as pa
        ln = self.byte_spans[:, 1] - self.byte_spans[:, 0]
        off = np.zeros(len(ln) + 1, dtype=np.int64)
        np.cumsum(ln, out=off[1:])
        idx = np.repeat(self.byte_spans[:, 0] - off[:-1], ln) + np.arange(off[-1], dtype=np.int64)
        data = np.ascontiguousarray(self.buf[idx])
        if off[-1] < 2 ** 31 - 1:
            values = pa.StringArray.from_buffers(len(ln), pa.py_buffer(off.astype(np.int32)), pa.py_buffer(data))
            return pa.ListArray.from_arrays(pa.array(self.tok_offsets.astype(np.int32), type=pa.int32()), values)
        values = pa.LargeStringArray.from_buffers(len(ln), pa.py_buffer(off), pa.py_buffer(data))
        return pa.LargeListArray.from_arrays(pa.array(self.tok_offsets, type=pa.int64()), values)


def tokenize_packed(buf: np.ndarray, offsets: np.ndarray, engine=None) -> PackedTokens:
    """Packed UTF-8 in (flat uint8 buffer + int64 offsets[S+1]), token byte ranges out; no Python loop per token."""
    e = engine or default_engine()
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    with e._lock:                                    # submit .. fetch .. byte ranges of the same batch
        e.submit(buf, offsets, SPANS)
        r = e.fetch()
        return PackedTokens(buf, e.token_bytes(), r.tok_offsets)


def featurize_batch(texts: Sequence[str], engine=None) -> List[List[LaToken]]:
    r = batch_arrays(texts, splits=False, spans=True, feats=True, engine=engine)
    out = []
    for i, t in enumerate(texts):
        sp, ft = r.string_spans(i), r.string_feats(i)
        out.append([LaToken(tok, int(s), int(e), ft[k].copy())
                    for k, (s, e) in enumerate(sp) for tok in (t[s:e].strip(),) if tok])
    return out


def split_mask_batch(texts: Sequence[str], engine=None) -> List[np.ndarray]:
    r = batch_arrays(texts, splits=True, spans=False, engine=engine)
    return [r.string_splits(i) for i in range(len(texts))]


def tokenize(text: str):
    """Yield the tokens of ``text`` (default_tokenizer.py:137-160)."""
    if len(text) == 0:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")  # reference: splits[0] = 1, :132
    r = batch_arrays([text], splits=False, spans=True)
    for s, e in r.spans:
        token = text[s:e].strip()
        if token:                                    # default_tokenizer.py:152-158
            yield token


def featurize(text: str):
    """Yield a LaToken per token of ``text`` (default_tokenizer.py:163-191)."""
    if len(text) == 0:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    r = batch_arrays([text], splits=False, spans=True, feats=True)
    for k, (s, e) in enumerate(r.spans):
        token = text[s:e].strip()
        if token:                                    # default_tokenizer.py:175-191
            yield LaToken(token, int(s), int(e), r.tok_feats[k].copy())
