"""GPU-backed stand-in for the reference's CPython extension module ``latok.latok``
(latok/core/src/latok/latok.c:373-378): the same three functions, the same int8 NumPy results,
computed by liblatok_b200.so on cuda:0.  Argument errors raise ValueError like the reference
(latok.c:40-50, 151-171, 292-312)."""
from __future__ import annotations

import numpy as np

from .engine import default_engine

__all__ = ["_gen_parse_matrix", "_gen_block_mask", "_combine_matrix_rows"]


def _gen_parse_matrix(*args) -> np.ndarray:
    """str -> int8[len(text), 25] feature matrix (gen_parse_matrix, latok.c:31-138)."""
    if len(args) < 1:
        raise ValueError("must specify string to generate the parse matrix for")   # latok.c:40-43
    text = args[0]
    if not isinstance(text, str):
        raise ValueError("Input string not in 'ready' state")                      # latok.c:47-50
    return default_engine().gen_parse_matrix(text)


def _gen_block_mask(*args) -> np.ndarray:
    """(a1, a2) -> int8 mask of ones with zeros over the a2-delimited blocks that hold an a1 mark
    (gen_block_mask, latok.c:140-258)."""
    if len(args) < 2:
        raise ValueError("must specify two aligning 1d numpy array args")          # latok.c:151-154
    return default_engine().gen_block_mask(args[0], args[1])


def _combine_matrix_rows(*args) -> np.ndarray:
    """(m, idxs) -> int8[m.shape[1]]: 2-D idxs = sum over rows of the product of the selected rows of m,
    1-D idxs = sum of the selected rows; -1 entries are ignored (combine_matrix_rows, latok.c:275-370)."""
    if len(args) < 2:
        raise ValueError("must specify 2d m and idxs matrices")                    # latok.c:292-295
    return default_engine().combine_matrix_rows(args[0], args[1])
