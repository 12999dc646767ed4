"""Utilities of the LaTok API (mirror of the reference's latok/core/latok_utils.py): thin wrappers
over the extension functions, the combo-matrix builder, feature names and the LaToken record."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ..latok import _gen_block_mask, _gen_parse_matrix


def gen_parse_matrix(text: str) -> np.ndarray:
    """Feature matrix of a string: one row of 25 int8 features per character (latok_utils.py:10-15)."""
    return _gen_parse_matrix(text)


def gen_block_mask(a1: np.ndarray, a2: np.ndarray) -> np.ndarray:
    """Mask of ones with zeros between the ones of ``a2`` wherever ``a1`` has a one, the ends of
    ``a2`` counting as ones (latok_utils.py:18-24)."""
    return _gen_block_mask(a1, a2)


def build_combo_matrix(idx_lists) -> np.ndarray:
    """list of lists of feature indices -> int8 matrix padded with -1; the indices of a row are
    AND-ed (multiplied), the rows OR-ed (added) by ``_combine_matrix_rows`` (latok_utils.py:27-56)."""
    width = max(len(row) for row in idx_lists)
    combo = np.full((len(idx_lists), width), -1, dtype=np.int8)
    for r, row in enumerate(idx_lists):
        combo[r, :len(row)] = row
    return combo


# display names of the 25 feature columns, in column order (latok_utils.py:60-86)
FEATURE_NAMES = (
    "Alpha AlphaNum Num Lower Upper Space Symbol Twitter @ : / . "
    "Prev_Alpha Next_Alpha Prev_AlphaNum Next_AlphaNum Prev_Lower Next_Lower Prev_Space Next_Space "
    "Prev_Symbol Next_@ Next_/ After_Next_Alpha After_Next_/"
).split()
NUM_FEATURES = len(FEATURE_NAMES)


@dataclass
class LaToken:
    """A token: its text, its [start_idx, end_idx) character range in the source string and its
    feature vector, the per-feature sum over the token's characters (latok_utils.py:92-116)."""
    text: str
    start_idx: int
    end_idx: int
    features: np.ndarray

    def weight(self, weighting=None):
        """Sum of the (optionally weighted) features (latok_utils.py:106-110)."""
        return np.sum((self.features * weighting) if weighting else self.features)

    def feature_weights(self):
        """{feature name: weight} for the non-zero features (latok_utils.py:112-116)."""
        return {FEATURE_NAMES[i]: self.features[i] for i in np.nonzero(self.features)[0]}
