"""The reference's own Python API, served by the GPU library: same names, same arrays, same errors
(latok.latok extension functions, latok_utils, default_tokenizer).  Needs a B200 (-m gpu)."""
import json
import subprocess
import sys

import numpy as np
import pytest

import corpus
from oracle import oracle

pytestmark = pytest.mark.gpu


def test_gen_parse_matrix_and_split_mask():
    from latok_b200.core.default_tokenizer import gen_split_mask
    from latok_b200.core.latok_utils import gen_parse_matrix
    for t in corpus.FIXTURES[:30] + corpus.fuzz_strings(3, 40, 50):
        if not t:
            continue
        m = gen_parse_matrix(t)
        assert m.dtype == np.int8 and m.shape == (len(t), 25) and m.flags["C_CONTIGUOUS"]
        assert np.array_equal(m, oracle.parse_matrix(t)), t
        s = gen_split_mask(m)
        assert s.dtype == np.int8 and np.array_equal(s, oracle.split_mask(oracle.parse_matrix(t))), t


def test_extension_functions_on_raw_arrays():
    from latok_b200.latok import _combine_matrix_rows, _gen_block_mask
    rng = np.random.default_rng(0)
    for n in [1, 2, 5, 31, 32, 33, 100, 1023, 1024, 1025, 5000]:
        for p1, p2 in [(0.0, 0.2), (0.2, 0.0), (0.1, 0.2), (0.5, 0.05), (0.02, 0.5)]:
            a1 = (rng.random(n) < p1).astype(np.int8)
            a2 = (rng.random(n) < p2).astype(np.int8)
            assert np.array_equal(_gen_block_mask(a1, a2), oracle.block_mask(a1, a2)), (n, p1, p2)
        m = rng.integers(0, 3, size=(25, n)).astype(np.int8)
        for idx in oracle.DEFAULT_RULES:
            assert np.array_equal(_combine_matrix_rows(m, idx), oracle.combine_rows(m, idx))
        mt = np.ascontiguousarray(m.T)  # [n, 25]; transposed view has strides (1, 25) like the tokenizer's
        assert np.array_equal(_combine_matrix_rows(mt.T, oracle.C_MASK), oracle.combine_rows(mt.T, oracle.C_MASK))
        rows = rng.integers(0, min(n, 120), size=int(rng.integers(1, 9))).astype(np.int8)
        big = rng.integers(0, 2, size=(max(n, 1), 25)).astype(np.int8)
        assert np.array_equal(_combine_matrix_rows(big, rows), oracle.combine_rows(big, rows))
    # other integer dtypes are accepted for the block mask (the reference only tests != 0)
    a1 = np.array([0, 0, 3, 0, 0, 0], dtype=np.int64)
    a2 = np.array([0, 1, 0, 0, 1, 0], dtype=np.int32)
    assert _gen_block_mask(a1, a2).tolist() == oracle.block_mask((a1 != 0), (a2 != 0)).tolist()


def test_error_behaviour_matches_reference():
    from latok_b200.latok import _combine_matrix_rows, _gen_block_mask, _gen_parse_matrix
    from latok_b200.core.default_tokenizer import tokenize, featurize
    with pytest.raises(ValueError):
        _gen_parse_matrix()
    with pytest.raises(ValueError):
        _gen_block_mask(np.zeros(3, np.int8))
    with pytest.raises(ValueError):
        _gen_block_mask(np.zeros(3, np.int8), np.zeros(4, np.int8))
    with pytest.raises(ValueError):
        _gen_block_mask(np.zeros((3, 2), np.int8), np.zeros((3, 2), np.int8))
    with pytest.raises(ValueError):
        _combine_matrix_rows(np.zeros((3, 2), np.int8))
    with pytest.raises(ValueError):
        _combine_matrix_rows(np.zeros(3, np.int8), np.zeros(2, np.int8))
    with pytest.raises(ValueError):   # the reference segfaults here; the boundary validates
        _combine_matrix_rows(np.zeros((3, 2), np.int16), np.zeros(2, np.int8))
    with pytest.raises(IndexError):   # Q1
        list(tokenize(""))
    with pytest.raises(IndexError):
        list(featurize(""))
    assert _gen_parse_matrix("").shape == (0, 25)


def test_tokenize_and_featurize(golden_dir):
    from latok_b200.core.default_tokenizer import (featurize, featurize_batch, split_mask_batch, tokenize,
                                                   tokenize_batch)
    from latok_b200.core.latok_utils import FEATURE_NAMES, LaToken
    recs = json.load(open(golden_dir / "reference_outputs.json"))["records"]
    for r in recs[:25]:
        assert list(tokenize(r["text"])) == r["tokens"]
        if "feats" in r:
            toks = list(featurize(r["text"]))
            assert all(isinstance(t, LaToken) for t in toks)
            assert [t.text for t in toks] == r["feat_texts"]
            assert [[t.start_idx, t.end_idx] for t in toks] == r["feat_spans"]
            assert [t.features.tolist() for t in toks] == r["feats"]
            for t in toks:
                assert t.weight() == int(np.sum(t.features))
                assert set(t.feature_weights()) <= set(FEATURE_NAMES)
    texts = [r["text"] for r in recs]
    assert tokenize_batch(texts) == [r["tokens"] for r in recs]
    fb = featurize_batch(texts)
    for r, toks in zip(recs, fb):
        if "feats" in r:
            assert [t.features.tolist() for t in toks] == r["feats"]
    for r, s in zip(recs, split_mask_batch(texts)):
        assert s.tolist() == r["splits"]
    # Q5: feature sums keep counting past position 127 and wrap as uint8
    tok = list(featurize("A" * 250))
    assert len(tok) == 1 and tok[0].features[0] == np.int8(250 - 256)


def test_install_as_latok_in_subprocess():
    code = (
        "import latok_b200; latok_b200.install_as_latok();"
        "from latok.latok import _gen_parse_matrix, _gen_block_mask, _combine_matrix_rows;"
        "from latok.core.default_tokenizer import tokenize, featurize, gen_split_mask;"
        "from latok.core.latok_utils import gen_parse_matrix, build_combo_matrix, LaToken, FEATURE_NAMES;"
        "import latok.core.offsets as oft;"
        "print(list(tokenize('This is a #test! Testing, Testing, 1 2 3')))"
    )
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(corpus.__file__).rsplit("/tests/", 1)[0])
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == str(oracle.tokens("This is a #test! Testing, Testing, 1 2 3"))
