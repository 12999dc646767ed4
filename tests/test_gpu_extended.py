"""Extended alignment sweeps (only with LATOK_EXTENDED=1 on a GPU box; NOT part of `-m gpu`, skipped under `-m "not gpu"`): strings a few bytes longer or shorter than
every unit of work drift through all alignments when they are laid end to end -- the kind of input that exposed the
token-end bug fixed in round 1 (test_gpu_parity.py::test_token_covering_a_whole_range).  Written after the round's GPU
budget was spent, so they have not been run yet: run `LATOK_EXTENDED=1 pytest tests/test_gpu_extended.py` first thing next round and
move what passes into the `gpu` suite.
"""
import os

import numpy as np
import pytest

import corpus
from oracle import oracle
from test_gpu_parity import check_batch

pytestmark = [pytest.mark.gpu_extended,
              pytest.mark.skipif(not os.environ.get("LATOK_EXTENDED"), reason="extended GPU sweeps: set LATOK_EXTENDED=1 on a B200 box")]

RANGE, TILE5, STEP = 3968, 9 * 3968, 1024


@pytest.fixture(scope="module")
def engine():
    from latok_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _drift(ch, unit_chars, count, spread=12):
    """`count` strings of ch * n with n cycling through unit_chars - spread .. unit_chars + spread."""
    return [ch * (unit_chars - spread + (k % (2 * spread + 1))) for k in range(count)]


@pytest.mark.parametrize("ch,width", [("x", 1), ("é", 2), ("日", 3), ("\U00020000", 4), (",", 1), ("A", 1)])
def test_split_free_strings_around_a_range(engine, ch, width):
    # no split point inside (or, for ',' and 'A', a symbol / upper-case run): one token per string, ends drift over the
    # closer search windows of consecutive ranges
    texts = _drift(ch, RANGE // width, 600) + _drift(ch, 2 * RANGE // width, 300) + _drift(ch, STEP // width, 400, 5)
    check_batch(engine, texts, 7, label=f"drift {ch!r}")


def test_split_free_strings_around_a_tile(engine):
    texts = _drift("x", TILE5, 60, 20) + _drift("日", TILE5 // 3, 60, 20) + _drift("x", 7936, 120, 10)
    check_batch(engine, texts, 15, label="drift tile")


@pytest.mark.parametrize("sep", [" ", " a@b,c@d ", "　", " #tag "])
def test_one_separator_drifting_through_long_strings(engine, sep):
    # a single closer / multi-mark chunk inside an otherwise split-free string of about two ranges
    texts = []
    for k in range(500):
        n = 2 * RANGE + (k % 37) - 18
        cut = (k * 131) % n
        texts.append("x" * cut + sep + "y" * (n - cut))
    check_batch(engine, texts, 7, label=f"one separator {sep!r}")


def test_token_bytes_around_words_and_groups(engine):
    # token byte ranges: multi-byte strings around the 32-byte words and the 128 KB scan groups of latok_tokbytes.cu
    from latok_b200.core.default_tokenizer import tokenize_packed
    from latok_b200.engine import pack_strings
    rng = np.random.default_rng(9)
    alphabet = ["a", "é", "日", "\U0001F600", " ", ",", "B"]
    texts = ["".join(rng.choice(alphabet, size=n)) for n in list(range(1, 200, 3)) + [131072 // 2 + d for d in range(-3, 4)]]
    texts += ["é" * n + " z" for n in range(1, 70)] + ["日" * (43690 + d) + " q" for d in range(-2, 3)]
    texts = [t for t in texts if t]
    pt = tokenize_packed(*pack_strings(texts), engine=engine)
    for i, t in enumerate(texts):
        assert pt.tokens(i) == oracle.tokens(t), i
