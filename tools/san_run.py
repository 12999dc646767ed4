"""Small mixed workload for compute-sanitizer (memcheck / racecheck): fixtures, fuzz, long documents with space-free
runs and multi-mark chunks, both kernels.  python tools/san_run.py"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import corpus
from oracle import oracle
from latok_b200.engine import Engine

texts = list(corpus.FIXTURES) + corpus.fuzz_strings(3, 1500, 100, "marks") + corpus.fuzz_strings(4, 60, 3000, "mixed")
texts += ["x" * 9000 + " a@b" + ",q" * 4000 + " end", "a@b,c@d,e@f one,two three,four five,six " * 300, "日本語、テキスト。 café " * 900]
with Engine(0) as e:
    for what in (3, 7, 15):
        r = e.run(texts, what)
        o = oracle.tokenize_batch(texts, matrix=bool(what & 8), feats=bool(what & 4))
        assert np.array_equal(r.splits, o["splits"]) and np.array_equal(r.spans, o["spans"]), what
        assert np.array_equal(r.tok_offsets, o["tok_offsets"]) and np.array_equal(r.char_offsets, o["char_offsets"])
        print("what", what, "ok", r.n_chars, r.n_tokens)
    # token byte ranges (latok_tokbytes.cu)
    from latok_b200.core.default_tokenizer import tokenize_packed
    from latok_b200.engine import pack_strings
    keep = [t for t in texts if t]
    pt = tokenize_packed(*pack_strings(keep), engine=e)
    for i in (0, 5, len(keep) - 3, len(keep) - 2, len(keep) - 1):
        assert pt.tokens(i) == oracle.tokens(keep[i])
    print("token bytes ok", len(pt.byte_spans))
