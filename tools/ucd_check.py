#!/usr/bin/env python3
"""Parity of a library built for ANOTHER Unicode version (tools/regen_classes.py + gen_tables.py) against
an oracle built over the same class ranges: every code point (in strings of consecutive code points, with and without
separators) + fuzz, all outputs.  Run on a GPU box:

    LATOK_B200_LIB=<library> LATOK_ORACLE_LIB=<oracle .so> python tools/ucd_check.py

Building the pair for the UCD of the running Python (on the CPU box; put both files somewhere inside the repository so
that they travel to the GPU box, e.g. latok_b200/_variants/, which is git-ignored):

    python tools/regen_classes.py --out /tmp/ucd.txt
    LATOK_CLASSES=/tmp/ucd.txt LATOK_LOW_LIMIT=0x32400 LATOK_B200_LIB_OUT=$PWD/latok_b200/_variants/liblatok_ucd.so \
        python -m latok_b200.build                       # library with the new tables (then rebuild the default one!)
    LATOK_CLASSES=/tmp/ucd.txt LATOK_LOW_LIMIT=0x32400 python tools/gen_tables.py /tmp/gen
    mkdir -p /tmp/orc/_gen && cp oracle/latok_oracle.c /tmp/orc/ && cp /tmp/gen/oracle_runs.h /tmp/orc/_gen/
    gcc -O2 -shared -fPIC /tmp/orc/latok_oracle.c -o latok_b200/_variants/liblatok_oracle_ucd.so
    python -m latok_b200.build                           # back to the UCD-11 tables of the parity tests
"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import corpus
from oracle import oracle
from latok_b200.engine import Engine

cps = [c for c in range(0x110000) if not 0xD800 <= c <= 0xDFFF]
texts = ["".join(map(chr, cps[i:i + 997])) for i in range(0, len(cps), 997)]
texts += [" ".join(map(chr, cps[i:i + 500])) for i in range(0, len(cps), 500)]
texts += ["a".join(map(chr, cps[i:i + 301])) + " A@b " for i in range(0, len(cps), 301)]
texts += corpus.fuzz_strings(51, 3000, 200, "mixed")
texts = [t for t in texts if t]
o = oracle.tokenize_batch(texts, matrix=True, feats=True)
with Engine(0) as e:
    for what in (3, 7, 15):
        r = e.run(texts, what)
        assert np.array_equal(r.splits, o["splits"]) and np.array_equal(r.spans, o["spans"]), what
        assert np.array_equal(r.tok_offsets, o["tok_offsets"]) and np.array_equal(r.char_offsets, o["char_offsets"]), what
        if what & 4:
            assert np.array_equal(r.tok_feats, o["tok_feats"]), what
        if what & 8:
            assert np.array_equal(r.matrix, o["matrix"]), what
print(f"ucd check ok: {len(texts)} strings, {o['n_chars']} characters, {o['n_tokens']} tokens; "
      f"feature word of U+31350 (CJK extension H, new after UCD 11): {oracle.base_features(0x31350):03X}")
