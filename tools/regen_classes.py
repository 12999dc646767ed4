#!/usr/bin/env python3
"""Regenerate the LaTok character-class ranges from a NEWER Unicode Character Database (opt-in).

The reference builds its type table with scripts/unicode/makeunicodedata.py from UCD 11.0.0.  The part of that
script the tokenization path depends on is the flag assignment at makeunicodedata.py:158-200,249-258:

    ALPHA      category in Lm Lt Lu Ll Lo            LOWER / UPPER  derived properties Lowercase / Uppercase
    SPACE      category Zs or bidirectional WS, B, S  PRINTABLE      ' ' or category not C*, Z*
    NUMERIC    the record has a numeric value         TWITTER        @ # $ ^ ;  '@' ':' '/' '.' flag themselves

This tool applies the same rules to the UCD that ships with the running Python (`unicodedata`, 15.0.0 on
CPython 3.12) and writes a ranges file in the format of latok_b200/data/ucd11_latok_classes.txt:

    python tools/regen_classes.py --out latok_b200/data/ucd15_latok_classes.txt [--report]

The build keeps reading the UCD-11 file (bit-exact parity with the reference).  `--report` compares the result
with the UCD-11 file: every difference must be a code point the newer UCD (re)defined (UCD 15: 3 redefined,
11 794 newly assigned).  tools/gen_tables.py takes another ranges file through LATOK_CLASSES=<file>
LATOK_LOW_LIMIT=0x32400, but the packed two-stage table holds at most 256 distinct 128-code-point blocks (UCD 11
uses 254) and UCD 15 needs more: widening the stage-1 entries to 16 bits in the kernels is the remaining step
before a UCD-15 library can be built.
"""
import argparse
import sys
import unicodedata
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
UCD11 = ROOT / "latok_b200" / "data" / "ucd11_latok_classes.txt"


def base_features(cp: int) -> int:
    """12 base feature bits of one code point: the flags of makeunicodedata.py:158-200,249-258 projected the way
    latok.c:87-98 reads them (bit i = feature column i, offsets.py:24-35)."""
    ch = chr(cp)
    cat = unicodedata.category(ch)
    if cat == "Cn" or cat == "Cs" or cat == "Co":      # unassigned / surrogate / private use: no record -> no features
        return 0
    bidi = unicodedata.bidirectional(ch)
    alpha = cat in ("Lm", "Lt", "Lu", "Ll", "Lo")
    lower, upper = _has_prop_lower(ch), _has_prop_upper(ch)
    space = cat == "Zs" or bidi in ("WS", "B", "S")
    printable = cp == 0x20 or cat[0] not in ("C", "Z")
    numeric = unicodedata.numeric(ch, None) is not None
    alnum = alpha or numeric
    symbol = printable and not alnum and not space
    bits = [alpha, alnum, numeric, lower, upper, space, symbol, cp in (0x40, 0x23, 0x24, 0x5E), cp == 0x40, cp == 0x3A,
            cp == 0x2F, cp == 0x2E]
    return sum(1 << i for i, b in enumerate(bits) if b)


def _has_prop_lower(ch: str) -> bool:
    # str.islower() on one character is exactly the derived property Lowercase (Objects/unicodectype.c)
    return ch.islower()


def _has_prop_upper(ch: str) -> bool:
    return ch.isupper()


def runs(feat):
    out, start = [], 0
    for cp in range(1, 0x110000 + 1):
        if cp == 0x110000 or feat[cp] != feat[start]:
            out.append((start, cp - 1, feat[start]))
            start = cp
    return out


def read_ranges(path: Path):
    feat = [0] * 0x110000
    for line in path.read_text().splitlines():
        if not line or line.startswith("#"):
            continue
        a, b, f = line.split()
        for cp in range(int(a, 16), int(b, 16) + 1):
            feat[cp] = int(f, 16)
    return feat


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", help="ranges file to write")
    ap.add_argument("--report", action="store_true", help="compare with the UCD-11 ranges file")
    args = ap.parse_args()
    feat = [base_features(cp) for cp in range(0x110000)]
    rr = runs(feat)
    if args.out:
        with open(args.out, "w") as f:
            f.write(f"# LaTok base-feature classes per code point, UCD {unicodedata.unidata_version} + LaTok's 5 special flags.\n")
            f.write("# Derived data (tools/regen_classes.py); columns: first last features(hex, bit i = feature column i)\n")
            f.write(f"# runs={len(rr)} distinct_feature_words={len(set(x[2] for x in rr))}\n")
            for a, b, v in rr:
                f.write(f"{a:X} {b:X} {v:03X}\n")
    if args.report:
        old = read_ranges(UCD11)
        changed = [cp for cp in range(0x110000) if old[cp] != feat[cp]]
        newly = [cp for cp in changed if old[cp] == 0]
        print(f"UCD {unicodedata.unidata_version}: runs={len(rr)} classes={len(set(x[2] for x in rr))}; "
              f"{len(changed)} code points differ from UCD 11 ({len(newly)} of them had no features in UCD 11)")
        for cp in [c for c in changed if old[c] != 0][:20]:
            print(f"  U+{cp:04X} {unicodedata.name(chr(cp), '?')}: {old[cp]:03X} -> {feat[cp]:03X}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
